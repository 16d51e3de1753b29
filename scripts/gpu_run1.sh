set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py > gpurun_out/bench_r1_b.json 2> gpurun_out/bench_r1_b.err; tail -3 gpurun_out/bench_r1_b.err; cat gpurun_out/bench_r1_b.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
