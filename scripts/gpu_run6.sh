set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_r1_j.json 2> gpurun_out/bench_r1_j.err; tail -n 2 gpurun_out/bench_r1_j.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b200q -c 3000 --csv --log-file gpurun_out/launches_r1j.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --awq-layers 1 --moe-layers 2 --moe-steps 2 --moe-awq-experts 2 --moe-block-experts 8 > gpurun_out/ncu_j.log 2>&1
tail -n 1 gpurun_out/ncu_j.log | cut -c1-200
for k in W4A16_ASYM; do
  python scripts/ncu_kernels.py $k > gpurun_out/kp_${k}_j.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:b200q --launch-skip 3 -c 1 -f -o gpurun_out/prof_r1j_$k python scripts/ncu_kernels.py $k > gpurun_out/ncu_${k}_j.log 2>&1
  cat gpurun_out/kp_${k}_j.log
done
