set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_r1_f.json 2> gpurun_out/bench_r1_f.err; tail -n 2 gpurun_out/bench_r1_f.err
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b200q -c 2000 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --awq-layers 1 --moe-layers 2 --moe-steps 2 --moe-awq-experts 2 > gpurun_out/ncu_f.log 2>&1
tail -n 2 gpurun_out/ncu_f.log | cut -c1-300
python scripts/ncu_kernels.py NVFP4 > gpurun_out/kp_NVFP4_f.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:nvfp4_fused --launch-skip 3 -c 1 -f -o gpurun_out/prof_r1f_NVFP4 python scripts/ncu_kernels.py NVFP4 > gpurun_out/ncu_NVFP4_f.log 2>&1
cat gpurun_out/kp_NVFP4_f.log
