set -x
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b200q -c 400 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
for k in W4A16_ASYM FP8_BLOCK NVFP4 INT4_G32_SYM; do
  python scripts/ncu_kernels.py $k > gpurun_out/kp_$k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:b200q --launch-skip 3 -c 1 -f -o gpurun_out/prof_r1_$k python scripts/ncu_kernels.py $k > gpurun_out/ncu_$k.log 2>&1
  cat gpurun_out/kp_$k.log
done
python scripts/ncu_awq_gemm.py > gpurun_out/kp_awq_gemm.log 2>&1 && \
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:awq_gemm_loss_kernel --launch-skip 1 -c 1 -f -o gpurun_out/prof_r1_awq_gemm python scripts/ncu_awq_gemm.py > gpurun_out/ncu_awq_gemm.log 2>&1
tail -3 gpurun_out/ncu_awq_gemm.log
timeout 300 python scripts/bench_awq_gemm.py 32768 2>&1 | tail -4
ls -la gpurun_out
