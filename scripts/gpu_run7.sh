set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 1200 python bench.py > gpurun_out/bench_r1_k.json 2> gpurun_out/bench_r1_k.err; tail -n 2 gpurun_out/bench_r1_k.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r1_k_ref.json 2> gpurun_out/bench_r1_k_ref.err; tail -c 600 gpurun_out/bench_r1_k_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b200q -c 3000 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --awq-layers 1 --moe-layers 2 --moe-steps 2 --moe-awq-experts 2 --moe-block-experts 8 > gpurun_out/ncu_k.log 2>&1
tail -n 1 gpurun_out/ncu_k.log | cut -c1-200
for k in FP8_CHANNEL NVFP4; do
  python scripts/ncu_kernels.py $k > gpurun_out/kp_${k}_k.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:b200q --launch-skip 3 -c 1 -f -o gpurun_out/prof_r1k_$k python scripts/ncu_kernels.py $k > gpurun_out/ncu_${k}_k.log 2>&1
  cat gpurun_out/kp_${k}_k.log
done
