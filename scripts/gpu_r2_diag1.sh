#!/bin/bash
# round 2, call 1: reproduce the driver's exact command and localise the gate_up_proj slow launches
set -x
mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > gpurun_out/r2/smi0.txt
FAST="--awq-layers 0 --moe-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1"
# 1) the driver's command, as is (full legs) -- once
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench_driver_full.json 2> gpurun_out/r2/bench_driver_full.err
# 2) headline only, fresh process each, per-step dump, 10 ms clock sampling
for i in 1 2 3 4; do
  B200Q_BENCH_LMS=10 B200Q_BENCH_DIAG=gpurun_out/r2/diag_hold_$i.json python bench.py --gpus 1 --steps 20 --warmup 5 $FAST > gpurun_out/r2/b_hold_$i.json 2> gpurun_out/r2/b_hold_$i.err
done
for i in 1 2 3; do
  B200Q_BENCH_HOLD=0 B200Q_BENCH_LMS=10 B200Q_BENCH_DIAG=gpurun_out/r2/diag_nohold_$i.json python bench.py --gpus 1 --steps 20 --warmup 5 $FAST > gpurun_out/r2/b_nohold_$i.json 2> gpurun_out/r2/b_nohold_$i.err
done
# 3) longer run for comparison
B200Q_BENCH_LMS=10 B200Q_BENCH_DIAG=gpurun_out/r2/diag_200.json python bench.py --gpus 1 --steps 200 --warmup 5 $FAST > gpurun_out/r2/b_200.json 2> gpurun_out/r2/b_200.err
tail -n 3 gpurun_out/r2/*.err
