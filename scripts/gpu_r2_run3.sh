#!/bin/bash
set -x
mkdir -p gpurun_out/r2
for v in 0 1 2 3 4 5; do B200Q_AB_TAG=var$v B200Q_TMA_VAR=$v python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma2.jsonl; done
B200Q_AB_TAG=bracket B200Q_TMA_BRACKET=1 python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma2.jsonl
cat gpurun_out/r2/ab_tma2.jsonl
# AWQ: parity of the ping-pong loss GEMM, then timing + unfiltered launch list
python -m pytest tests/test_gpu_awq.py -x -q > gpurun_out/r2/pytest3_awq.log 2>&1; tail -3 gpurun_out/r2/pytest3_awq.log
python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain.log 2>&1; tail -2 gpurun_out/r2/awq_layer_plain.log
python scripts/bench_awq_gemm.py > gpurun_out/r2/awq_gemm.log 2>&1; tail -12 gpurun_out/r2/awq_gemm.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/awq_layer_launches.csv python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_ncu.log 2>&1
tail -2 gpurun_out/r2/awq_layer_ncu.log
wc -l gpurun_out/r2/awq_layer_launches.csv
