#!/bin/bash
set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_awq.py -x -q -k "attention_core" > gpurun_out/r2/pytest8_attn.log 2>&1; echo "rc=$?" >> gpurun_out/r2/pytest8_attn.log
tail -25 gpurun_out/r2/pytest8_attn.log
timeout 120 python scripts/bench_attn.py > gpurun_out/r2/bench_attn.log 2>&1; cat gpurun_out/r2/bench_attn.log
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest8.log
tail -5 gpurun_out/r2/pytest8.log
timeout 200 python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain8.log 2>&1; tail -1 gpurun_out/r2/awq_layer_plain8.log
B200Q_ATTN_SDPA=1 timeout 200 python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain8_sdpa.log 2>&1; tail -1 gpurun_out/r2/awq_layer_plain8_sdpa.log
timeout 300 python scripts/bench_schemes.py > gpurun_out/r2/schemes8.log 2>&1; cp gpurun_out/schemes.json gpurun_out/r2/schemes8.json; cat gpurun_out/r2/schemes8.log
