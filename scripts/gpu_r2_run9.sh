#!/bin/bash
set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_awq.py -x -q -k "attention_core" > gpurun_out/r2/pytest9_attn.log 2>&1; echo "rc=$?" >> gpurun_out/r2/pytest9_attn.log
tail -8 gpurun_out/r2/pytest9_attn.log
timeout 120 python scripts/bench_attn.py > gpurun_out/r2/bench_attn9.log 2>&1; cat gpurun_out/r2/bench_attn9.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest9.log
tail -5 gpurun_out/r2/pytest9.log
timeout 200 python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain9.log 2>&1; tail -1 gpurun_out/r2/awq_layer_plain9.log
B200Q_ATTN_SDPA=1 timeout 200 python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain9_sdpa.log 2>&1; tail -1 gpurun_out/r2/awq_layer_plain9_sdpa.log
