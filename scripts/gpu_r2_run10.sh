#!/bin/bash
set -x
mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_awq.py -x -q -k "attention_core" > gpurun_out/r2/pytest10_attn.log 2>&1; echo "rc=$?" >> gpurun_out/r2/pytest10_attn.log
tail -8 gpurun_out/r2/pytest10_attn.log
timeout 120 python scripts/bench_attn.py > gpurun_out/r2/bench_attn10.log 2>&1; cat gpurun_out/r2/bench_attn10.log
timeout 300 python -m pytest tests/test_gpu_awq.py tests/test_gpu_awq_fullsize.py -x -q > gpurun_out/r2/pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest10.log
tail -4 gpurun_out/r2/pytest10.log
timeout 200 python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain10.log 2>&1; tail -1 gpurun_out/r2/awq_layer_plain10.log
B200Q_ATTN_SDPA=1 timeout 200 python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain10_sdpa.log 2>&1; tail -1 gpurun_out/r2/awq_layer_plain10_sdpa.log
