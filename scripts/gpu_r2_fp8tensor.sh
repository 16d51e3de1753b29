#!/bin/bash
# round 2: bf16 fast path of per-tensor FP8 -- parity, then the scheme table line
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x -k "tensor or fp8 or unfused or registered or stack" 2>&1 | tail -3
timeout 300 python scripts/bench_schemes.py 2>&1 | grep "FP8"; cp gpurun_out/schemes.json gpurun_out/r2/schemes_fp8tensor.json
