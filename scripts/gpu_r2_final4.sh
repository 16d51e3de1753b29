#!/bin/bash
# round 2, last pass: INT8 per-channel fast path tests, every-scheme table, then the whole GPU suite + smoke on the final tree
mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_compress.py tests/test_gpu_fullsize_vs_ct.py -m gpu -q -x -k "int8_channel or channel_kernel" 2>&1 | tail -3
timeout 300 python scripts/bench_schemes.py 2>&1 | tail -14; cp gpurun_out/schemes.json gpurun_out/r2/schemes_final.json
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest_f4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_f4.log; tail -3 gpurun_out/r2/pytest_f4.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_f4.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2/smoke_f4.log
