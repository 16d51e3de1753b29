#!/bin/bash
# round 2, closing check on the final tree: whole GPU suite + smoke()
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest_f6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_f6.log; tail -3 gpurun_out/r2/pytest_f6.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_f6.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2/smoke_f6.log
