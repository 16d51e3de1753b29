#!/bin/bash
# round 2, closing pass on the final tree: whole GPU suite, smoke(), the driver's default bench command
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest_f5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_f5.log; tail -3 gpurun_out/r2/pytest_f5.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_f5.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2/smoke_f5.log
(time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2/bench_f5.json 2> gpurun_out/r2/bench_f5.err; tail -4 gpurun_out/r2/bench_f5.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_f5.json') if l.startswith('{')][-1]); print(round(d['value']), round(d['roofline']['frac'],3), d['clocks']); print(json.dumps(d['legs'])); print(d['parity']); print(d['roofline']['per_class'])
"
