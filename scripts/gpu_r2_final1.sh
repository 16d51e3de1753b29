#!/bin/bash
set -x
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest_f1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_f1.log; tail -3 gpurun_out/r2/pytest_f1.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_f1.log 2>&1; tail -2 gpurun_out/r2/smoke_f1.log
(time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2/bench_f1.json 2> gpurun_out/r2/bench_f1.err; tail -6 gpurun_out/r2/bench_f1.err
(time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2/bench_f1_ref.json 2> gpurun_out/r2/bench_f1_ref.err; tail -4 gpurun_out/r2/bench_f1_ref.err
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --awq-layers 0 --moe-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1 --glm-units 0 --no-parity --no-strong > gpurun_out/r2/bench_f1_b.json 2> gpurun_out/r2/bench_f1_b.err
timeout 900 python bench.py --gpus 1 --steps 200 --warmup 5 --awq-layers 0 --moe-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1 --glm-units 0 --no-parity --no-strong > gpurun_out/r2/bench_f1_200.json 2> gpurun_out/r2/bench_f1_200.err
python -c "
import json
for f in ('bench_f1','bench_f1_b','bench_f1_200'):
    d=json.loads([l for l in open('gpurun_out/r2/%s.json'%f) if l.startswith('{')][-1]); print(f, round(d['value']), round(d['roofline']['frac'],3), d['clocks'])
d=json.loads([l for l in open('gpurun_out/r2/bench_f1.json') if l.startswith('{')][-1]); print(json.dumps(d['legs'])); print(d['parity'])
"
