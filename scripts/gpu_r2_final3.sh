#!/bin/bash
# round 2, last pass: new graph / concurrent-class test, the default bench line at N = 1 (profiles/r2_bench.json)
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_fullsize_vs_ct.py -m gpu -q -k "concurrent or arena" 2>&1 | tail -3
(time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2/bench_f3.json 2> gpurun_out/r2/bench_f3.err; tail -5 gpurun_out/r2/bench_f3.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_f3.json') if l.startswith('{')][-1]); print(round(d['value']), round(d['roofline']['frac'],3), d['clocks']); print(json.dumps(d['legs'])); print(d['parity'])
"
