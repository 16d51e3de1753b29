#!/bin/bash
set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -q > gpurun_out/r2/pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest7.log
tail -6 gpurun_out/r2/pytest7.log
python scripts/bench_schemes.py > gpurun_out/r2/schemes7.log 2>&1; cp gpurun_out/schemes.json gpurun_out/r2/schemes7.json; cat gpurun_out/r2/schemes7.log
for w in 7 9 10; do B200Q_AB_TAG=warps$w B200Q_TMA_WARPS=$w python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma5.jsonl; done
B200Q_AB_TAG=default python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma5.jsonl
cat gpurun_out/r2/ab_tma5.jsonl
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2/bench7.json 2> gpurun_out/r2/bench7.err; tail -3 gpurun_out/r2/bench7.err
python -c "
import json
d=json.load(open('gpurun_out/r2/bench7.json')); print(json.dumps(d['legs'])); print(d['parity'])
"
