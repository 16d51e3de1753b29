#!/bin/bash
# round 2: single-evaluation FP8 128x128 block kernel -- parity tests, then the headline step with and without it
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x -k "block or fp8 or arena or fullsize" 2>&1 | tail -3
FLAGS="--gpus 1 --steps 20 --warmup 5 --awq-layers 0 --moe-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1 --no-parity"
for v in "" "B200Q_FP8_BLOCK_BRACKET=1"; do
  env $v python bench.py $FLAGS 2>/dev/null | python -c "
import sys, json
d = [json.loads(l) for l in sys.stdin if l.startswith('{')][-1]
pc = d['roofline']['per_class']
print('$v', round(d['value']), round(d['roofline']['frac'], 3), {k: round(v['GBps_algorithmic']) for k, v in pc.items()}, json.dumps({k: d['legs'][k]['frac'] for k in d['legs'] if k.startswith('glm') or k.startswith('headline')}))"
done
