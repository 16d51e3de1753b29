#!/bin/bash
# round 2: expert-parallel layer-wide MoE mapping (ring-ordered combine) on N GPUs: single-device chain test, then the bench leg + parity
mkdir -p gpurun_out/r2
N=${1:-2}
[ -n "$SKIP_PYTEST" ] || timeout 600 python -m pytest tests/test_gpu_awq.py -m gpu -q -x -k "moe or routed" 2>&1 | tail -3
FLAGS="--steps 2 --warmup 3 --awq-layers 0 --moe-layers 0 --moe-awq-experts $N --moe-block-experts ${2:-256} --glm-units 0 --no-cpu-baseline --e2e-steps 1 --no-strong"
for mode in ${MODES-"" "--moe-block-token-sharded"}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N $FLAGS $mode > gpurun_out/r2/ep_n$N$mode.json 2> gpurun_out/r2/ep_n$N$mode.err
  echo "rc=$? mode=$mode"; tail -3 gpurun_out/r2/ep_n$N$mode.err
  python - <<PY
import json
for line in open('gpurun_out/r2/ep_n$N$mode.json'):
    if line.startswith('{"metric"'):
        d = json.loads(line)
        print(json.dumps(d['legs'].get('moe_awq_layer_mapping')), d.get('parity'))
PY
done
