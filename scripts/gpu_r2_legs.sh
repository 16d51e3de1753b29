#!/bin/bash
# round 2: the strong legs that replay a CUDA graph per step (headline_strong, glm_fp8, moe_nvfp4) on N GPUs, other legs off
mkdir -p gpurun_out/r2
N=${1:-1}
FLAGS="$EXTRA --steps 5 --warmup 3 --awq-layers 0 --moe-awq-experts 0 --moe-block-experts 0 --glm-file-gb 0 --no-parity --no-cpu-baseline --e2e-steps 1"
if [ "$N" = 1 ]; then
  timeout 600 python bench.py --gpus 1 $FLAGS > gpurun_out/r2/legs_n1.json 2> gpurun_out/r2/legs_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N $FLAGS > gpurun_out/r2/legs_n$N.json 2> gpurun_out/r2/legs_n$N.err
fi
echo "rc=$?"; grep -v "Warning: \[PG" gpurun_out/r2/legs_n$N.err | tail -3
python - <<PY
import json
d=[json.loads(l) for l in open('gpurun_out/r2/legs_n$N.json') if l.startswith('{"metric"')][-1]
print(json.dumps(d['legs']))
PY
