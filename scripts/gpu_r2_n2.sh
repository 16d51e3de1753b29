#!/bin/bash
set -x
mkdir -p gpurun_out/r2
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2/bench_n$N.json 2> gpurun_out/r2/bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/r2/bench_n$N.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2/bench_n$N.json') if l.startswith('{\"metric\"')][-1]; print(json.dumps(d['legs'])); print(d['parity']); print(d['e2e'])
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5 --warmup 1 --cpu-awq-tokens 0 > gpurun_out/r2/bench_ref_n$N.json 2> gpurun_out/r2/bench_ref_n$N.err; echo "ref rc=$?"; head -c 400 gpurun_out/r2/bench_ref_n$N.json
