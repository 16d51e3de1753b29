#!/bin/bash
set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest4.log
tail -6 gpurun_out/r2/pytest4.log
for v in 0 ; do B200Q_AB_TAG=var$v-waitmoved B200Q_TMA_VAR=$v python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma2.jsonl; done
B200Q_AB_TAG=bracket-waitmoved B200Q_TMA_BRACKET=1 python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma2.jsonl
tail -2 gpurun_out/r2/ab_tma2.jsonl
python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain4.log 2>&1; tail -2 gpurun_out/r2/awq_layer_plain4.log
(time python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2/bench4.json 2> gpurun_out/r2/bench4.err; tail -5 gpurun_out/r2/bench4.err
(time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2/bench4_ref.json 2> gpurun_out/r2/bench4_ref.err; tail -4 gpurun_out/r2/bench4_ref.err
B200Q_TMA_BRACKET=1 python bench.py --gpus 1 --steps 20 --warmup 5 --awq-layers 0 --moe-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1 --glm-units 0 --no-parity > gpurun_out/r2/bench4_bracket.json 2> gpurun_out/r2/bench4_bracket.err
