#!/bin/bash
set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest6.log
tail -4 gpurun_out/r2/pytest6.log
for v in 8 9 10; do B200Q_AB_TAG=var$v B200Q_TMA_VAR=$v python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma4.jsonl; done
cat gpurun_out/r2/ab_tma4.jsonl
python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain6.log 2>&1; tail -1 gpurun_out/r2/awq_layer_plain6.log
python bench.py --gpus 1 --steps 20 --warmup 5 --awq-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1 --glm-units 0 --no-parity --no-strong > gpurun_out/r2/bench6_nvfp4.json 2> gpurun_out/r2/bench6_nvfp4.err
B200Q_FP4_NO_LOC=1 python bench.py --gpus 1 --steps 20 --warmup 5 --awq-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1 --glm-units 0 --no-parity --no-strong > gpurun_out/r2/bench6_nvfp4_noloc.json 2> gpurun_out/r2/bench6_nvfp4_noloc.err
python -c "
import json
for f in ('bench6_nvfp4','bench6_nvfp4_noloc'):
    d=json.load(open('gpurun_out/r2/%s.json'%f)); print(f, d['moe_nvfp4']['value'], d['moe_nvfp4']['roofline']['frac'])
"
# 2-GPU run of the whole bench (new legs: strong headline, parity, GLM with file leg)
