#!/bin/bash
set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest5.log
tail -4 gpurun_out/r2/pytest5.log
for v in 0 6 7; do B200Q_AB_TAG=var$v B200Q_TMA_VAR=$v python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma3.jsonl; done
B200Q_AB_TAG=bracket B200Q_TMA_BRACKET=1 python scripts/ab_tma2.py 2>&1 | tail -1 >> gpurun_out/r2/ab_tma3.jsonl
cat gpurun_out/r2/ab_tma3.jsonl
python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_plain5.log 2>&1; tail -1 gpurun_out/r2/awq_layer_plain5.log
FAST="--awq-layers 0 --moe-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1 --glm-units 0 --no-parity --no-strong"
B200Q_TMA_BRACKET=1 ncu --set full --clock-control none --import-source on -k regex:group_tma -c 2 -o gpurun_out/r2/ncu_int4_bracket -f python bench.py --gpus 1 --steps 1 --warmup 3 $FAST > gpurun_out/r2/ncu_bracket.log 2>&1
ncu -i gpurun_out/r2/ncu_int4_bracket.ncu-rep --page raw --csv > gpurun_out/r2/ncu_int4_bracket_raw.csv 2>/dev/null
ncu --set full --clock-control none --import-source on -k regex:group_tma -c 2 -o gpurun_out/r2/ncu_int4_one2 -f python bench.py --gpus 1 --steps 1 --warmup 3 $FAST > gpurun_out/r2/ncu_one2.log 2>&1
ncu -i gpurun_out/r2/ncu_int4_one2.ncu-rep --page raw --csv > gpurun_out/r2/ncu_int4_one2_raw.csv 2>/dev/null
ls -la gpurun_out/r2 | tail -4
