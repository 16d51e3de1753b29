#!/bin/bash
# round 2, call 2: parity of the single-evaluation kernels + full-size tests vs live CT; A/B of the INT4 kernel variants under the
# driver's exact headline command; ncu instruction counts of the dominant kernel
set -x
mkdir -p gpurun_out/r2
python -m pytest tests -m gpu -x -q > gpurun_out/r2/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest2.log
tail -5 gpurun_out/r2/pytest2.log
FAST="--awq-layers 0 --moe-layers 0 --moe-awq-experts 0 --no-cpu-baseline --e2e-steps 1"
for v in one bracket atomic; do
  for i in 1 2; do
    case $v in
      one) E="" ;;
      bracket) E="B200Q_TMA_BRACKET=1" ;;
      atomic) E="B200Q_ZP_ATOMIC=1" ;;
    esac
    env $E B200Q_BENCH_DIAG=gpurun_out/r2/diag2_${v}_$i.json python bench.py --gpus 1 --steps 20 --warmup 5 $FAST > gpurun_out/r2/b2_${v}_$i.json 2> gpurun_out/r2/b2_${v}_$i.err
    grep per-launch gpurun_out/r2/b2_${v}_$i.err
  done
done
# other schemes (g32 sym etc.)
python scripts/bench_schemes.py > gpurun_out/r2/schemes_one.log 2>&1; cp gpurun_out/schemes.json gpurun_out/r2/schemes_one.json
B200Q_TMA_BRACKET=1 python scripts/bench_schemes.py > gpurun_out/r2/schemes_bracket.log 2>&1; cp gpurun_out/schemes.json gpurun_out/r2/schemes_bracket.json
# ncu: launch list + full set on the INT4 asym kernel
ncu --set full --clock-control none --import-source on -k regex:group_tma -c 2 -o gpurun_out/r2/ncu_int4_one -f python bench.py --gpus 1 --steps 1 --warmup 3 $FAST > gpurun_out/r2/ncu_one.log 2>&1
ncu -i gpurun_out/r2/ncu_int4_one.ncu-rep --page raw --csv > gpurun_out/r2/ncu_int4_one_raw.csv 2>/dev/null
ls -la gpurun_out/r2 | tail -5
