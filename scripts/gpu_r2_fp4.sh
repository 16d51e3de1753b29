# round 2: fused NVFP4 kernel A/B: next-tile prefetch variants (one process per variant)
mkdir -p gpurun_out/r2
for v in "" "B200Q_FP4_MINB=14" "B200Q_FP4_MINB=13" "B200Q_FP4_MINB=14 B200Q_FP4_NT=8"; do
  env $v B200Q_AB_TAG="${v:-default}" timeout 300 python scripts/ab_fp4_v2.py 2>&1 | tail -1
done | tee gpurun_out/r2/ab_fp4_v6.log
