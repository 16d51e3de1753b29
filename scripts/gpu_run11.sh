set -x
for c in 3 4 5 6; do
B200Q_DECODE_CTAS=$c B200Q_DECODE_INT4_PRE=c$c python scripts/bench_decompress.py W4A16 W4A16_ASYM INT4_G32_SYM NVFP4 FP8_BLOCK FP8_CHANNEL 2>&1 | tail -6
done
