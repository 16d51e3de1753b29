for c in 2 3 4 6 8; do
echo "== CTAS $c"
B200Q_DECODE_CTAS=$c B200Q_DECODE_INT4_PRE=d$c python scripts/bench_decompress.py W4A16 INT4_G32_SYM FP8_BLOCK FP8_G32 FP8_CHANNEL 2>&1 | tail -5
done
