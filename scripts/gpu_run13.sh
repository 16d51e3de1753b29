for shape in 32,2560,9728 512,768,2048; do
for c in 3 4 5 8; do
echo "== shape $shape CTAS $c"
B200Q_BENCH_SHAPE=$shape B200Q_DECODE_CTAS=$c B200Q_DECODE_INT4_PRE=e$c python scripts/bench_decompress.py W4A16 W4A16_ASYM INT4_G32_SYM NVFP4 2>&1 | tail -4
done
done
