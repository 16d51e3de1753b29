set -x
B200Q_QPACK=rows timeout 600 python -m pytest tests/test_gpu_compress.py -m gpu -q -k "quantize_pack or unfused or fullsize" 2>&1 | tail -3
B200Q_QPACK=tma timeout 600 python -m pytest tests/test_gpu_compress.py -m gpu -q -k "quantize_pack or unfused" 2>&1 | tail -3
for m in rows tma; do B200Q_QPACK=$m B200Q_BENCH_TAG=$m python scripts/bench_qparams_paths.py W4A16 W4A16_ASYM INT4_G32_SYM FP8_G32 2>&1 | grep quantize_pack; done
