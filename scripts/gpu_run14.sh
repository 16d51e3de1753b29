set -x
( time timeout 1000 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
python scripts/bench_decompress.py 2>&1 | tail -9
B200Q_BENCH_SHAPE=32,2560,9728 B200Q_DECODE_INT4_PRE=s2 python scripts/bench_decompress.py 2>&1 | tail -9
