set -x
( time timeout 1000 python -m pytest tests -m gpu -x -q --durations=8 ) 2>&1 | tail -16
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python scripts/bench_decompress.py 2>&1 | tail -9
B200Q_DECODE_INT4_PRE=none python scripts/bench_decompress.py W4A16 W4A16_ASYM INT4_G32_SYM INT4_G32_ASYM 2>&1 | tail -4
B200Q_DECODE_INT4_PRE=all python scripts/bench_decompress.py W4A16 W4A16_ASYM INT4_G32_SYM INT4_G32_ASYM 2>&1 | tail -4
