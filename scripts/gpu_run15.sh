set -x
( time timeout 1000 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -8
python scripts/bench_observers.py 2>&1 | tail -11
B200Q_MINMAX_LEGACY=1 B200Q_WMEAN_LEGACY=1 B200Q_BENCH_TAG=legacy python scripts/bench_observers.py 2>&1 | tail -11
