set -x
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -12
python scripts/bench_qparams_paths.py NVFP4 2>&1 | tail -3
