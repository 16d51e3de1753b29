set -x
( time timeout 1000 python -m pytest tests -m gpu -q ) 2>&1 | tail -25
python scripts/bench_qparams_paths.py W4A16_ASYM NVFP4 2>&1 | tail -8
