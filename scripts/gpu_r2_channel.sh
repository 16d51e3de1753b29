#!/bin/bash
# round 2: single-evaluation channel kernel -- parity tests of every CHANNEL format, then the scheme table (default and B200Q_CHANNEL_BRACKET=1)
mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests -m gpu -q -x -k "channel or fullsize_bit_exact or registered or unfused or golden" 2>&1 | tail -3
timeout 300 python scripts/bench_schemes.py 2>&1 | grep "CHANNEL"; cp gpurun_out/schemes.json gpurun_out/r2/schemes_channel_one.json
B200Q_CHANNEL_BRACKET=1 timeout 300 python scripts/bench_schemes.py 2>&1 | grep "CHANNEL"
