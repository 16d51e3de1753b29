#!/bin/bash
# round 2, second half: the whole GPU suite, smoke(), the default bench line and the NVFP4 ncu capture on the final tree
mkdir -p gpurun_out/r2
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2/pytest_f2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2/pytest_f2.log; tail -3 gpurun_out/r2/pytest_f2.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2/smoke_f2.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2/smoke_f2.log
(time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5) > gpurun_out/r2/bench_f2.json 2> gpurun_out/r2/bench_f2.err; tail -5 gpurun_out/r2/bench_f2.err
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2/bench_f2.json') if l.startswith('{')][-1]); print(round(d['value']), round(d['roofline']['frac'],3), d['clocks']); print(json.dumps(d['legs'])); print(d['parity'])
"
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k regex:nvfp4_fused --launch-skip 3 -c 1 -o gpurun_out/r2/prof_nvfp4_v3 python scripts/ncu_kernels.py NVFP4 > gpurun_out/r2/prof_nvfp4_v3.log 2>&1
ncu -i gpurun_out/r2/prof_nvfp4_v3.ncu-rep --page raw --csv > gpurun_out/r2/prof_nvfp4_v3_raw.csv 2>/dev/null
rm -f gpurun_out/r2/prof_nvfp4_v3.ncu-rep
python scripts/ncu_kernels.py NVFP4 2>&1 | tail -1
