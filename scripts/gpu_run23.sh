set -x
( time timeout 600 python -m pytest tests -m gpu -q ) 2>&1 | tail -6
( time timeout 300 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err ) 2>&1 | tail -4
tail -n 2 gpurun_out/bench_final2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final2.json')); print(d['value'], d['roofline']['frac'], d['awq']['value'], d['awq']['roofline']['frac'], d['moe_nvfp4']['value'], d['e2e']['value'])"
