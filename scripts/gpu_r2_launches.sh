#!/bin/bash
# round 2, final tree: UNFILTERED ncu launch list of a reduced bench.py run (all legs, library kernels included)
mkdir -p gpurun_out/r2
FLAGS="--gpus 1 --steps 2 --warmup 3 --awq-layers 1 --moe-layers 2 --moe-steps 2 --moe-awq-experts 2 --moe-block-experts 8 --glm-units 8 --glm-file-gb 0 --no-cpu-baseline --e2e-steps 1 --no-parity"
timeout 300 python bench.py $FLAGS > gpurun_out/r2/launches_bench_plain.log 2>&1; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2/launches_bench_final.csv python bench.py $FLAGS > gpurun_out/r2/launches_bench_final.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2/launches_bench_final.csv')) if len(r) > 10]
hdr = rows[0]; ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    try:
        agg[r[ki][:80]][0] += 1; agg[r[ki][:80]][1] += float(r[vi].replace(',', ''))
    except Exception:
        pass
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{v[1] / tot * 100:6.2f} %  {v[0]:5d} x  {k}")
PY
du -sh gpurun_out/r2/launches_bench_final.csv
