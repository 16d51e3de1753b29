ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:b200q -c 3000 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --awq-layers 1 --moe-layers 2 --moe-steps 2 --moe-awq-experts 2 --moe-block-experts 4 > gpurun_out/ncu_h.log 2>&1
tail -n 1 gpurun_out/ncu_h.log | cut -c1-200
