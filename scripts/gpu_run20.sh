set -x
( time timeout 900 python -m pytest tests -m gpu -q ) 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python scripts/bench_qparams_paths.py > gpurun_out/qp_final.log 2>&1; tail -26 gpurun_out/qp_final.log
python scripts/bench_observers.py > gpurun_out/obs_final.log 2>&1; tail -11 gpurun_out/obs_final.log
( time timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | tail -4
tail -n 3 gpurun_out/bench_final.err; cut -c1-600 gpurun_out/bench_final.json
