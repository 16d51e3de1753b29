# round 2: ncu --set full of the lean fused NVFP4 kernel (same script / shape as the round-1/2 captures) + the A/B with the new defaults
mkdir -p gpurun_out/r2
python scripts/ncu_kernels.py NVFP4 2>&1 | tail -1 | tee gpurun_out/r2/nvfp4_v2_plain.log
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -f -k regex:nvfp4_fused --launch-skip 3 -c 1 -o gpurun_out/r2/prof_nvfp4_v2 python scripts/ncu_kernels.py NVFP4 > gpurun_out/r2/prof_nvfp4_v2.log 2>&1
ncu -i gpurun_out/r2/prof_nvfp4_v2.ncu-rep --page raw --csv > gpurun_out/r2/prof_nvfp4_v2_raw.csv 2>/dev/null
ncu -i gpurun_out/r2/prof_nvfp4_v2.ncu-rep --page source --csv > gpurun_out/r2/prof_nvfp4_v2_source.csv 2>/dev/null
rm -f gpurun_out/r2/prof_nvfp4_v2.ncu-rep
for v in "" "B200Q_FP4_NT=4" ; do
  env $v B200Q_AB_TAG="${v:-default}" timeout 300 python scripts/ab_fp4_v2.py 2>&1 | tail -1
done | tee gpurun_out/r2/ab_fp4_v2b.log
du -sh gpurun_out
