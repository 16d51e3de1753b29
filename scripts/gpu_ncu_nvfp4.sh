set -x
export B200Q_FP4_SLACK=6
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:nvfp4_resident --launch-skip 3 -c 1 -f -o gpurun_out/prof_nvfp4_persist python scripts/ncu_kernels.py NVFP4 > gpurun_out/ncu_nvfp4_persist.log 2>&1
tail -3 gpurun_out/ncu_nvfp4_persist.log
