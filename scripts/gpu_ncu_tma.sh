ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:group_tma --launch-skip 3 -c 1 -f -o gpurun_out/prof_tma_fma python scripts/ncu_kernels.py W4A16_ASYM > gpurun_out/ncu_tma_fma.log 2>&1
tail -n 2 gpurun_out/ncu_tma_fma.log
