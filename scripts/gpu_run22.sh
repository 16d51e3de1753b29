set -x
python scripts/ncu_new_kernels.py 2>&1 | tail -2
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'decode_int4_batch|elementwise_fast|minmax_group_bf16|minmax_block128|nvfp4_supplied' --launch-skip 6 -c 6 -f -o gpurun_out/prof_r1_new python scripts/ncu_new_kernels.py > gpurun_out/ncu_new.log 2>&1
tail -3 gpurun_out/ncu_new.log; ls -la gpurun_out/prof_r1_new.ncu-rep
