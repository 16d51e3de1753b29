#!/bin/bash
# round 2 profile pass: ncu --set full of the dominant kernels (free-running and base clocks) -> raw CSV pages (the .ncu-rep files are
# deleted on the box: gpurun brings back at most 64 MiB), unfiltered launch lists
set -x
mkdir -p gpurun_out/r2
cap() {  # name, clock-control, kernel regex, skip, command...
  local name=$1 clk=$2 rx=$3 skip=$4; shift 4
  timeout 300 ncu --set full --clock-control $clk --kernel-name-base demangled -f -k regex:$rx --launch-skip $skip -c 1 -o gpurun_out/r2/$name "$@" > gpurun_out/r2/$name.log 2>&1
  ncu -i gpurun_out/r2/$name.ncu-rep --page raw --csv > gpurun_out/r2/${name}_raw.csv 2>/dev/null
  rm -f gpurun_out/r2/$name.ncu-rep
}
cap prof_int4_asym none group_tma 3 python scripts/ncu_kernels.py W4A16_ASYM
cap prof_int4_asym_baseclk base group_tma 3 python scripts/ncu_kernels.py W4A16_ASYM
cap prof_int4_g32 none group_tma 3 python scripts/ncu_kernels.py INT4_G32_SYM
cap prof_fp8_block none block_fp8 3 python scripts/ncu_kernels.py FP8_BLOCK
cap prof_nvfp4 none nvfp4_fused 3 python scripts/ncu_kernels.py NVFP4
cap prof_awq_fq none awq_fq_grid 3 python scripts/ncu_awq_fq.py
cap prof_attn none attn_core 3 python scripts/bench_attn.py
cap prof_awq_gemm none awq_gemm_loss 1 python scripts/ncu_awq_gemm.py 32768
cap prof_rope none qk_norm_rope 3 python scripts/ncu_awq_layer.py
# launch lists (unfiltered)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2/awq_layer_launches2.csv python scripts/ncu_awq_layer.py > gpurun_out/r2/awq_layer_ncu2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r2/launches_bench.csv python bench.py --gpus 1 --steps 2 --warmup 3 --awq-layers 1 --moe-layers 2 --moe-steps 2 --moe-awq-experts 2 --moe-block-experts 8 --glm-units 8 --glm-file-gb 0 --no-cpu-baseline --e2e-steps 1 --no-parity > gpurun_out/r2/launches_bench.log 2>&1
python scripts/ncu_awq_fq.py > gpurun_out/r2/awq_fq_plain.log 2>&1; cat gpurun_out/r2/awq_fq_plain.log
du -sh gpurun_out
