for v in "" "B200Q_FP4_FMA=1"; do echo "variant: $v"; env $v python scripts/ncu_kernels.py NVFP4; done
