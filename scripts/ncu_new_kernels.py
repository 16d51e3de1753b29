"""Dev aid (GPU): two launches each of the streaming kernels added at the end of round 1, on L2-exceeding inputs, for one ncu capture:
decode_int4_batch_kernel (asym g128), elementwise_fast_kernel (FP8 128x128 quantize, INT4 g128 asym fake_quantize), minmax_group_bf16_kernel
(g128), minmax_block128_bf16_kernel, nvfp4_supplied_kernel (quantize_pack)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
w3 = synth_stack(list(range(8)), 9728, 2560, 0, dev)   # 398 MB
w = w3.reshape(-1, 2560)
for _ in range(2):
    a = PRESETS["W4A16_ASYM"]
    sd = ops.compress_weight(w3, a)
    ops.decompress_int_packed(sd["weight_packed"], sd["weight_scale"], sd["weight_zero_point"], w3.shape, a)
    mn, mx = ops.observe_minmax(w, a)
    s, z = ops.calculate_qparams(mn, mx, a)
    ops.fake_quantize(w, s, z, a)
    b = PRESETS["FP8_BLOCK"]
    mn, mx = ops.observe_minmax(w, b)
    s, z = ops.calculate_qparams(mn, mx, b)
    ops.quantize(w, s, None, b, dtype=torch.float8_e4m3fn)
    n = PRESETS["NVFP4"]
    sd = ops.compress_weight(w, n)
    ops.quantize_pack(w, sd["weight_scale"], None, n, global_scale=sd["weight_global_scale"])
torch.cuda.synchronize()
print("ok")
