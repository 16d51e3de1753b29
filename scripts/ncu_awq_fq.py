"""Dev aid (GPU): the AWQ per-ratio weight update (awq_fq_fast.cu) on the Qwen3-4B gate/up shape, 20 ratios, for ncu capture / timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq
from quantizers_b200.scheduler import PRESETS

dev = torch.device("cuda", 0)
N, K, R = 9728, 2560, 20
g = torch.Generator(device=dev).manual_seed(1)
w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
scales = torch.exp(torch.randn(R, K, generator=g, device=dev) * 0.5)
out = torch.empty(R, N, K, dtype=torch.bfloat16, device=dev)
for name in ("W4A16_ASYM", "INT4_G32_SYM"):
    for _ in range(3):
        awq.scaled_fake_quantize_grid(w, scales, PRESETS[name], out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        awq.scaled_fake_quantize_grid(w, scales, PRESETS[name], out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{name}: {ms*1e3:.1f} us for {N}x{K} x {R} ratios = {N*K*R/ms/1e9:.2f} T element-ratios/s = {N*K*R/ms/1e9/148*1e3:.1f} G/s per SM; "
          f"traffic (2 + 2R) B/elem -> {N*K*(2+2*R)/ms/1e6:.0f} GB/s", flush=True)
