"""Dev aid (GPU): time the fused tcgen05 AWQ loss GEMM at the Qwen3-4B mapping shapes against torch (cuBLAS) doing
the same work un-fused.  usage: bench_awq_gemm.py [T] [shape ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq

dev = torch.device("cuda", 0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
shapes = {"down": (9728, 2560), "o": (4096, 2560), "expert_w2": (1536, 3072)}
R = 20
for name, (K, N) in shapes.items():
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.randn(T, K, generator=g, device=dev).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
    wq = (w[None].float() + 0.001 * torch.randn(R, N, K, generator=g, device=dev)).to(torch.bfloat16)
    flops = 2.0 * T * K * N * (R + 1)
    for _ in range(2):
        l = awq.gemm_loss_fused(x, w, wq)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        l = awq.gemm_loss_fused(x, w, wq)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    # torch un-fused
    def torch_path():
        ref = x @ w.t()
        acc = torch.zeros(R, device=dev)
        for r in range(R):
            y = x @ wq[r].t()
            acc[r] = (ref - y).float().pow(2).sum()
        return acc
    lt = torch_path()
    torch.cuda.synchronize()
    e0.record()
    lt = torch_path()
    e1.record()
    torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1)
    rel = ((l.double() - lt.double()).abs() / lt.double()).max().item()
    print(f"{name}: T={T} K={K} N={N} R={R}: fused {ms:.2f} ms = {flops/ms/1e9:.0f} TFLOP/s | torch {ms_t:.2f} ms = {flops/ms_t/1e9:.0f} TFLOP/s | max rel diff {rel:.2e}", flush=True)
