"""Dev aid (GPU): one launch of the fused AWQ loss GEMM at the Qwen3-4B down_proj mapping shape, for ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq

dev = torch.device("cuda", 0)
T, K, N, R = int(sys.argv[1]) if len(sys.argv) > 1 else 8192, 9728, 2560, 4
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randn(T, K, generator=g, device=dev).to(torch.bfloat16)
w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
wq = (w[None].float() + 0.001 * torch.randn(R, N, K, generator=g, device=dev)).to(torch.bfloat16)
for _ in range(2):
    l = awq.gemm_loss_fused(x, w, wq)
torch.cuda.synchronize()
print(l.tolist())
