"""GPU: q/k RMSNorm + RoPE kernel (b200q_qk_norm_rope) on the Qwen3-4B attention geometry, T = 32 768 tokens."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import _lib as L, awq

H, HKV, D, S, T = 32, 8, 128, 512, 32768
dev = torch.device("cuda", 0)
qkv = torch.randn(T, (H + 2 * HKV) * D, device=dev).to(torch.bfloat16)
qn = torch.ones(D, device=dev, dtype=torch.bfloat16)
par = awq.AttentionParent(torch.zeros(8, H * D, dtype=torch.bfloat16, device=dev), H, HKV, D, S, qn, qn)
fn = lambda: L.check(L.lib().b200q_qk_norm_rope(L.ptr(qkv), T, H, HKV, D, S, L.ptr(qn), L.ptr(qn), L.ptr(par.cos), L.ptr(par.sin), ctypes.c_float(1e-6),
                                                 L.stream_ptr(dev)))
for _ in range(3):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
nbytes = 2 * T * (H + HKV) * D * 2
print(f"ROPE_BATCH={os.environ.get('B200Q_ROPE_BATCH', '4')}: {ms*1e3:.1f} us, {nbytes/ms/1e6:.0f} GB/s (read + write of the q/k columns), {nbytes/ms/1e6/6549.4:.3f} of the HBM roofline")
