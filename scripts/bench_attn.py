"""GPU: the tcgen05 attention core vs torch SDPA on the AWQ attention-parent shape (64 samples x 512 tokens, 32 q / 8 kv heads of 128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import _lib as L

B, S, H, HKV, D = 64, 512, 32, 8, 128
T = B * S
qkv = torch.randn(T, (H + 2 * HKV) * D, device="cuda").to(torch.bfloat16)
lib = L.lib()
out = torch.empty((T, H * D), dtype=torch.bfloat16, device="cuda")
nws = int(lib.b200q_attention_workspace(T, HKV, D, S))
ws = torch.empty(nws, dtype=torch.uint8, device="cuda")
def ours():
    L.check(lib.b200q_attention_core(L.ptr(qkv), T, H, HKV, D, S, L.ptr(out), L.ptr(ws), nws, L.stream_ptr(qkv.device)))
q, k, v = qkv.split([H * D, HKV * D, HKV * D], dim=-1)
q = q.unflatten(-1, (H, D)).unflatten(0, (B, S)).transpose(1, 2)
k = k.unflatten(-1, (HKV, D)).unflatten(0, (B, S)).transpose(1, 2)
v = v.unflatten(-1, (HKV, D)).unflatten(0, (B, S)).transpose(1, 2)
def sdpa():
    return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
for name, fn in (("tcgen05 core", ours), ("torch SDPA", sdpa)):
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
    print(f"{name}: {ms*1e3:.1f} us  ({4.0*S*S*D*H*B/ms/1e9:.0f} dense-equivalent TFLOP/s)", flush=True)
ref = sdpa().transpose(1, 2).reshape(T, H * D)
ours()
print("max |diff| vs SDPA:", float((out.float() - ref.float()).abs().max()))
