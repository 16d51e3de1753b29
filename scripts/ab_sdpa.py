"""Dev aid (GPU): which torch SDPA backend is fastest for the AWQ attention parent's core (B=64, H=32, Hkv=8, S=512, d=128, strided q/k/v)."""
import torch, time
from torch.nn.attention import SDPBackend, sdpa_kernel
dev = torch.device("cuda", 0)
B, S, H, HKV, D = 64, 512, 32, 8, 128
qkv = torch.randn(B * S, (H + 2 * HKV) * D, device=dev, dtype=torch.bfloat16)
q, k, v = qkv.split([H * D, HKV * D, HKV * D], dim=-1)
q = q.unflatten(-1, (H, D)).unflatten(0, (B, S)).transpose(1, 2)
k = k.unflatten(-1, (HKV, D)).unflatten(0, (B, S)).transpose(1, 2)
v = v.unflatten(-1, (HKV, D)).unflatten(0, (B, S)).transpose(1, 2)
ref = None
for name, be in (("default", None), ("flash", SDPBackend.FLASH_ATTENTION), ("cudnn", SDPBackend.CUDNN_ATTENTION), ("efficient", SDPBackend.EFFICIENT_ATTENTION)):
    try:
        def run():
            if be is None:
                return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
            with sdpa_kernel(be):
                return torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=True, enable_gqa=True)
        for _ in range(3):
            o = run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(21):
            o = run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if ref is None:
            ref = o.float()
        err = (o.float() - ref).abs().max().item()
        print(f"{name:10s}: {ms:7.2f} ms for 21 calls ({4*S*S*D*H*B*21/ms/1e9:.0f} TF/s dense-equivalent), max |diff| vs default {err:.3e}, out stride {o.stride()}", flush=True)
    except Exception as e:  # noqa: BLE001
        print(f"{name:10s}: failed: {str(e)[:150]}", flush=True)
