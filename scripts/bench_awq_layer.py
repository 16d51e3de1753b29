"""Dev aid (GPU): time the AWQ search of one Qwen3-4B decoder layer (config 1) stage by stage."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq
from quantizers_b200.scheduler import PRESETS, synth_awq_layer

dev = torch.device("cuda", 0)
T = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cfg = dict(n_heads=32, n_kv=8, head_dim=128, seq_len=512)
w, acts = synth_awq_layer(0, T, dev)
args = PRESETS["W4A16_ASYM"]
flops = awq.decoder_layer_flops(T, 2560, 9728, 32, 8, 128, 512)

def timed(f, n=1):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): r = f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, r

for rep in range(2):
    ms, res = timed(lambda: awq.search_decoder_layer({k: v.clone() for k, v in w.items()}, acts, args, **cfg))
    print(f"layer search: {ms:.1f} ms = {flops/ms/1e9:.0f} TFLOP/s, {1e3/ms:.2f} layers/s; ratios {[(k, v[1]) for k, v in res.items()]}", flush=True)
# stage breakdown
ms, _ = timed(lambda: awq.compute_best_scale(acts["down_in"], [w["down"]], awq.linear_parent, args)); print(f"  down mapping     {ms:.1f} ms")
ms, _ = timed(lambda: awq.compute_best_scale(acts["mlp_in"], [w["gate"], w["up"]], awq.MLPParent(w["down"]), args)); print(f"  gate/up mapping  {ms:.1f} ms")
ap = awq.AttentionParent(w["o"], 32, 8, 128, 512, w["q_norm"], w["k_norm"])
ms, _ = timed(lambda: awq.compute_best_scale(acts["attn_in"], [w["q"], w["k"], w["v"]], ap, args)); print(f"  q/k/v mapping    {ms:.1f} ms")
wall = torch.stack([torch.cat([w["gate"], w["up"]])] * 21)
ms, h = timed(lambda: awq.gemm_project(acts["mlp_in"], wall, swiglu=True)); print(f"    swiglu project x21  {ms:.1f} ms = {21*2*2.0*T*2560*9728/ms/1e9:.0f} TF/s")
ms, _ = timed(lambda: awq.gemm_loss_pairs(h[0], h[1:], w["down"], None)); print(f"    down loss (A varies) x21  {ms:.1f} ms = {21*2.0*T*2560*9728/ms/1e9:.0f} TF/s")
del h, wall
wall = torch.stack([torch.cat([w["q"], w["k"], w["v"]])] * 21)
ms, qkv = timed(lambda: awq.gemm_project(acts["attn_in"], wall)); print(f"    qkv project x21  {ms:.1f} ms = {21*2.0*T*2560*6144/ms/1e9:.0f} TF/s")
ms, a = timed(lambda: ap.core(qkv[0])); print(f"    attention core x1  {ms:.2f} ms (x21 = {21*ms:.1f})")
ms, _ = timed(lambda: [awq.scaled_fake_quantize(w["down"], torch.ones(9728, device=dev), args) for _ in range(20)]); print(f"    scaled_fake_quantize down x20  {ms:.2f} ms")
