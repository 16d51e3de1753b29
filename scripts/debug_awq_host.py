"""Dev aid (GPU): host-side time of every step of one expert-mapping search (where does the CPU block?)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq, scheduler as S
from quantizers_b200 import _lib as L

dev = torch.device("cuda", 0)
E, T = 6, 32768
qargs = S.PRESETS["INT4_G32_SYM"]
w1, w3, w2, xs = S.synth_moe_awq_experts(0, range(E), T, dev)
awq.search_expert_mappings(xs, w2.clone(), qargs, smooth_weight=w3.clone())
torch.cuda.synchronize()
n_grid = 20
tot = {}
def tick(name, t0):
    dt = (time.perf_counter() - t0) * 1e3
    tot.setdefault(name, []).append(dt)
    return time.perf_counter()
t_all = time.perf_counter()
for e in range(E):
    x, w = xs[e], w2[e]
    t = time.perf_counter()
    xsum = awq.abs_sum_cols(x); t = tick("abs_sum_cols", t)
    x_mean = xsum / float(x.shape[0]); t = tick("div", t)
    w_mean = awq.compute_layer_means([w], qargs.group_size); t = tick("w_mean", t)
    scales = awq.awq_scales(x_mean, w_mean, [i / n_grid for i in range(n_grid)], True); t = tick("awq_scales", t)
    acc = torch.zeros(n_grid + 1, dtype=torch.float32, device=dev); t = tick("zeros", t)
    w_all = awq.workspace.get("w_all", (n_grid + 1, w.shape[0], w.shape[1]), w.dtype, dev); t = tick("ws.get", t)
    w_all[0].copy_(w); t = tick("copy ref", t)
    awq.scaled_fake_quantize_grid(w, scales, qargs, w_all[1:]); t = tick("fq_grid", t)
    sums = awq.gemm_loss_pairs(x.contiguous(), None, w_all[0], w_all[1:]); t = tick("gemm_loss", t)
    acc[:n_grid] = sums; t = tick("acc slice", t)
    acc[n_grid] = float(T * w.shape[0]); t = tick("acc scalar", t)
    best = awq._first_min_device(acc); t = tick("first_min", t)
    bs = scales.index_select(0, best.clamp(min=0)).reshape(-1); t = tick("index_select", t)
    awq.smooth([w], w3[e], bs); t = tick("smooth", t)
host = (time.perf_counter() - t_all) * 1e3
torch.cuda.synchronize()
print(f"host enqueue total {host:.1f} ms, with sync {(time.perf_counter() - t_all) * 1e3:.1f} ms for {E} experts")
for k, v in tot.items():
    print(f"  {k:14s} " + " ".join(f"{d:7.2f}" for d in v))
