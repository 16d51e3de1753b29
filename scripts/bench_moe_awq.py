"""Dev aid (GPU): per-expert w3 -> w2 AWQ mapping search (MiniMax-M2.1 shapes), a few experts, CUDA-event timed."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq, scheduler as S

dev = torch.device("cuda", 0)
E = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
qargs = S.PRESETS["INT4_G32_SYM"]
w1, w3, w2, xs = S.synth_moe_awq_experts(0, range(E), T, dev)
def run(w2, w3):
    if os.environ.get("SYNC_PER_EXPERT"):
        res = []
        for e in range(E):
            r = awq.compute_best_scale(xs[e], [w2[e]], awq.linear_parent, qargs)
            awq.smooth([w2[e]], w3[e], r[0])
            res.append(r)
        return res
    return awq.search_expert_mappings(xs, w2, qargs, smooth_weight=w3)


run(w2.clone(), w3.clone())  # full warm-up pass: every kernel loaded, workspaces grown
torch.cuda.synchronize()
for rep in range(2):
    a, b = w2.clone(), w3.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    res = run(a, b)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{E} experts: {ms / E:.2f} ms/expert (wall {(time.perf_counter() - t0) * 1e3 / E:.2f}), "
          f"{awq.expert_mapping_flops(T, 1536, 3072) * E / ms / 1e9:.0f} TFLOP/s; ratios {[r[1] for r in res]}")
