"""Dev aid (GPU): per-launch timing jitter of the fused AWQ loss GEMM (same call repeated), with clocks / power sampled alongside."""
import os, sys, subprocess, time, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq

dev = torch.device("cuda", 0)
T, K, N, R = 32768, 1536, 3072, 20
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(T, K, generator=g, device=dev).to(torch.bfloat16)
w = (torch.randn(R + 1, N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown",
                        "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
for _ in range(3):
    awq.gemm_loss_pairs(x, None, w[0], w[1:])
torch.cuda.synchronize()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
evs = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
evs[0].record()
for i in range(n):
    awq.gemm_loss_pairs(x, None, w[0], w[1:])
    evs[i + 1].record()
torch.cuda.synchronize()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(n)]
smi.terminate()
out = smi.stdout.read().strip().splitlines()
fl = (R + 1) * 2.0 * T * K * N
print("per-launch ms:", " ".join(f"{v:.1f}" for v in ms))
print(f"median {statistics.median(ms):.2f} ms = {fl / statistics.median(ms) / 1e9:.0f} TF/s; max {max(ms):.1f}; min {min(ms):.2f}")
rows = [[c.strip() for c in l.split(",")] for l in out if l.count(",") >= 5]
if rows:
    clk = [float(r[0]) for r in rows]; pw = [float(r[1]) for r in rows]
    print(f"smi samples {len(rows)}: clocks min/median/max {min(clk):.0f}/{statistics.median(clk):.0f}/{max(clk):.0f} MHz, power max {max(pw):.0f} W, temp max {max(float(r[2]) for r in rows):.0f} C")
    print("reasons active:", {k: sum(1 for r in rows if r[3 + i].lower().startswith("active")) for i, k in enumerate(("sw_power_cap", "hw_slowdown", "sw_thermal"))})
    print("clock trace:", " ".join(f"{c:.0f}" for c in clk[::max(len(clk) // 40, 1)]))
