"""GPU: throughput of the stand-alone observer / statistics kernels against the HBM roofline.

The fused compress kernels carry their own observer; these are the entry points the calibration loop calls on their own:
  b200q_minmax        weight min/max per channel / group / 128x128 block        (memoryless_minmax, O1)
  b200q_global_scale  per-tensor min/max -> NVFP4 global scale                  (get_global_scale / static_minmax, O2/O3)
  b200q_abs_sum_cols  sum_t |x[t, k]|                                           (_accumulate_mean, O5)
  b200q_wmean_accumulate  sum_rows |w| / (group absmax + 1e-6)                  (_compute_layer_means, O5)
Algorithmic bytes = the input read once (2 B / element); outputs are negligible.  Writes gpurun_out/observers.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq, ops
from quantizers_b200.scheduler import PRESETS, SchemeArgs, synth_stack

dev = torch.device("cuda", 0)
peak = 6549.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass

w = synth_stack(list(range(32)), 9728, 2560, 0, dev)                       # 1.59 GB of weights
torch.manual_seed(4321)
x = (torch.randn(32768, 9728, device=dev) * (1 + 3 * torch.rand(9728, device=dev))).to(torch.bfloat16)  # 637 MB of activations, T = 64 x 512
wd = w[:8].reshape(-1, 2560).contiguous()                                    # down_proj-like stack for w_mean: rows x K

cases = [
    ("minmax channel (weights)", w.numel() * 2, lambda: ops.observe_minmax(w, PRESETS["FP8_CHANNEL"])),
    ("minmax group 128 (weights)", w.numel() * 2, lambda: ops.observe_minmax(w, PRESETS["W4A16_ASYM"])),
    ("minmax group 32 (weights)", w.numel() * 2, lambda: ops.observe_minmax(w, PRESETS["INT4_G32_SYM"])),
    ("minmax group 16 (weights)", w.numel() * 2, lambda: ops.observe_minmax(w, SchemeArgs(4, "float", True, "group", 16))),
    ("minmax block 128x128 (weights)", w.numel() * 2, lambda: ops.observe_minmax(w, PRESETS["FP8_BLOCK"])),
    ("global scale (activations [32768, 9728])", x.numel() * 2, lambda: ops.observe_global_scale(x)),
    ("global scale (weight stack as one tensor)", w.numel() * 2, lambda: ops.observe_global_scale(w)),
    ("per-matrix global scales (32 matrices)", w.numel() * 2, lambda: ops.weight_global_scales(w)),
    ("abs_sum_cols (activations [32768, 9728])", x.numel() * 2, lambda: awq.abs_sum_cols(x)),
    ("abs_sum_cols (activations [32768, 2560])", 32768 * 2560 * 2, None),
    ("w_mean accumulate g128 ([77824, 2560])", wd.numel() * 2, lambda: awq.compute_layer_means([wd], 128)),
]
x_small = x[:, :2560].contiguous()
cases[9] = (cases[9][0], cases[9][1], lambda: awq.abs_sum_cols(x_small))

rows = []
for name, nbytes, fn in cases:
    try:
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print(name, "ERROR", str(e)[:200], flush=True)
        continue
    ms = e0.elapsed_time(e1) / 10
    gbps = nbytes / ms / 1e6
    rows.append({"kernel": name, "ms": ms, "bytes": nbytes, "GBps": gbps, "frac_of_hbm_peak": gbps / peak})
    print(f"{name:46s} {ms*1e3:8.1f} us  {gbps:6.0f} GB/s  {gbps/peak:.3f}", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
tag = os.environ.get("B200Q_BENCH_TAG", "")
json.dump({"peak_GBps": peak, "rows": rows}, open(f"gpurun_out/observers{'_' + tag if tag else ''}.json", "w"), indent=1)
