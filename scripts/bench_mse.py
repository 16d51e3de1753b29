"""Dev aid (GPU): throughput of the one-pass MSE observer (20 grid points per chunk)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack
dev = torch.device("cuda", 0)
for name in ("W4A16_ASYM", "INT4_G32_SYM", "FP8_CHANNEL"):
    w = synth_stack([0, 1], 9728, 2560, 0, dev)
    for _ in range(2):
        ops.observe_mse_minmax(w, PRESETS[name])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.observe_mse_minmax(w, PRESETS[name])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name:13s}: {ms:7.3f} ms for {w.numel()*2/1e6:.0f} MB = {w.numel()*2/ms/1e6:6.0f} GB/s bf16-in, {w.numel()*20/ms/1e6:.0f} G fake-quant+pow evaluations / s", flush=True)
