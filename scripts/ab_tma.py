"""Dev aid (GPU): A/B the ALU-pipe relief variants of the TMA group kernel (B200Q_TMA_LEGACY_ALU=1 selects the old code) --
run once per setting (the switch is read once per process)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
tag = "legacy" if os.environ.get("B200Q_TMA_LEGACY_ALU") else "fma"
for n in ["W4A16_ASYM", "W4A16", "INT4_G32_SYM", "FP8_G32"]:
    w = synth_stack(list(range(16)), 9728, 2560, 0, dev)
    for _ in range(3):
        ops.compress_weight(w, PRESETS[n])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.compress_weight(w, PRESETS[n])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    alg = PRESETS[n].bytes_per_element() * w.numel() / ms / 1e6
    print(f"{tag:6s} {n:13s}: {ms*1e3:7.1f} us, {alg:5.0f} GB/s algorithmic ({alg/6549.4:.3f} of HBM peak)", flush=True)
