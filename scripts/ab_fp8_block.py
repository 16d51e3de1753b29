"""Dev aid (GPU): FP8 128x128 block compress timing on the bench shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack
dev = torch.device("cuda", 0)
a = PRESETS["FP8_BLOCK"]
for (E, R, C) in [(36, 4096, 2560), (72, 1024, 2560), (36, 2560, 4096), (16, 9728, 2560)]:
    w = synth_stack(list(range(E)), R, C, 0, dev)
    for _ in range(3):
        ops.compress_weight(w, a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.compress_weight(w, a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"[{E},{R},{C}]: {ms*1e3:7.1f} us, {3.0*w.numel()/ms/1e6/6549.4:.3f} of HBM peak", flush=True)
