"""Dev aid (GPU): NVFP4 compress with caller-supplied global scales (single pass, nvfp4_flat_kernel) vs the fused single-launch
kernel that also computes them -- how much of the fused time is the compress pass itself?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
a = PRESETS["NVFP4"]
for (E, R, C) in [(128, 768, 2048), (512, 768, 2048)]:
    w = synth_stack(list(range(E)), R, C, 0, dev)
    gs = ops.weight_global_scales(w)
    for name, fn in (("flat (gs given)", lambda: ops.compress_weight(w, a, global_scale=gs)), ("fused", lambda: ops.compress_weight(w, a)),
                     ("global scales only", lambda: ops.weight_global_scales(w))):
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
        print(f"[{E},{R},{C}] {name:20s}: {ms*1e3:7.1f} us, {w.numel()*2/ms/1e6:5.0f} GB/s bf16-in", flush=True)
