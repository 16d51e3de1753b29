"""Dev aid (GPU): NVFP4 fused compress on DENSE shapes (sibling spans of 20-100 MB) next to the MoE expert shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
a = PRESETS["NVFP4"]
for shape, n, span in (((9728, 2560), 32, 2), ((2560, 9728), 32, 1), ((4096, 2560), 36, 1), ((768, 2048), 2048, 2), ((14336, 4096), 8, 2)):
    w = synth_stack(list(range(n)), shape[0], shape[1], 0, dev)
    for _ in range(3):
        ops.compress_weight(w, a, fuse_span=span)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.compress_weight(w, a, fuse_span=span)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    alg = 2.5625 * w.numel() / ms / 1e6
    print(f"{n} x {shape} span {span}: {ms*1e3:7.1f} us, {alg:5.0f} GB/s algorithmic ({alg/6549.4:.3f})", flush=True)
    del w
