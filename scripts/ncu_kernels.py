"""Dev aid (GPU): launch each headline fused-compress kernel a few times on L2-exceeding inputs, for ncu capture."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
names = sys.argv[1:] or ["W4A16_ASYM", "INT4_G32_SYM", "FP8_BLOCK", "NVFP4", "FP8_G32"]
shapes = {"NVFP4": (128, 768, 2048)}
for n in names:
    E, R, C = shapes.get(n, (8, 9728, 2560))
    w = synth_stack(list(range(E)), R, C, 0, dev)
    torch.cuda.synchronize()
    for _ in range(3):
        ops.compress_weight(w, PRESETS[n])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.compress_weight(w, PRESETS[n])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"{n}: {w.numel()*2/ms/1e6:.0f} GB/s bf16-in, {PRESETS[n].bytes_per_element()*w.numel()/ms/1e6:.0f} GB/s algorithmic, {ms*1e3:.1f} us")
