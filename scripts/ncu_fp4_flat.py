"""Dev aid (GPU): launch the single-pass NVFP4 kernel (global scales supplied) a few times for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack
dev = torch.device("cuda", 0)
w = synth_stack(list(range(256)), 768, 2048, 0, dev)
gs = ops.weight_global_scales(w)
for _ in range(6):
    ops.compress_weight(w, PRESETS["NVFP4"], global_scale=gs)
torch.cuda.synchronize()
