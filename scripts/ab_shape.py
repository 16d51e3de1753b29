"""Dev aid (GPU): does the row length matter?  Same bytes, different [n, rows, cols] views, per-launch CUDA-event timing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS

dev = torch.device("cuda", 0)
torch.manual_seed(0)
base = (torch.randn(72 * 9728 * 2560 // 4, device=dev, dtype=torch.float32) * 0.02).to(torch.bfloat16).repeat(4)
shapes = {"W4A16_ASYM": [(72, 9728, 2560), (72, 2560, 9728), (36, 2560, 9728), (36, 9728, 5120), (18, 9728, 10240), (72, 4864, 5120)],
          "FP8_BLOCK": [(36, 4096, 2560), (36, 2560, 4096), (72, 1024, 2560), (18, 4096, 5120)]}
for n, lst in shapes.items():
    for rep in range(2):
        for shp in lst:
            numel = shp[0] * shp[1] * shp[2]
            w = base[:numel].view(shp)
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
            alg = PRESETS[n].bytes_per_element() * numel / ms / 1e6
            print(f"{n:11s} {str(shp):20s}: {ms*1e3:7.1f} us, {alg:5.0f} GB/s algorithmic ({alg/6549.4:.3f})", flush=True)
