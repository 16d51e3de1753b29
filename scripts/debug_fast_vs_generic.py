"""Dev aid (GPU): compare the fast bf16 kernels against the generic templates on the adversarial sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS

allv = torch.arange(0x0001, 0x7F80, dtype=torch.int32).to(torch.int16).view(torch.bfloat16).float()
rows = []
for scale_max in (1.0, 0.0371, 7.5, 3.0e-3, 1.0e30, 1.0e-30, 448.0):
    v = allv[(allv <= scale_max)][-(127 * 64):]
    v = v[: (v.numel() // 127) * 127]
    blk = v.reshape(-1, 127)
    blk = torch.cat([torch.full((blk.shape[0], 1), scale_max), blk * torch.where(torch.arange(127) % 2 == 0, 1.0, -1.0)], dim=1)
    rows.append(blk)
w = torch.cat(rows).to(torch.bfloat16)
w = w[: (w.shape[0] // 8) * 8].cuda()
name = sys.argv[1] if len(sys.argv) > 1 else "NVFP4"
a = ops.compress_weight(w, PRESETS[name])
os.environ["B200Q_DISABLE_FAST"] = "1"
b = ops.compress_weight(w, PRESETS[name])
for k in a:
    x, y = a[k], b[k]
    if x.dtype == torch.float8_e4m3fn:
        x, y = x.view(torch.uint8), y.view(torch.uint8)
    bad = (x != y)
    print(k, "mismatch", int(bad.sum()), "/", x.numel())
    if bad.any() and x.ndim == 2:
        idx = bad.nonzero()[:12]
        for r, c in idx.tolist():
            print("  row", r, "col", c, "fast", int(x[r, c]), "generic", int(y[r, c]), "w", w[r, 2 * c:2 * c + 2].float().tolist() if k == "weight_packed" else "")
        print("  rows with mismatches:", bad.any(dim=1).nonzero().flatten()[:40].tolist())
