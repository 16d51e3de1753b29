"""Dev aid (GPU): show where the CUDA mse observer and the restated oracle pick different grid points."""
import os, sys
os.environ.setdefault("TORCHDYNAMO_DISABLE", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import llmc_restated as R, oracle as O
from tests.util import synth_weight
from tests.test_gpu_compress import Args
from quantizers_b200 import ops

w = synth_weight(96, 640, torch.bfloat16, 21)
geom = O.Geom(O.GROUP, 128)
mn, mx = O.minmax(w, geom)
E = []
for i in range(20):
    p = 1 - i / 100
    s, z = O.calculate_qparams(p * mn, p * mx, O.INT, 4, False)
    q = O.fake_quantize(w, s, z, geom, O.INT, 4)
    q -= w; q.abs_(); q.pow_(2.4)
    E.append(torch.sum(q.reshape(96, -1, 128), dim=-1))
E = torch.stack(E)
rmn, rmx = R.mse_minmax(w, geom, O.INT, 4, False)
gmn, gmx = ops.observe_mse_minmax(w.cuda(), Args("int4_g128_asym"))
gmn, gmx = gmn.cpu().reshape(rmn.shape), gmx.cpu().reshape(rmx.shape)
bad = ((gmn != rmn) | (gmx != rmx)).nonzero()
print("mismatches", len(bad), "of", rmn.numel())
for r, c in bad[:6].tolist():
    ratios_g = (gmn[r, c].float() / mn[r, c].float()).item(), (gmx[r, c].float() / mx[r, c].float()).item()
    ratios_r = (rmn[r, c].float() / mn[r, c].float()).item(), (rmx[r, c].float() / mx[r, c].float()).item()
    print((r, c), "gpu p~", ratios_g, "ref p~", ratios_r, "raw", mn[r, c].item(), mx[r, c].item())
    print("   ref errs", [f"{v:.4e}" for v in E[:, r, c].float().tolist()])
