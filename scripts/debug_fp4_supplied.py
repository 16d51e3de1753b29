"""GPU debug: first mismatches of the supplied-scale NVFP4 quantize_pack against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from quantizers_b200 import ops
from tests.test_gpu_compress import Args
from tests.util import geom_of, synth_weight

rows, cols = (int(v) for v in (sys.argv[1:3] or (3, 16)))
geom, args = geom_of("nvfp4"), Args("nvfp4")
gen = torch.Generator().manual_seed(rows + cols)
ws = [synth_weight(rows, cols, torch.bfloat16, 70 + i) for i in range(2)]
gs = O.generate_gparam(min(float(w.float().min()) for w in ws), max(float(w.float().max()) for w in ws), torch.bfloat16)
for i, w in enumerate(ws):
    mn, mx = O.minmax(w, geom)
    s, _ = O.calculate_qparams(mn, mx, O.FP4, 4, True, gs)
    codes = s.to(torch.float8_e4m3fn).view(torch.uint8).to(torch.int16)
    codes = (codes + torch.randint(-2, 3, codes.shape, generator=gen, dtype=torch.int16)).clamp(1, 0x7e).to(torch.uint8)
    s = codes.view(torch.float8_e4m3fn).to(torch.bfloat16)
    s.view(-1)[0] = 0.0
    if s.numel() > 2:
        s.view(-1)[1] = 0.3
    q_o = O.quantize(w, s, torch.zeros(s.shape, dtype=torch.float8_e4m3fn), geom, O.FP4, 4, gs)
    want = q_o[:, 0::2] | (q_o[:, 1::2] << 4)
    got = ops.quantize_pack(w.cuda(), s.cuda(), None, args, global_scale=gs.cuda()).cpu()
    bad = (got != want).nonzero()
    print(f"matrix {i}: {len(bad)} bytes differ; gs={float(gs)}")
    for r, c in bad[:6].tolist():
        g = (2 * c) // 16
        print(f"  row {r} byte {c}: got {int(got[r, c]):#04x} want {int(want[r, c]):#04x}  x=({float(w[r, 2*c])!r}, {float(w[r, 2*c+1])!r}) scale={float(s[r, g])!r} code={int(s[r, g].to(torch.float8_e4m3fn).view(torch.uint8)):#04x}")
