import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import llmc_restated as R, oracle as O
from quantizers_b200 import awq
from tests.test_gpu_compress import Args
g = torch.Generator().manual_seed(2)
T,K,N=1024,512,256
x = (torch.randn(T, K, generator=g) * (1 + 3 * torch.rand(K, generator=g))).to(torch.bfloat16)
w = (torch.randn(N, K, generator=g) * 0.02).to(torch.bfloat16); w[:, 7] *= 20
xm,_ = R.accumulate_abs_mean([x]); geom=O.Geom(O.GROUP,128)
wm = R.compute_layer_means([w], 128); s = R.awq_scales(xm, wm, 0.35, True)
ref = R.scaled_fake_quantize(w, s, geom, O.INT, 4, False)
got = awq.scaled_fake_quantize(w.cuda(), s.cuda(), Args("int4_g128_asym")).cpu()
bad = (got.view(torch.int16) != ref.view(torch.int16))
print("bad", int(bad.sum()))
idx = bad.nonzero()[:10]
ws = (w.float()*s.view(1,-1)).to(torch.bfloat16)
mn,mx = O.minmax(ws, geom); sc,zp = O.calculate_qparams(mn,mx,O.INT,4,False)
for r,c in idx.tolist():
    print(r,c,"w",w[r,c].item(),"ws",ws[r,c].item(),"s",s[c].item(),"scale",sc[r,c//128].item(),"zp",zp[r,c//128].item(),"got",got[r,c].item(),"ref",ref[r,c].item())
# also ones scale
one = torch.ones(K)
ref1 = R.scaled_fake_quantize(w, one, geom, O.INT, 4, False)
got1 = awq.scaled_fake_quantize(w.cuda(), one.cuda(), Args("int4_g128_asym")).cpu()
print("bad with unit scales", int((got1.view(torch.int16) != ref1.view(torch.int16)).sum()))
