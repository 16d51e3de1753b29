"""One AWQ search of a Qwen3-4B decoder layer (config 1 shapes) between cudaProfilerStart/Stop, for an UNFILTERED ncu launch list
(library kernels included): ncu --profile-from-start off --metrics gpu__time_duration.sum ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import awq
from quantizers_b200 import scheduler as S

dev = torch.device("cuda", 0)
T = int(os.environ.get("AWQ_TOKENS", 64 * 512))
w, acts = S.synth_awq_layer(0, T, dev)
cfg = dict(n_heads=32, n_kv=8, head_dim=128, seq_len=512)
q = S.PRESETS["W4A16_ASYM"]
for _ in range(2):
    awq.search_decoder_layer({k: v.clone() for k, v in w.items()}, acts, q, **cfg)
torch.cuda.synchronize()
ww = {k: v.clone() for k, v in w.items()}
torch.cuda.synchronize()
torch.cuda.profiler.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
res = awq.search_decoder_layer(ww, acts, q, **cfg)
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("layer ms", e0.elapsed_time(e1), {k: v[1] for k, v in res.items()})
