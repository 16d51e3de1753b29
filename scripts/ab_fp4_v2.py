"""GPU A/B of the fused NVFP4 kernel (B200Q_FP4_FUSED_V1 / B200Q_FP4_NT / B200Q_FP4_FMA are read once per process -> one process per
variant).  Shapes: MiniMax-M2 expert stacks (gate/up share a global scale: span 2), a dense layer, a ragged one.  Caller-owned
outputs, 5 warm + 20 timed launches; prints a checksum per shape so the variants can be compared bit for bit."""
import hashlib, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
tag = os.environ.get("B200Q_AB_TAG", "default")
a = PRESETS["NVFP4"]
res = {"tag": tag}
for name, (E, R, C, span) in (("moe_gate_up", (512, 1536, 3072, 2)), ("moe_down", (256, 3072, 1536, 1)), ("dense", (16, 9728, 2560, 1)),
                              ("qkv", (12, 4096, 4096, 3)), ("ragged", (6, 1000, 2064, 3))):
    w = synth_stack(list(range(E)), R, C, 0, dev)
    out = ops.compress_outputs(w.shape, a, w.dtype, dev, fuse_span=span)
    for _ in range(5):
        ops.compress_weight(w, a, fuse_span=span, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.compress_weight(w, a, fuse_span=span, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    h = hashlib.sha1()
    for k in sorted(out):
        if not k.startswith("_") and torch.is_tensor(out[k]) and out[k].is_cuda:
            h.update(out[k].contiguous().view(torch.uint8).cpu().numpy().tobytes())
    res[name] = {"GB/s": round(a.bytes_per_element() * w.numel() / ms / 1e6), "sha": h.hexdigest()[:12]}
    del out, w
print(json.dumps(res), flush=True)
