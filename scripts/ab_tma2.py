"""GPU A/B of the INT4 group kernel variants (B200Q_TMA_VAR / B200Q_TMA_BRACKET, read once per process -> one process per variant).
Shapes: the two INT4 launches of the headline step + the AWQ recipes' g32 sym.  Caller-owned outputs, 5 warm + 20 timed launches."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
tag = os.environ.get("B200Q_AB_TAG", "default")
res = {"tag": tag}
for shape_name, (E, R, C) in (("gate_up", (72, 9728, 2560)), ("down", (36, 2560, 9728))):
    w = synth_stack(list(range(E)), R, C, 0, dev)
    for preset in ("W4A16_ASYM", "W4A16", "INT4_G32_SYM"):
        a = PRESETS[preset]
        out = ops.compress_outputs(w.shape, a, w.dtype, dev)
        for _ in range(5):
            ops.compress_weight(w, a, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.compress_weight(w, a, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        alg = a.bytes_per_element() * w.numel() / ms / 1e6
        res[f"{shape_name}:{preset}"] = round(alg)
        del out
    del w
print(json.dumps(res), flush=True)
