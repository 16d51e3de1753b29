"""Dev aid (GPU): time the fused NVFP4 compress (|max| -> global scale -> codes, one launch): the shared-memory resident kernel
against the L2-resident two-pass kernel (B200Q_FP4_PERSISTENT=0), and check they agree bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, synth_stack

dev = torch.device("cuda", 0)
a = PRESETS["NVFP4"]
DBG = os.environ.get('B200Q_FP4_DBG', '0') != '0'
shapes = [(256, 768, 2048, 2), (128, 2048, 768, 1), (64, 1536, 3072, 2), (6, 1000, 2064, 3), (1024, 768, 2048, 2)]
for (E, R, C, span) in shapes:
    w = synth_stack(list(range(E)), R, C, 0, dev)
    outs = {}
    for mode in ("1c1s4", "1c4s3", "1c4s5", "1c5s3", "1c5s5", "0"):
        os.environ["B200Q_FP4_PERSISTENT"] = mode[0]
        if len(mode) > 1:
            os.environ["B200Q_FP4_CFG"] = mode[2]
            os.environ["B200Q_FP4_SLACK"] = mode[4:]
        for _ in range(3):
            o = ops.compress_weight(w, a, fuse_span=span)
        torch.cuda.synchronize()
        outs[mode] = o
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.compress_weight(w, a, fuse_span=span)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f"[{E},{R},{C}] span {span} resident={mode}: {ms*1e3:.1f} us, {a.bytes_per_element()*w.numel()/ms/1e6:.0f} GB/s algorithmic "
              f"({a.bytes_per_element()*w.numel()/ms/1e6/6549.4:.3f} of HBM peak)", flush=True)
    same = all(all(torch.equal(outs[m][k].view(torch.uint8) if outs[m][k].dtype != torch.float32 else outs[m][k],
                           outs["0"][k].view(torch.uint8) if outs["0"][k].dtype != torch.float32 else outs["0"][k]) for k in outs["0"]) for m in outs if m != "0")
    print("   resident == two-pass:", same if not DBG else "(debug run)", flush=True)
