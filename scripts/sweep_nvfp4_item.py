"""Dev aid (GPU): item-kernel knobs (B200Q_FP4_NT tiles per item, B200Q_FP4_LOOKAHEAD_MB) of the fused NVFP4 compress."""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from quantizers_b200 import ops
    from quantizers_b200.scheduler import PRESETS, synth_stack
    dev = torch.device("cuda", 0)
    a = PRESETS["NVFP4"]
    out = []
    for (E, R, C, span) in [(256, 768, 2048, 2), (128, 2048, 768, 1), (64, 1536, 3072, 2), (1024, 768, 2048, 2)]:
        w = synth_stack(list(range(E)), R, C, 0, dev)
        for _ in range(3):
            ops.compress_weight(w, a, fuse_span=span)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.compress_weight(w, a, fuse_span=span)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out.append(f"{a.bytes_per_element()*w.numel()/ms/1e6/6549.4:.3f}")
    print(os.environ.get("B200Q_FP4_NT"), os.environ.get("B200Q_FP4_LOOKAHEAD_MB"), " ".join(out), flush=True)
else:
    for nt in ("4", "8"):
        for la in ("40",):
            env = dict(os.environ, B200Q_FP4_NT=nt, B200Q_FP4_LOOKAHEAD_MB=la)
            subprocess.run([sys.executable, __file__, "child"], env=env)
