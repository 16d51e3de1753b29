"""GPU: fused observe -> qparams -> quantize -> pack throughput of every scheme the reference's recipes use (plus the other CT
strategies the generic kernels cover), one 1.6 GB bf16 stack per launch, CUDA events, 3 cold + 20 timed launches.  Writes
gpurun_out/schemes.json; the committed copy is profiles/r1_schemes.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from quantizers_b200 import ops
from quantizers_b200.scheduler import PRESETS, SchemeArgs, synth_stack

dev = torch.device("cuda", 0)
peak = 6549.4
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
schemes = dict(PRESETS)
schemes["INT4_G32_ASYM"] = SchemeArgs(4, "int", False, "group", 32)
schemes["INT4_CHANNEL_SYM"] = SchemeArgs(4, "int", True, "channel")
schemes["INT8_CHANNEL_SYM"] = SchemeArgs(8, "int", True, "channel")
schemes["FP8_G128"] = SchemeArgs(8, "float", True, "group", 128)
schemes["FP8_TENSOR"] = SchemeArgs(8, "float", True, "tensor")
w = synth_stack(list(range(32)), 9728, 2560, 0, dev)   # 32 x [9728, 2560] bf16 = 1.59 GB
rows = []
for name, a in schemes.items():
    kw = {"fuse_span": 2} if name == "NVFP4" else {}
    try:
        for _ in range(3):
            out = ops.compress_weight(w, a, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            out = ops.compress_weight(w, a, **kw)
        e1.record()
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        rows.append({"scheme": name, "error": str(e)[:200]})
        print(name, "ERROR", str(e)[:200], flush=True)
        continue
    ms = e0.elapsed_time(e1) / 20
    out_bytes = sum(v.numel() * v.element_size() for k, v in out.items() if k != "weight_shape")
    alg = (w.numel() * 2 + out_bytes) / ms / 1e6
    rows.append({"scheme": name, "ms": ms, "bf16_in_GBps": w.numel() * 2 / ms / 1e6, "algorithmic_GBps": alg, "frac_of_hbm_peak": alg / peak,
                 "alg_bytes_per_element": (w.numel() * 2 + out_bytes) / w.numel()})
    print(f"{name:17s} {ms*1e3:8.1f} us  {w.numel()*2/ms/1e6:6.0f} GB/s bf16-in  {alg:6.0f} GB/s algorithmic  {alg/peak:.3f}", flush=True)
    del out
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"stack": list(w.shape), "peak_hbm_gbs": peak, "rows": rows}, open("gpurun_out/schemes.json", "w"), indent=1)
