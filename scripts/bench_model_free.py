"""Dev aid (GPU): end-to-end throughput of model_free_ptq on a synthetic sharded checkpoint in /tmp (page cache, so this is the
host <-> device pipeline, not the disk): file -> pinned staging -> H2D -> fused FP8 128x128 kernel -> D2H -> pwrite."""
import json, os, shutil, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from safetensors.torch import save_file
from quantizers_b200.model_free import model_free_ptq

src, dst = "/tmp/mfp_src", "/tmp/mfp_dst"
shutil.rmtree(src, ignore_errors=True); shutil.rmtree(dst, ignore_errors=True)
os.makedirs(src)
n_shards, per_shard = int(sys.argv[1]) if len(sys.argv) > 1 else 4, 12
g = torch.Generator().manual_seed(0)
total = 0
for s in range(n_shards):
    t = {}
    for i in range(per_shard):
        l = s * per_shard + i
        t[f"model.layers.{l}.mlp.experts.0.gate_proj.weight"] = (torch.randn(1536, 2048, generator=g) * 0.02).to(torch.bfloat16)
        t[f"model.layers.{l}.mlp.experts.0.down_proj.weight"] = (torch.randn(2048, 1536, generator=g) * 0.02).to(torch.bfloat16)
        t[f"model.layers.{l}.mlp.shared.up_proj.weight"] = (torch.randn(10240, 2048, generator=g) * 0.02).to(torch.bfloat16)
        t[f"model.layers.{l}.input_layernorm.weight"] = torch.ones(2048, dtype=torch.bfloat16)
    total += sum(v.numel() * 2 for v in t.values())
    save_file(t, os.path.join(src, f"model-{s + 1:05d}-of-{n_shards:05d}.safetensors"), metadata={"format": "pt"})
model_free_ptq(src, dst, scheme="FP8_BLOCK", ignore=["lm_head"], max_workers=1)   # warm-up: CUDA context, library load
for workers in (1, 2, 4, 16):
    shutil.rmtree(dst, ignore_errors=True)
    t0 = time.perf_counter()
    st = model_free_ptq(src, dst, scheme="FP8_BLOCK", ignore=["lm_head"], max_workers=workers)
    dt = time.perf_counter() - t0
    print(f"max_workers={workers}: {total / 1e9:.2f} GB bf16 in {dt:.2f} s = {total / dt / 1e9:.1f} GB/s; {st}", flush=True)
shutil.rmtree(src, ignore_errors=True); shutil.rmtree(dst, ignore_errors=True)
