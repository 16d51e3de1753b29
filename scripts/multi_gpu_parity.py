"""Multi-GPU parity check (run under torchrun on >= 2 GPUs; NCCL):
  1. token-sharded AWQ search (|x| sums, token counts and [n_grid] losses all-reduced) == the unsharded search: same argmin
     ratio, losses within 1e-3 relative, identical best scales up to fp32 summation order;
  2. token-sharded layer-wide MoE mapping == unsharded;
  3. expert-sharded NVFP4 compress: every rank's slice is bit-identical to the same experts compressed in one piece (no collective).
Prints one JSON line from rank 0."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from quantizers_b200 import awq, ops, scheduler as S

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
out = {"world": world}
g = torch.Generator(device=dev).manual_seed(7)          # same seed on every rank: identical tensors
T, K, N = 8192, 1536, 3072
x = (torch.randn(T, K, generator=g, device=dev) * (1 + 3 * torch.rand(K, generator=g, device=dev))).to(torch.bfloat16)
w = (torch.randn(N, K, generator=g, device=dev) * 0.02).to(torch.bfloat16)
qa = S.PRESETS["INT4_G32_SYM"]
s_full, r_full, l_full = awq.compute_best_scale(x, [w], awq.linear_parent, qa)
tok = S.partition(T, world, rank)
s_sh, r_sh, l_sh = awq.compute_best_scale(x[tok.start:tok.stop].contiguous(), [w], awq.linear_parent, qa, process_group=dist.group.WORLD)
out["awq_linear"] = {"same_ratio": r_full == r_sh, "max_rel_loss_diff": max(abs(a - b) / b for a, b in zip(l_sh, l_full)),
                     "max_rel_scale_diff": float(((s_sh - s_full).abs() / s_full.abs()).max())}
E, H, I, k = 8, 512, 768, 2
w1 = (torch.randn(E, I, H, generator=g, device=dev) * 0.05).to(torch.bfloat16)
w3 = (torch.randn(E, I, H, generator=g, device=dev) * 0.05).to(torch.bfloat16)
w2 = (torch.randn(E, H, I, generator=g, device=dev) * 0.05).to(torch.bfloat16)
xm = (torch.randn(T, H, generator=g, device=dev) * (1 + 3 * torch.rand(H, generator=g, device=dev))).to(torch.bfloat16)
p = torch.softmax(torch.randn(T, E, generator=g, device=dev), dim=-1)
tw, ti = torch.topk(p, k, dim=-1)
tw = tw / tw.sum(-1, keepdim=True)
_, r_full, l_full = awq.search_moe_block_mapping(xm, w1, w3, w2, ti, tw, qa)
sl = slice(tok.start, tok.stop)
_, r_sh, l_sh = awq.search_moe_block_mapping(xm[sl].contiguous(), w1, w3, w2, ti[sl].contiguous(), tw[sl].contiguous(), qa, process_group=dist.group.WORLD)
out["awq_moe_block"] = {"same_ratio": r_full == r_sh, "max_rel_loss_diff": max(abs(a - b) / b for a, b in zip(l_sh, l_full))}
experts = 16
wall = S.synth_stack(list(range(experts * 2)), 768, 2048, 0, dev)     # gate/up pairs of 16 experts
whole = ops.compress_weight(wall, S.PRESETS["NVFP4"], fuse_span=2)
mine = S.partition(experts, world, rank)
part = ops.compress_weight(wall[2 * mine.start:2 * mine.stop].contiguous(), S.PRESETS["NVFP4"], fuse_span=2)
same = all(torch.equal(part[kk].view(torch.uint8) if part[kk].dtype != torch.float32 else part[kk],
                       (whole[kk][2 * mine.start:2 * mine.stop]).view(torch.uint8) if part[kk].dtype != torch.float32 else whole[kk][2 * mine.start:2 * mine.stop])
           for kk in part)
flag = torch.tensor([1 if same else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
out["nvfp4_expert_shards_bit_identical"] = bool(flag.item())
if rank == 0:
    print(json.dumps(out), flush=True)
dist.destroy_process_group()
