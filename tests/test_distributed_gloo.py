"""CPU, world_size 2, gloo: the host-side logic of the N > 1 path -- unit partitioning, the statistics exchange
(MIN / MAX / SUM all-reduce) and the token-sharded AWQ search finishing with the same argmin on every rank.
The per-rank arithmetic is the CPU oracle here (the CUDA kernels need a GPU); what is under test is the partition /
reduction / selection code in quantizers_b200.scheduler and quantizers_b200.awq that runs unchanged over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world=2):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r in out:
        if isinstance(r, tuple) and r and r[0] == "error":
            raise AssertionError(r[1])
    return sorted(out, key=lambda t: t[0])


def _entry(fn, rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        q.put((rank, fn(rank, world)))
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback

        q.put(("error", f"rank {rank}: {e}\n{traceback.format_exc()}"))


# ------------------------------------------------------------------------------------------------ partitioning
def test_partition_covers_all_units_once():
    from quantizers_b200.scheduler import owner_of, partition

    for n in (1, 7, 36, 128, 18432):
        for w in (1, 2, 3, 4, 8):
            seen = []
            for r in range(w):
                rg = partition(n, w, r)
                seen.extend(rg)
                for u in rg:
                    assert owner_of(u, n, w) == r
            assert seen == list(range(n))
            sizes = [len(partition(n, w, r)) for r in range(w)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        partition(4, 2, 2)


def test_partition_balanced_covers_every_class_unit_once_and_balances_elements():
    from quantizers_b200 import scheduler as S

    for spec, n in ((S.qwen3_4b(layers=1), 36), (S.glm47_flash(units=1), 48), (S.qwen3_30b_a3b(layers=1, experts=1), 16)):
        sizes = {m.name: m.rows * m.cols * m.per_unit for m in spec.matrices}
        for w in (1, 2, 3, 5, 8):
            parts = S.partition_balanced(spec, n, w)
            assert len(parts) == w
            for m in spec.matrices:
                seen = sorted(u for p in parts for u in p.get(m.name, ()))
                assert seen == list(range(n)), (spec.name, m.name, w)
            loads = [sum(sizes[k] * len(v) for k, v in p.items()) for p in parts]
            assert max(loads) - min(loads) <= max(sizes.values())          # LPT: within one largest item
            assert parts == S.partition_balanced(spec, n, w)                # deterministic: every rank computes the same table
    # NVFP4 siblings that share a global scale ACROSS classes stay on one rank
    tied = S.ModelSpec("t", 4, "layer", [S.MatrixSpec("q", 64, 64, "NVFP4", 1, "qkv"), S.MatrixSpec("kv", 32, 64, "NVFP4", 2, "qkv"),
                                         S.MatrixSpec("o", 64, 64, "NVFP4")])
    for p in S.partition_balanced(tied, 4, 3):
        assert p.get("q", []) == p.get("kv", [])
    # Qwen3-4B on 8 ranks: whole layers give 5 / 4 (max / mean = 1.11); the class partition is within 0.5 %
    spec = S.qwen3_4b(layers=1)
    sizes = {m.name: m.rows * m.cols * m.per_unit for m in spec.matrices}
    loads = [sum(sizes[k] * len(v) for k, v in p.items()) for p in S.partition_balanced(spec, 36, 8)]
    assert max(loads) / (sum(loads) / 8) < 1.005


def _stats_worker(rank, world):
    from quantizers_b200.scheduler import allreduce_stats

    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(64, 32, generator=g)
    mins, maxs, sums = x.amin(0), x.amax(0), x.abs().sum(0)
    allreduce_stats(mins, maxs, sums)
    return mins.tolist(), maxs.tolist(), sums.tolist()  # plain lists: tensors in an mp.Queue need the sender alive


def test_allreduce_stats_matches_unsharded():
    res = _run(_stats_worker)
    xs = [torch.randn(64, 32, generator=torch.Generator().manual_seed(100 + r)) for r in range(2)]
    full = torch.cat(xs)
    for _, (mins, maxs, sums) in res:
        assert mins == full.amin(0).tolist() and maxs == full.amax(0).tolist()  # MIN/MAX: bit-identical for any world size
        assert torch.allclose(torch.tensor(sums), full.abs().sum(0), rtol=1e-6)


def _problem():
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(512, 256, generator=g) * (1 + 3 * torch.rand(256, generator=g))).to(torch.bfloat16)
    w = (torch.randn(128, 256, generator=g) * 0.02).to(torch.bfloat16)
    w[:, 7] *= 20
    return x, w


def _awq_worker(rank, world):
    """Token-sharded search: each rank owns half of the calibration tokens."""
    from oracle import llmc_restated as R
    from oracle import oracle as O
    from quantizers_b200 import awq

    x, w = _problem()
    xs = x[rank * 256:(rank + 1) * 256]
    geom = O.Geom(O.GROUP, 32)
    xsum = xs.abs().float().sum(0)
    cnt = torch.tensor([float(xs.shape[0])], dtype=torch.float64)
    x_mean = awq.reduce_token_stats(xsum, cnt)
    w_mean = R.compute_layer_means([w], 32)
    n_grid = 20
    acc = torch.zeros(n_grid + 1)
    ref = R.linear_parent([w], xs)
    for i in range(n_grid):
        s = R.awq_scales(x_mean, w_mean, i / n_grid, True)
        wq = R.scaled_fake_quantize(w, s, geom, O.INT, 4, True)
        acc[i] = (ref - R.linear_parent([wq], xs)).float().pow(2).sum()
    acc[n_grid] = ref.numel()
    best_i, losses = awq.reduce_and_select(acc)
    return best_i, losses, x_mean.tolist()


def test_token_sharded_awq_search_agrees_with_unsharded():
    from oracle import llmc_restated as R
    from oracle import oracle as O

    res = _run(_awq_worker)
    x, w = _problem()
    _, r_ref, l_ref = R.compute_best_scale([x], [w], R.linear_parent, O.Geom(O.GROUP, 32), O.INT, 4, True)
    (_, (b0, l0, m0)), (_, (b1, l1, m1)) = res
    assert b0 == b1 and l0 == l1 and m0 == m1      # every rank ends with identical totals and argmin
    assert b0 / 20 == r_ref
    assert max(abs(a - b) / b for a, b in zip(l0, l_ref)) < 1e-3


def test_reduce_and_select_first_minimum_and_failure():
    from quantizers_b200 import awq

    acc = torch.tensor([3.0, 1.0, 1.0, 2.0, 10.0])
    i, losses = awq.reduce_and_select(acc, distributed=False)
    assert i == 1 and losses == [0.3, 0.1, 0.1, 0.2]
    with pytest.raises(RuntimeError):
        awq.reduce_and_select(torch.tensor([float("nan"), float("inf"), 4.0]), distributed=False)


# ------------------------------------------------------------------------------------------------ expert-parallel ring
def _ring_problem():
    g = torch.Generator().manual_seed(11)
    T, E, K, H = 96, 9, 3, 16
    contrib = (torch.randn(T, E, H, generator=g) * 3).to(torch.bfloat16)          # expert e's weighted output for token t
    topk = torch.stack([torch.randperm(E, generator=g)[:K] for _ in range(T)])   # routing: K distinct experts per token
    x = torch.randn(T, 5, generator=g)
    return T, E, K, H, contrib, topk, x


def _index_add_reference(T, H, contrib, topk, experts, start=None):
    """transformers' sparse-MoE block: ``final.index_add_(0, tokens_of_e, out_e)`` per expert in ascending order, bf16 tensor."""
    final = torch.zeros(T, H, dtype=torch.bfloat16) if start is None else start.clone()
    for e in experts:
        tok = (topk == e).any(dim=1).nonzero().squeeze(1)
        if tok.numel():
            final.index_add_(0, tok, contrib[tok, e])
    return final


def _ring_worker(rank, world):
    """Each rank owns an ascending expert range and a token shard of unequal length; shards are all-gathered, the running output
    travels 0 -> 1 -> ... -> N-1 (the exchange of awq.search_moe_block_mapping_ep with gloo send/recv)."""
    from quantizers_b200 import awq
    from quantizers_b200.scheduler import partition

    T, E, K, H, contrib, topk, x = _ring_problem()
    cuts = [0, 40, 96] if world == 2 else [0, 17, 50, 96]
    mine = slice(cuts[rank], cuts[rank + 1])
    x_all = awq._all_gather_rows(x[mine].contiguous(), dist.group.WORLD)
    topk_all = awq._all_gather_rows(topk[mine].contiguous(), dist.group.WORLD)
    assert torch.equal(x_all, x) and torch.equal(topk_all, topk)
    experts = partition(E, world, rank)
    running = None
    if rank > 0:
        running = torch.empty(T, H, dtype=torch.bfloat16)
        dist.recv(running, src=rank - 1)
    running = _index_add_reference(T, H, contrib, topk_all, experts, running)
    if rank + 1 < world:
        dist.send(running, dst=rank + 1)
        return None
    return running.view(torch.int16).tolist()


@pytest.mark.parametrize("world", [2, 3])
def test_expert_parallel_ring_reproduces_the_unsharded_rounding_sequence(world):
    T, E, K, H, contrib, topk, _ = _ring_problem()
    want = _index_add_reference(T, H, contrib, topk, range(E))
    res = _run(_ring_worker, world)
    got = torch.tensor(res[-1][1], dtype=torch.int16).view(torch.bfloat16)
    assert torch.equal(got.view(torch.int16), want.view(torch.int16))
    # the alternative exchange -- fp32 sum of per-rank partial outputs, rounded once -- is a different function
    from quantizers_b200.scheduler import partition

    parts = sum(_index_add_reference(T, H, contrib, topk, partition(E, world, r)).float() for r in range(world))
    assert not torch.equal(parts.to(torch.bfloat16).view(torch.int16), want.view(torch.int16))
