"""MoE calibration plumbing (quantizers_b200/moe_calibration.py): per-expert Linears + calibrate-all-experts forward, checked
against transformers' own sparse block (CPU, fp32)."""
import pytest
import torch

from quantizers_b200 import moe_calibration as MC


def _hf_block(E=8, H=64, I=32, k=2, norm=True, seed=0):
    qm = pytest.importorskip("transformers.models.qwen3_moe.modeling_qwen3_moe")
    from transformers import Qwen3MoeConfig

    cfg = Qwen3MoeConfig(hidden_size=H, moe_intermediate_size=I, num_experts=E, num_experts_per_tok=k, norm_topk_prob=norm,
                         num_hidden_layers=1, num_attention_heads=2, num_key_value_heads=2, vocab_size=32)
    torch.manual_seed(seed)
    blk = qm.Qwen3MoeSparseMoeBlock(cfg)
    with torch.no_grad():
        for p in blk.parameters():
            p.copy_(torch.randn_like(p) * 0.2)
    return blk


class _Holder(torch.nn.Module):
    def __init__(self, blk):
        super().__init__()
        self.mlp = blk

    def forward(self, x):
        y = self.mlp(x)
        return y[0] if isinstance(y, tuple) else y


@pytest.mark.parametrize("norm", [True, False])
def test_block_output_unchanged(norm):
    blk = _hf_block(norm=norm)
    model = _Holder(blk)
    x = torch.randn(3, 17, 64)
    with torch.no_grad():
        ref = model(x)
        with MC.moe_calibrate_all_experts(model) as names:
            assert names == ["mlp"]
            assert isinstance(model.mlp, MC.CalibrationSparseMoeBlock) and model.mlp.calibrate_all_experts
            got_all = model(x)
        assert not model.mlp.calibrate_all_experts          # linearized for good, sparse again
        got_sparse = model(x)
    torch.testing.assert_close(got_all, ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(got_sparse, ref, rtol=1e-5, atol=1e-6)
    lin = [n for n, m in model.named_modules() if isinstance(m, torch.nn.Linear)]
    assert "mlp.experts.0.gate_proj" in lin and "mlp.experts.7.down_proj" in lin and len(lin) == 24


def test_every_expert_sees_every_token():
    model = _Holder(_hf_block(E=8, k=2))
    x = torch.randn(2, 16, 64)
    seen = {}

    def hook(name):
        def f(mod, args):
            seen[name] = seen.get(name, 0) + args[0].shape[0]
        return f

    with torch.no_grad(), MC.moe_calibrate_all_experts(model):
        hs = [m.register_forward_pre_hook(hook(n)) for n, m in model.named_modules() if n.endswith("gate_proj")]
        model(x)
        assert len(seen) == 8 and set(seen.values()) == {32}
        for h in hs:
            h.remove()
    seen.clear()
    with torch.no_grad():
        hs = [m.register_forward_pre_hook(hook(n)) for n, m in model.named_modules() if n.endswith("gate_proj")]
        model(x)
        for h in hs:
            h.remove()
    assert sum(seen.values()) == 32 * 2                      # sparse: each token reaches exactly k experts


def test_linear_router_and_module_list_experts():
    """transformers-4 style block: Linear router, ModuleList of MLP experts, top_k / norm_topk_prob on the block."""
    torch.manual_seed(1)
    E, H, I, k = 4, 32, 16, 2

    class MLP(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w1, self.w3, self.w2 = torch.nn.Linear(H, I, bias=False), torch.nn.Linear(H, I, bias=False), torch.nn.Linear(I, H, bias=False)

        def forward(self, x):
            return self.w2(torch.nn.functional.silu(self.w1(x)) * self.w3(x))

    class Block(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate = torch.nn.Linear(H, E, bias=False)
            self.experts = torch.nn.ModuleList([MLP() for _ in range(E)])
            self.top_k, self.norm_topk_prob = k, True

        def forward(self, x):
            shp = x.shape
            x = x.reshape(-1, H)
            p = torch.softmax(self.gate(x), dim=-1, dtype=torch.float)
            w, idx = torch.topk(p, k, dim=-1)
            w = (w / w.sum(-1, keepdim=True)).to(x.dtype)
            out = torch.zeros_like(x)
            for t in range(x.shape[0]):
                for j in range(k):
                    out[t] += w[t, j] * self.experts[int(idx[t, j])](x[t:t + 1])[0]
            return out.reshape(shp)

    model = _Holder(Block())
    x = torch.randn(1, 9, H)
    with torch.no_grad():
        ref = model(x)
        with MC.moe_calibrate_all_experts(model):
            got = model(x)
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-6)
    assert model.mlp.experts[0].w1.weight.shape == (I, H)    # expert modules kept as they were


def test_blocks_with_shared_experts_are_left_alone():
    blk = _hf_block()
    blk.shared_expert = torch.nn.Linear(64, 64)
    model = _Holder(blk)
    assert MC.replace_moe_blocks(model) == []


def test_linearize_rejects_inconsistent_shapes():
    class Bad(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.gate_up_proj = torch.nn.Parameter(torch.zeros(2, 8, 4))
            self.down_proj = torch.nn.Parameter(torch.zeros(2, 4, 5))

    with pytest.raises(ValueError):
        MC.linearize_experts(Bad())
