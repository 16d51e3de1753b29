"""CPU: the AWQ / observer restatement (oracle/llmc_restated.py) -- internal consistency and, when live
compressed_tensors is importable, agreement of its quantization leg with the CT fake_quantize it stands in for."""
import pytest
import torch

from oracle import ct_live as L
from oracle import llmc_restated as R
from oracle import oracle as O
from tests.util import assert_bits_equal


def _problem(T=192, K=256, N=96, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(T, K, generator=g) * (1 + 3 * torch.rand(K, generator=g))).to(torch.bfloat16)
    w = (torch.randn(N, K, generator=g) * 0.02).to(torch.bfloat16)
    w[:, 7] *= 20
    return x, w


def test_best_scale_search_is_sane():
    x, w = _problem()
    geom = O.Geom(O.GROUP, 32)
    s, ratio, losses = R.compute_best_scale([x[:96], x[96:]], [w], R.linear_parent, geom, O.INT, 4, True)
    assert len(losses) == 20 and all(l > 0 and l == l for l in losses)
    assert ratio == losses.index(min(losses)) / 20  # first minimum
    assert s.shape == (256,) and torch.isfinite(s).all()
    # ratio 0 => scales are all ones (x_mean^0 / (w_mean^1 + 1e-4), normalised by sqrt(max*min)) only when w_mean is flat;
    # duo_scaling off at ratio 0 gives exactly ones => plain RTN loss
    s0 = R.awq_scales(torch.rand(256) + 0.1, None, 0.0, False)
    assert torch.equal(s0, torch.ones(256))


@pytest.mark.skipif(not L.available(), reason="compressed_tensors not importable")
def test_scaled_fake_quantize_matches_ct_primitives():
    """The inner step built on the oracle == the same step built on live CT calls."""
    from compressed_tensors.quantization.lifecycle.forward import fake_quantize
    from compressed_tensors.quantization.utils.helpers import calculate_qparams

    x, w = _problem(seed=3)
    xm, _ = R.accumulate_abs_mean([x])
    for name, geom, qtype, nb, sym in (("int4_g32_sym", O.Geom(O.GROUP, 32), O.INT, 4, True),
                                       ("int4_g128_asym", O.Geom(O.GROUP, 128), O.INT, 4, False),
                                       ("fp8_g32", O.Geom(O.GROUP, 32), O.FP8, 8, True)):
        _, args = L.format_args(name)
        wm = R.compute_layer_means([w], geom.group)
        s = R.awq_scales(xm, wm, 0.35, True)
        got = R.scaled_fake_quantize(w, s, geom, qtype, nb, sym)
        ws = w.clone().mul_(s.view(1, -1))
        mn, mx = L.observe_minmax(ws, args)
        sc, zp = calculate_qparams(mn, mx, args)
        ref = (fake_quantize(ws, sc, zp, args) / s.view(1, -1)).to(w.dtype)
        assert_bits_equal(got, ref, name)


def test_observers_and_mse():
    x, w = _problem(seed=5)
    obs = R.MinMaxObserver("static_minmax")
    obs.update(torch.tensor([-1.0]), torch.tensor([2.0]))
    mn, mx = obs.update(torch.tensor([-0.5]), torch.tensor([3.0]))
    assert mn.item() == -1.0 and mx.item() == 3.0
    ema = R.MinMaxObserver("minmax")
    ema.update(torch.tensor([0.0]), torch.tensor([1.0]))
    mn, mx = ema.update(torch.tensor([1.0]), torch.tensor([2.0]))
    assert abs(mx.item() - 1.01) < 1e-6
    geom = O.Geom(O.GROUP, 32)
    bmn, bmx = R.mse_minmax(w, geom, O.INT, 4, True)
    mn0, mx0 = O.minmax(w, geom)
    assert (bmx.float() <= mx0.float()).all() and (bmn.float() >= mn0.float()).all()
    gs = R.activation_global_scale([x[:10], x[10:0], x[10:]])  # middle batch empty: skipped
    amax = max(abs(float(x.float().min())), abs(float(x.float().max())))
    assert gs.item() == O.generate_gparam(-amax, amax, torch.bfloat16).item()
