"""The exactness argument behind the single-evaluation kernels (quant_group_tma.cu ONE, awq_fq_fast.cu, div_const_bf16_1):

    for bf16 x and a bf16 (or short-constant) divisor s, the exact quotient x / s is never closer than 2^-17 (relative) to a bf16
    rounding boundary, so ANY approximation of the quotient with relative error below 2^-21 -- x * rcp.approx(s) is within 2^-22 --
    rounds to the same bf16 as the reference's correctly rounded fp32 division.

Brute force over every pair of 8-bit significands (exponents only shift both sides), plus a direct numpy emulation of the two
chains on random operands.  CPU only."""
import numpy as np


def _min_tie_distance(num_sig, den):
    """min over numerators (8-bit significands 128..255, any exponent) of the relative distance of num / den to the nearest
    9-bit-odd significand (= a bf16 rounding boundary)."""
    best = 1.0
    for k in range(0, 12):                       # enough binades to cover every alignment of the quotient's leading bit
        q = num_sig * 2.0 ** k / den
        e = np.floor(np.log2(q))
        t = q / 2.0 ** e * 256                    # quotient significand scaled so that bf16 values are even integers
        d = np.abs(t - (2 * np.floor(t / 2) + 1))
        rel = d / t
        rel = rel[d > 0] if np.any(d == 0) else rel
        best = min(best, float(rel.min()))
    return best


def test_bf16_over_bf16_quotient_never_near_a_rounding_boundary():
    sig = np.arange(128, 256, dtype=np.float64)
    worst = min(_min_tie_distance(sig, float(ms)) for ms in range(128, 256))
    assert worst > 2.0 ** -17, worst             # 1 / (255 * 511) = 2^-16.99 is the analytic bound
    assert worst > 16 * 2.0 ** -21               # > 16x the slack of a reciprocal multiply


def test_constant_divisors_of_calculate_qparams():
    sig = np.arange(128, 256, dtype=np.float64)
    for div in (15.0, 7.5, 448.0, 6.0):          # (max - min) / 15, absmax / 7.5, absmax / 448, absmax / 6
        assert _min_tie_distance(sig, div) > 2.0 ** -14, div


def _bf16_rne(x32):
    u = x32.astype(np.float32).view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16).astype(np.uint32) << 16
    return r.view(np.float32)


def test_reciprocal_multiply_equals_division_on_random_operands():
    rng = np.random.default_rng(0)
    x = _bf16_rne(rng.standard_normal(2_000_000).astype(np.float32) * 0.05)
    s = _bf16_rne(np.abs(rng.standard_normal(2_000_000)).astype(np.float32) * 0.01 + 1e-4)
    ref = _bf16_rne((x / s).astype(np.float32))                       # fp32 division (correctly rounded), then bf16
    r = (np.float32(1.0) / s).astype(np.float32)
    for perturb in (1.0, 1 - 2.0 ** -22, 1 + 2.0 ** -22):             # rcp.approx is within 1 ulp of this
        got = _bf16_rne((x * (r * np.float32(perturb)).astype(np.float32)).astype(np.float32))
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))


def test_nvfp4_group_scale_is_one_multiply_for_every_bf16_statistic():
    """NVFP4 local scale T(absmax / 6) (helpers.py calculate_qparams, TENSOR_GROUP): the fused kernel's compress pass computes it as
    bf16(absmax * fp32(1/6)) with no range test (quant_tile_fast.cu fp4_compress_group).  Exhaustive over every non-negative bf16
    including subnormals and +inf, against torch's bf16 division (fp32 divide, round to bf16)."""
    import torch

    bits = torch.arange(0, 0x7F81, dtype=torch.int32)                 # 0 .. +inf
    a = (bits << 16).view(torch.float32)
    ref = (a.to(torch.bfloat16) / 6).view(torch.int16)
    one = (a * torch.tensor(1.0 / 6.0, dtype=torch.float32)).to(torch.bfloat16).view(torch.int16)
    assert torch.equal(ref, one)
