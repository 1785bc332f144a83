"""CPU simulation of the arithmetic the window-attention kernels run on kind::f16 (csrc/wmsa3d.cu):
fp16 hi / lo operand pairs under per-tensor power-of-two scales.  numpy float16 restates what the kernels do per element;
the references are float64.  (The kernels themselves are compared with the reference on the GPU: tests/test_gpu_swin.py.)"""
import numpy as np
import pytest


def split_scale(amax):
    """f16_split_scale (csrc/tc05.cuh): s = 2^(140 - eb) for the biased exponent eb of amax, clamped; inv = 1 / s."""
    eb = (np.float32(amax).view(np.uint32) >> 23) & 0xff
    se = int(np.clip(267 - int(eb), 1, 254))
    s = np.uint32(se << 23).view(np.float32)
    inv = np.uint32((254 - se) << 23).view(np.float32)
    return s, inv


def pair(x):
    """fp16 hi / lo pair of a float32 array: hi = rn16(x), lo = rn16(x - hi)."""
    x = np.asarray(x, np.float32)
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi, lo


@pytest.mark.parametrize("amax", [1e-30, 3e-7, 1e-3, 0.7, 1.0, 5.3, 1e4, 6e18])
def test_split_scale_is_an_exact_power_of_two_below_2_to_14(amax):
    s, inv = split_scale(amax)
    assert float(s) * float(inv) == 1.0
    assert float(np.float32(amax)) * float(s) < 2.0 ** 14
    assert float(np.float32(amax)) * float(s) >= 2.0 ** 12          # and not wastefully small
    m, e = np.frexp(float(s))
    assert m == 0.5                                                   # a power of two


def test_pairs_carry_22_mantissa_bits():
    g = np.random.default_rng(0)
    x = (g.standard_normal(20000) * 3.0).astype(np.float32)
    s, _ = split_scale(np.abs(x).max())
    xs = x * s
    hi, lo = pair(xs)
    err = np.abs(xs.astype(np.float64) - (hi.astype(np.float64) + lo.astype(np.float64)))
    # relative 2^-22 where lo is a normal fp16, else the subnormal spacing of fp16 (2^-24) halved
    assert np.all(err <= np.maximum(np.abs(xs) * 2.0 ** -22, 2.0 ** -25))
    assert np.all(np.isfinite(hi.astype(np.float32))) and np.all(np.isfinite(lo.astype(np.float32)))


@pytest.mark.parametrize("n_keys,spread", [(392, 4.0), (392, 30.0), (98, 1.0)])
def test_pv_product_on_pairs_is_fp32_grade(n_keys, spread):
    """O = P V with P = 2^(t - m + 10) and V * sv as fp16 pairs, three products (hi*hi + hi*lo + lo*hi), fp32 accumulation;
    undone by sv_inv / l where l carries P's factor 2^10."""
    g = np.random.default_rng(1)
    t = (g.standard_normal((64, n_keys)) * spread).astype(np.float32)          # log2-domain scores
    v = (g.standard_normal((n_keys, 32)) * 1.7).astype(np.float32)
    m = t.max(1, keepdims=True)
    p = np.exp2((t - (m - np.float32(10.0))).astype(np.float32)).astype(np.float32)
    assert p.max() <= 1024.0
    sv, sv_inv = split_scale(np.abs(v).max() * 1.3)                            # the range scalar covers q and k as well
    ph, pl = pair(p)
    vh, vl = pair(v * sv)
    f = lambda a: a.astype(np.float32)
    acc = (f(pl) @ f(vh)).astype(np.float32) + (f(ph) @ f(vl)).astype(np.float32) + (f(ph) @ f(vh)).astype(np.float32)
    l = p.sum(1, keepdims=True, dtype=np.float32)
    out = acc * (sv_inv / l)
    p64 = np.exp2(t.astype(np.float64) - m.astype(np.float64))
    ref = (p64 / p64.sum(1, keepdims=True)) @ v.astype(np.float64)
    assert np.abs(out - ref).max() <= 3e-6 * np.abs(ref).max()


def test_ds_operand_stays_inside_fp16_under_worst_case_inputs():
    """dS_e = P (dP_acc - D sd sq) 2^-20 with |dO sd|, |v sq| < 2^14 and 32 channels: |dS_e| <= 2^14 < 65504."""
    sq_v, sd_do = 2.0 ** 14, 2.0 ** 14                      # worst case: every scaled element at the bound
    dp_acc = 32 * sq_v * sd_do                              # all 32 products aligned
    d_scaled = 32 * sq_v * sd_do                            # D_i = dO . O, |O| <= max|v|
    assert 1.0 * (dp_acc + d_scaled) * 2.0 ** -20 <= 2.0 ** 14


def test_in_place_operand_columns_do_not_overlap_unread_scores():
    """Backward: a row thread (column half h) owns score columns [16h, 16h + 16) of a chunk and writes its dS pairs back
    into them (hi pairs in the first 8, lo pairs in the next 8); the accumulate MMAs read K step k at 16k (hi) / 16k + 8 (lo).
    Forward: a 32-key chunk's P pairs take columns [0, 16) (hi) and [16, 32) (lo) of the chunk, a 16-key last chunk 8 + 8."""
    for h in range(2):
        own = set(range(16 * h, 16 * h + 16))
        hi_cols, lo_cols = set(range(16 * h, 16 * h + 8)), set(range(16 * h + 8, 16 * h + 16))
        assert hi_cols | lo_cols == own and not (hi_cols & lo_cols)
        assert hi_cols == set(range(16 * h, 16 * h + 8)) and min(lo_cols) == 16 * h + 8          # what the issuer addresses
    for cols in (32, 16):
        hi0, lo0 = 0, cols // 2
        assert lo0 + cols // 2 <= cols                                                            # never past the chunk
