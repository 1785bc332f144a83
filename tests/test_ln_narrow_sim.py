"""CPU restatement of the sub-warp layout of the narrow-row LayerNorm kernels (csrc/layernorm.cu: ln_fwd_narrow_kernel /
ln_bwd_narrow_kernel): LPR lanes per row, 32 / LPR rows per warp pass, column float4 index j * LPR + sub.  Checks that the
dispatch's (LPR, VPL) pairs cover every column of a row exactly once, that the xor-shuffle ladders stay inside a row
group, and that per-group Welford statistics merged with Chan's formula in the kernel's order equal the statistics of the
whole chunk.  (The kernels run against float64 references on the GPU: tests/test_gpu_swin.py.)"""
import numpy as np
import pytest


def fwd_layout(c4):
    """the forward dispatch of vitta_ln_fwd_amax for 16 < C/4 <= 64"""
    return (8, 3) if c4 <= 24 else (8, 4) if c4 <= 32 else (16, 3) if c4 <= 48 else (16, 4)


def bwd_layout(c4):
    """the backward dispatch of vitta_ln_bwd_amax for 16 < C/4 <= 48"""
    return (8, 3) if c4 <= 24 else (16, 2) if c4 <= 32 else (16, 3)


@pytest.mark.parametrize("c4", list(range(17, 65)))
def test_lane_column_map_covers_each_column_once(c4):
    for lpr, vpl in ([fwd_layout(c4)] + ([bwd_layout(c4)] if c4 <= 48 else [])):
        assert lpr * vpl >= c4
        cols = [j * lpr + sub for sub in range(lpr) for j in range(vpl) if j * lpr + sub < c4]
        assert sorted(cols) == list(range(c4))


@pytest.mark.parametrize("lpr", [8, 16])
def test_shuffle_ladders_stay_inside_their_lanes(lpr):
    lanes = np.arange(32)
    # row reduction: xor offsets lpr/2 .. 1 never leave the row group
    o = lpr // 2
    while o:
        assert np.all((lanes ^ o) // lpr == lanes // lpr)
        o >>= 1
    # group merge: xor offsets lpr .. 16 keep the column (lane % lpr) and reach every other group
    reach = {l: {l} for l in range(32)}
    o = lpr
    while o < 32:
        assert np.all((lanes ^ o) % lpr == lanes % lpr)
        reach = {l: reach[l] | reach[l ^ o] for l in range(32)}
        o <<= 1
    for l in range(32):
        assert reach[l] == {m for m in range(32) if m % lpr == l % lpr}


@pytest.mark.parametrize("lpr,rows", [(8, 64), (8, 61), (16, 64), (16, 7), (8, 3)])
def test_group_welford_plus_chan_merge_equals_chunk_statistics(lpr, rows):
    g = 32 // lpr
    rng = np.random.default_rng(rows * lpr)
    y = (rng.standard_normal(rows) * 2.0 + 0.7).astype(np.float32)       # one column of the chunk's output rows
    n = np.zeros(g, np.float32)
    mean = np.zeros(g, np.float32)
    m2 = np.zeros(g, np.float32)
    for r in range(rows):                                                  # row r goes to group r % g (pass r // g)
        k = r % g
        n[k] += 1
        d = y[r] - mean[k]
        mean[k] = np.float32(mean[k] + d * (np.float32(1.0) / n[k]))
        m2[k] = np.float32(m2[k] + d * (y[r] - mean[k]))
    o = 1
    while o < g:                                                           # xor ladder over the groups, as the kernel runs it
        nn, mm, qq = n.copy(), mean.copy(), m2.copy()
        for k in range(g):
            nb, mb, m2b = n[k ^ o], mean[k ^ o], m2[k ^ o]
            nt = n[k] + nb
            f = nb / nt if nt > 0 else np.float32(0)
            d = mb - mean[k]
            qq[k] = m2[k] + m2b + d * d * n[k] * f
            mm[k] = mean[k] + d * f
            nn[k] = nt
        n, mean, m2 = nn, mm, qq
        o <<= 1
    ref_mean = y.astype(np.float64).mean()
    ref_m2 = ((y.astype(np.float64) - ref_mean) ** 2).sum()
    assert n[0] == rows
    assert abs(mean[0] - ref_mean) <= 2e-6 * max(1.0, abs(ref_mean))
    assert abs(m2[0] - ref_m2) <= 2e-5 * max(1.0, ref_m2)
