"""CPU: byte-level simulation of the in-place fp16 operand conversions of the opt-in ``f16x3`` kernels.

The kernels rewrite fp32 TMA tiles in shared memory into the canonical UMMA operand layouts without any barrier wider
than a warp, which only works if (a) every thread's destination lines were read by its own warp and (b) the bytes land
where the shared-memory descriptors say.  These tests replay the thread -> address maps of
``gemm_tf32.cu`` (``F16`` split warps) and ``wgrad_tf32.cu`` (``F16`` X conversion, dY transposition) on a numpy byte
array and decode the result with the canonical layout formulas (CUTLASS ``make_umma_desc``: K-major SWIZZLE_128B rows of
64 fp16; MN-major SWIZZLE_128B atoms ``((8,n),(8,k)):((1,LBO),(8,SBO))``).  They pin the index arithmetic; the
hardware behaviour itself is covered by the opt-in GPU tests."""
import numpy as np
import pytest


def _f16_pieces(x):
    h = x.astype(np.float16)
    return h, (x - h.astype(np.float32)).astype(np.float16)


def test_gemm_f16_split_writes_canonical_k_major_tiles():
    rng = np.random.default_rng(1)
    a = rng.standard_normal((128, 64)).astype(np.float32)
    stage = np.zeros(2 * 16384, dtype=np.uint8)                 # raw box 0 (K 0..31) | raw box 1 (K 32..63)
    for box in range(2):
        for r in range(128):
            for c in range(8):                                   # TMA SWIZZLE_128B: 16-byte chunk c of row r at c ^ (r % 8)
                o = box * 16384 + r * 128 + ((c ^ (r & 7)) << 4)
                stage[o:o + 16] = a[r, box * 32 + 4 * c: box * 32 + 4 * c + 4].view(np.uint8)
    out = stage.copy()
    written_by = {}
    for warp in range(8):                                        # split warps 2..9
        reads = []
        for lane in range(32):
            row, half = warp * 16 + (lane & 15), lane >> 4
            line = half * 16384 + row * 128
            reads.append((row, half, [stage[line + ((c ^ (row & 7)) << 4):][:16].view(np.float32).copy()
                                      for c in range(8)]))
        lines_read = {(h, r) for r, h, _ in reads}
        for row, half, v in reads:                               # after __syncwarp
            for j in range(4):
                hi, lo = _f16_pieces(np.concatenate([v[2 * j], v[2 * j + 1]]))
                off = row * 128 + (((half * 4 + j) ^ (row & 7)) << 4)
                for tile, val in ((0, hi), (1, lo)):
                    assert (tile, row) in lines_read             # a line is only overwritten by the warp that read it
                    assert (tile, off) not in written_by
                    written_by[(tile, off)] = warp
                    out[tile * 16384 + off: tile * 16384 + off + 16] = val.view(np.uint8)
    hi_ref, lo_ref = _f16_pieces(a)
    for tile, ref in ((0, hi_ref), (1, lo_ref)):
        got = np.zeros((128, 64), np.float16)
        for r in range(128):
            for k in range(64):                                  # canonical K-major SWIZZLE_128B, fp16
                o = tile * 16384 + r * 128 + (((k // 8) ^ (r & 7)) << 4) + (k % 8) * 2
                got[r, k] = out[o:o + 2].view(np.float16)[0]
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("bn", [64, 128, 192, 256])
def test_wgrad_f16_x_conversion_writes_canonical_mn_major_atoms(bn):
    rng = np.random.default_rng(bn)
    pairs, boxes = bn // 64, bn // 32
    x = rng.standard_normal((32, bn)).astype(np.float32)         # [pixel][channel] of one stage
    st = np.zeros(boxes * 4096, dtype=np.uint8)
    for b in range(boxes):
        for r in range(32):
            for c in range(8):
                o = b * 4096 + r * 128 + ((c ^ (r & 7)) << 4)
                st[o:o + 16] = x[r, b * 32 + 4 * c: b * 32 + 4 * c + 4].view(np.uint8)
    out = st.copy()
    for warp in range(4):                                        # converting warps 6..9: 8 pixel rows of EVERY box each
        reads = []
        for lane in range(32):
            row, half = warp * 8 + (lane & 7), (lane >> 3) & 1
            for it in range((pairs + 1) // 2):
                g = 2 * it + (lane >> 4)
                if g < pairs:
                    line = (2 * g + half) * 4096 + row * 128
                    reads.append((row, half, g, [st[line + ((c ^ (row & 7)) << 4):][:16].view(np.float32).copy()
                                                 for c in range(8)]))
        rows_read = {r for r, _, _, _ in reads}
        assert rows_read == set(range(warp * 8, warp * 8 + 8))
        for row, half, g, v in reads:                            # after __syncwarp: destinations stay inside the warp's rows
            for j in range(4):
                hi, lo = _f16_pieces(np.concatenate([v[2 * j], v[2 * j + 1]]))
                off = row * 128 + (((half * 4 + j) ^ (row & 7)) << 4)
                out[g * 4096 + off: g * 4096 + off + 16] = hi.view(np.uint8)
                out[(pairs + g) * 4096 + off: (pairs + g) * 4096 + off + 16] = lo.view(np.uint8)
    hi_ref, lo_ref = _f16_pieces(x)
    for base, ref in ((0, hi_ref), (pairs * 4096, lo_ref)):      # descriptors: LBO = 4096 (next 64 channels), SBO = 1024
        got = np.zeros((32, bn), np.float16)
        for p in range(32):
            for c in range(bn):
                o = base + (c // 64) * 4096 + (p // 8) * 1024 + (p % 8) * 128 + ((((c % 64) // 8) ^ (p % 8)) << 4) \
                    + (c % 8) * 2
                got[p, c] = out[o:o + 2].view(np.float16)[0]
        assert np.array_equal(got, ref)


def test_wgrad_dy_transposition_reads_the_atom32b_swizzle():
    """dY boxes keep TMA's SWIZZLE_128B_ATOM_32B: pixel r, channel c at r*128 + (((c>>3) ^ (r&3)) << 5) + (c&7)*4; lane =
    channel reads pixel pairs (2u, 2u+1) for the packed fp16 columns of tensor memory."""
    rng = np.random.default_rng(3)
    dy = rng.standard_normal((32, 32)).astype(np.float32)
    box = np.zeros(4096, np.uint8)
    for r in range(32):
        for c in range(32):
            o = r * 128 + (((c >> 3) ^ (r & 3)) << 5) + (c & 7) * 4
            box[o:o + 4] = dy[r, c:c + 1].view(np.uint8)
    for lane in range(32):
        c8, base = lane >> 3, (lane & 7) * 4
        for u in range(16):
            for r in (2 * u, 2 * u + 1):
                assert box[base + r * 128 + ((c8 ^ (r & 3)) << 5):][:4].view(np.float32)[0] == dy[r, lane]


def test_f16_scale_rule_matches_the_kernel_and_never_overflows():
    """f16_split_scale (tc05.cuh): s = 2^(140 - biased exponent of amax), clamped; |x|*s < 2^14 and inv = 1/s exactly."""
    for amax in [1.0, 0.99999, 4.5, 65504.0, 3e-10, 1e-30, 2.0 ** -113, 1e38, 0.0]:
        eb = (np.float32(amax).view(np.uint32) >> 23) & 0xff
        se = min(max(267 - int(eb), 1), 254)
        s = np.uint32(se << 23).view(np.float32)
        inv = np.uint32((254 - se) << 23).view(np.float32)
        if 0 < se < 254:
            assert float(s) * float(inv) == 1.0
        if amax > 0 and se > 1:
            assert float(np.float32(amax)) * float(s) < 2.0 ** 14
        hi = np.float32(np.float32(amax) * s).astype(np.float16)
        assert np.isfinite(hi)
