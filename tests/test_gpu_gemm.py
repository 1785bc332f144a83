"""GPU: the tcgen05 3xTF32 GEMM / implicit-GEMM convolution (K6/K8) against float64 references.

Tolerance: the split keeps ~21 mantissa bits per product, so |err| <= 2e-6 * sum_k |a||b| (the same error class as an
fp32 FFMA dot product); a plain single-pass TF32 GEMM misses this bound by three orders of magnitude."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _err_ok(got, ref64, absprod64, k=1024, tol=2e-6):
    """fp32 accumulation error grows like sqrt(K) (random walk), exactly as for an FFMA dot product."""
    err = (got.double() - ref64).abs()
    bound = tol * max(1.0, (k / 1024.0) ** 0.5) * absprod64 + 1e-30
    worst = float((err / bound).max())
    assert worst <= 1.0, "max err/bound = %.3f (max |err| %.3e)" % (worst, float(err.max()))
    return float((err / (absprod64 + 1e-30)).max())


@pytest.mark.parametrize("m,n,k", [(128, 64, 32), (128, 128, 64), (256, 256, 128), (200, 96, 100), (1000, 320, 256),
                                   (6272, 2048, 512), (25088, 64, 256), (392, 384, 128), (37, 5, 36)])
@pytest.mark.parametrize("force_bn", [0, 64, 128, 256, 64 | 0x2000, 128 | 0x2000])   # 0x2000: tensor-memory A operands
def test_gemm_tf32x3_vs_float64(cuda_device, m, n, k, force_bn):
    from vitta_b200 import ops
    g = torch.Generator().manual_seed(m * 7 + n * 3 + k)
    a = torch.randn(m, k, generator=g).to(cuda_device)
    b = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda_device)
    bh, bl = ops.split_tf32(b)
    out = ops.gemm_tf32x3(a, bh, bl, n, force_bn=force_bn)
    torch.cuda.synchronize()
    ref = a.double() @ b.double().t()
    absprod = a.double().abs() @ b.double().abs().t()
    rel = _err_ok(out, ref, absprod, k)
    # and it is an fp32-grade result, not a TF32 one
    assert rel < 2e-6


@pytest.mark.parametrize("bn", [64, 128])
def test_gemm_operand_forms_agree(cuda_device, bn):
    """A from tensor memory and A from shared memory issue the same products on the same hi/lo values; only the number
    of accumulator chains (= the fp32 summation order) differs, so the results agree to fp32 rounding."""
    from vitta_b200 import ops
    g = torch.Generator().manual_seed(bn)
    m, n, k = 3000, 192, 296
    a = torch.randn(m, k, generator=g).to(cuda_device)
    b = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda_device)
    bh, bl = ops.split_tf32(b)
    t = ops.gemm_tf32x3(a, bh, bl, n, force_bn=bn)
    s = ops.gemm_tf32x3(a, bh, bl, n, force_bn=bn | 0x2000)
    assert float((t - s).abs().max()) <= 4e-6 * float(s.abs().max())


def test_gemm_epilogues(cuda_device):
    from vitta_b200 import ops
    g = torch.Generator().manual_seed(11)
    m, n, k = 777, 512, 128
    a = torch.randn(m, k, generator=g).to(cuda_device)
    b = (torch.randn(n, k, generator=g) / k ** 0.5).to(cuda_device)
    bias = torch.randn(n, generator=g).to(cuda_device)
    res = torch.randn(m, n, generator=g).to(cuda_device)
    bh, bl = ops.split_tf32(b)
    ref = a.double() @ b.double().t() + bias.double()
    out = ops.gemm_tf32x3(a, bh, bl, n, bias=bias)
    assert float((out.double() - ref).abs().max()) < 2e-5
    out = ops.gemm_tf32x3(a, bh, bl, n, bias=bias, act=1)
    assert float((out.double() - F.gelu(ref)).abs().max()) < 2e-5
    out = ops.gemm_tf32x3(a, bh, bl, n, bias=bias, residual=res)
    assert float((out.double() - (ref + res.double())).abs().max()) < 2e-5
    # strided A (a column slice of a wider matrix) and output into a strided view
    wide = torch.randn(m, 3 * k, generator=g).to(cuda_device)
    a2 = wide[:, k:2 * k]
    big = torch.zeros(m, 2 * n, device=cuda_device)
    ops.gemm_tf32x3(a2, bh, bl, n, out=big[:, n:])
    ref2 = a2.double() @ b.double().t()
    assert float((big[:, n:].double() - ref2).abs().max()) < 2e-5
    assert float(big[:, :n].abs().max()) == 0.0


@pytest.mark.parametrize("f,h,w,cin,cout,kh,stride,pad", [
    (4, 14, 14, 64, 128, 3, 1, 1), (3, 7, 7, 32, 64, 3, 1, 1), (2, 56, 56, 64, 64, 3, 1, 1), (5, 28, 28, 128, 128, 3, 1, 1),
    (2, 56, 56, 64, 256, 1, 1, 0), (3, 28, 28, 128, 128, 3, 2, 1), (3, 14, 14, 256, 512, 1, 2, 0), (2, 9, 5, 8, 24, 3, 1, 1),
    (2, 30, 30, 4, 64, 7, 2, 3), (16, 7, 7, 512, 512, 3, 1, 1),
])
def test_conv2d_tf32x3_vs_float64(cuda_device, f, h, w, cin, cout, kh, stride, pad):
    from vitta_b200 import ops
    g = torch.Generator().manual_seed(f + h * 3 + cin)
    x = torch.randn(f, cin, h, w, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, kh, kh, generator=g) / (cin * kh * kh) ** 0.5).to(cuda_device)
    wh, wl = ops.split_tf32(wt)
    y = ops.conv2d_tf32x3(x, wh, wl, cout, kh, kh, stride, pad)
    torch.cuda.synchronize()
    ref = F.conv2d(x.double(), wt.double(), None, stride, pad)
    absprod = F.conv2d(x.double().abs(), wt.double().abs(), None, stride, pad)
    assert y.shape == ref.shape
    _err_ok(y, ref, absprod, cin * kh * kh)


@pytest.mark.parametrize("cin,cout,kh", [(64, 128, 3), (128, 64, 1)])
def test_split_mode1_gives_data_gradient(cuda_device, cin, cout, kh):
    """dgrad of a stride-1 conv = the same implicit GEMM on grad_out with the mode-1 (transposed, rotated) operand."""
    from vitta_b200 import ops
    g = torch.Generator().manual_seed(5)
    f, h, w = 3, 14, 14
    pad = kh // 2
    x = torch.randn(f, cin, h, w, generator=g, dtype=torch.float64).to(cuda_device).requires_grad_(True)
    wt = (torch.randn(cout, cin, kh, kh, generator=g) / (cin * kh * kh) ** 0.5).to(cuda_device)
    go = torch.randn(f, cout, h, w, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    F.conv2d(x, wt.double(), None, 1, pad).backward(go.double())
    wh, wl = ops.split_tf32(wt, mode=1)
    gx = ops.conv2d_tf32x3(go, wh, wl, cin, kh, kh, 1, pad)
    scale = float(x.grad.abs().max())
    assert float((gx.double() - x.grad).abs().max()) < 3e-6 * scale * (cout * kh * kh) ** 0.5


@pytest.mark.parametrize("f,h,w,cin,cout,kh,stride,pad", [
    (4, 14, 14, 64, 128, 3, 1, 1), (32, 7, 7, 32, 64, 3, 1, 1), (2, 56, 56, 64, 64, 3, 1, 1), (6, 28, 28, 128, 128, 3, 1, 1),
    (2, 56, 56, 64, 256, 1, 1, 0), (3, 28, 28, 128, 128, 3, 2, 1), (3, 14, 14, 256, 512, 1, 2, 0), (2, 9, 5, 8, 24, 3, 1, 1),
    (16, 7, 7, 512, 512, 3, 1, 1), (64, 14, 14, 1024, 256, 1, 1, 0), (5, 14, 14, 36, 20, 1, 1, 0),
])
def test_conv2d_wgrad_tf32x3_vs_float64(cuda_device, f, h, w, cin, cout, kh, stride, pad):
    from vitta_b200 import ops
    g = torch.Generator().manual_seed(f * 5 + h + cout)
    x = torch.randn(f, cin, h, w, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kh) // stride + 1
    gy = torch.randn(f, cout, ho, wo, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    gw = ops.conv2d_wgrad_tf32x3(x, gy, cout, kh, kh, stride, pad)
    torch.cuda.synchronize()
    wt = torch.zeros(cout, cin, kh, kh, dtype=torch.float64, device=cuda_device, requires_grad=True)
    F.conv2d(x.double(), wt, None, stride, pad).backward(gy.double())
    # sum |x||gy| bound via the same convolution on absolute values
    wa = torch.zeros_like(wt, requires_grad=True)
    F.conv2d(x.double().abs(), wa, None, stride, pad).backward(gy.double().abs())
    assert gw.shape == wt.grad.shape and gw.is_contiguous()
    _err_ok(gw, wt.grad, wa.grad, f * ho * wo)


@pytest.mark.parametrize("f,h,cin,cout,kh,stride,pad", [(6, 28, 128, 128, 3, 1, 1), (16, 14, 256, 64, 1, 1, 0),
                                                        (3, 28, 64, 192, 3, 2, 1)])
def test_conv_operand_forms_agree(cuda_device, f, h, cin, cout, kh, stride, pad):
    """vitta_gemm_set_operand_form: shared-memory A operands (1) against tensor-memory A operands (2) for forward, data
    gradient and weight gradient -- same products, different accumulator-chain count, so equal to fp32 rounding."""
    from vitta_b200 import _lib, ops
    g = torch.Generator().manual_seed(f + cin)
    x = torch.randn(f, cin, h, h, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cout, cin, kh, kh, generator=g) / (cin * kh * kh) ** 0.5).to(cuda_device)
    res = []
    try:
        for form in (1, 2):
            _lib.call("vitta_gemm_set_operand_form", form)
            x1 = x.clone(memory_format=torch.channels_last).requires_grad_(True)
            w1 = wt.clone().requires_grad_(True)
            y = ops.conv2d(x1, w1, stride, pad)
            y.backward(torch.ones_like(y) * 0.5)
            res.append((y.detach(), x1.grad, w1.grad))
    finally:
        _lib.call("vitta_gemm_set_operand_form", 0)
    for a, b in zip(*res):
        assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())


def test_conv2d_autograd_matches_torch(cuda_device):
    """Conv2dFn (forward + dgrad + wgrad on tcgen05) against torch's float64 autograd, stride 1 and 2."""
    import vitta_b200
    from vitta_b200 import ops
    vitta_b200.set_fp32_exact()
    g = torch.Generator().manual_seed(21)
    for (f, h, cin, cout, kh, stride, pad) in [(4, 14, 64, 128, 3, 1, 1), (4, 28, 128, 128, 3, 2, 1), (8, 14, 256, 64, 1, 1, 0),
                                                (3, 28, 256, 512, 1, 2, 0),    # downsample conv: 3 of 4 pixel classes get 0
                                                (2, 15, 64, 64, 3, 2, 1),      # odd extent: ragged residue classes
                                                (5, 14, 512, 512, 3, 2, 1)]:
        x = torch.randn(f, cin, h, h, generator=g).to(cuda_device)
        wt = (torch.randn(cout, cin, kh, kh, generator=g) / (cin * kh * kh) ** 0.5).to(cuda_device)
        x1 = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
        w1 = wt.clone().requires_grad_(True)
        y1 = ops.conv2d(x1, w1, stride, pad)
        go = torch.randn(y1.shape, generator=g).to(cuda_device)
        y1.backward(go)
        x2 = x.double().requires_grad_(True)
        w2 = wt.double().requires_grad_(True)
        y2 = F.conv2d(x2, w2, None, stride, pad)
        y2.backward(go.double())
        for a, b, nm in ((y1, y2, "y"), (x1.grad, x2.grad, "gx"), (w1.grad, w2.grad, "gw")):
            scale = float(b.detach().abs().max())
            err = float((a.detach().double() - b.detach()).abs().max())
            # 3xTF32 drops the lo*lo term (2^-22 per product): the bound grows with sqrt(K); K = 4608 in the last case
            assert err < 5e-5 * scale, (nm, stride, kh, err, scale)


def test_weight_split_cache_survives_address_reuse(cuda_device):
    """Regression for the round-1 parity failure: allocate a weight, convolve, free it, allocate a DIFFERENT weight of
    the same shape (the caching allocator returns the same address, version counter 0 again) and convolve: each result
    must come from its own weight."""
    import gc
    import vitta_b200
    from vitta_b200 import ops
    vitta_b200.set_fp32_exact()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 128, 14, 14, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    ptrs, errs = [], []
    for i in range(4):
        wt = (torch.randn(128, 128, 3, 3, generator=g) / 34.0).to(cuda_device)
        w1 = wt.clone().requires_grad_(True)
        ptrs.append(w1.data_ptr())
        y = ops.conv2d(x, w1, 1, 1)
        y.sum().backward()
        ref = F.conv2d(x.double(), wt.double(), None, 1, 1)
        errs.append(float((y.detach().double() - ref).abs().max()) / float(ref.abs().max()))
        del wt, w1, y, ref
        gc.collect()
    assert max(errs) < 5e-5, (errs, ptrs)
    assert len(set(ptrs)) < len(ptrs), "the allocator never reused an address: the regression was not exercised"


@pytest.mark.parametrize("prec", ["f16x3", "tf32x3"])
def test_multi_tensor_split_refresh_is_bit_identical_to_per_tensor_splits(cuda_device, prec):
    """vitta_split_multi (all weights, both operand forms, one launch sequence after the SGD step) against the per-tensor
    vitta_split_* path: same pieces bit for bit, written into the existing buffers, stamps valid afterwards."""
    from vitta_b200 import ops
    before = ops.gemm_precision()
    ops.set_gemm_precision(prec)
    try:
        g = torch.Generator().manual_seed(11)
        shapes = [(64, 64, 3, 3), (256, 64, 1, 1), (128, 128, 3, 3), (96, 288), (40, 24, 3, 3), (12, 8)]
        ws = [torch.nn.Parameter((torch.randn(*sh, generator=g) * (10.0 ** -i)).to(cuda_device)) for i, sh in enumerate(shapes)]
        split = ops.weight_split_f16 if prec == "f16x3" else ops.weight_split
        first = [[tuple(t.clone() for t in split(w, m)) for m in (0, 1)] for w in ws]
        ptrs = [[tuple(t.data_ptr() for t in split(w, m)) for m in (0, 1)] for w in ws]
        # "optimizer step": change the weights through raw storage writes, announce it, refresh everything at once
        with torch.no_grad():
            for w in ws:
                w.data.mul_(1.5).add_(0.01)
        vers = [w._version for w in ws]
        ops.bump_weight_epoch()
        n = ops.refresh_weight_splits(ws)
        assert n == 2 * len(ws)
        for w, p0, f0 in zip(ws, ptrs, first):
            for m in (0, 1):
                got = split(w, m)                                  # cache hit: the refreshed buffers
                assert tuple(t.data_ptr() for t in got) == p0[m]  # refreshed in place
                make = ops.split_f16 if prec == "f16x3" else ops.split_tf32
                want = make(w.detach(), m)                         # fresh per-tensor split of the NEW weight
                for a, b, old in zip(got, want, f0[m]):
                    assert torch.equal(a, b)
                assert not torch.equal(got[0], f0[m][0])           # and really new
    finally:
        ops.set_gemm_precision(before)
