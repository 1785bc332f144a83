"""Autograd operators of the Video-Swin path over the C-ABI kernels (include/vitta_b200.h: K7, K8, K9).

Every operator works on the token matrix ``(rows, C)`` of a ``(B, D, H, W, C)`` activation, fp32, CUDA.  The
composite Functions (attention half-block, MLP half-block, PatchMerging, PatchEmbed3D) keep their intermediate
tensors inside one autograd node so that epilogues can be fused (bias / GELU / shortcut / DropPath in the GEMM
store, GELU' in the data-gradient GEMM, the hook gradient and the shortcut gradient in the LayerNorm backward).
No eager-torch arithmetic fallback exists: a missing library raises in ``_lib.load()``.
"""
import ctypes as C

import torch

from . import _lib, ops
from ._lib import call, ptr, stream_ptr


def _chk(t, what):
    if not t.is_cuda or t.dtype != torch.float32:
        raise _lib.VittaError("%s: vitta_b200 kernels need fp32 CUDA tensors (got %s %s); there is no CPU path"
                              % (what, t.dtype, t.device))


_ws = {}


def _workspace(kind, n, dev, zero):
    key = (kind, dev)
    w = _ws.get(key)
    if w is None or w.numel() < n:
        w = (torch.zeros if zero else torch.empty)(int(n), dtype=torch.float32, device=dev)
        _ws[key] = w
    return w


# ----------------------------------------------------------------------------------------------
# raw kernel wrappers (no autograd)
# ----------------------------------------------------------------------------------------------
def gemm(a, w, mode, bias=None, residual=None, act=0, aux_out=None, row_scale=None, rows_per_group=1, out=None,
         want_amax=False):
    """out[M, N] = a[M, K] @ op(w)  with the fused epilogue of vitta_gemm_tf32x3_ex.
    mode 0: w is (N, K) (forward of nn.Linear);  mode 1: w is (K, N) and its transpose is used (data gradient)."""
    m, k = a.shape
    n = w.shape[0] if mode == 0 else w.shape[1]
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32, device=a.device)
    ldr = residual.stride(0) if residual is not None else 0
    if ops.gemm_precision() == "f16x3" and k % 8 == 0 and a.is_contiguous():
        # opt-in fp16 operand split; the amax of the activation operand is a separate pass for now
        hi, lo, wam = ops.weight_split_f16(w, mode)
        if want_amax:
            # the result feeds another fp16-split GEMM: its range comes out of this epilogue instead of an extra pass
            am = ops.new_amax(a.device)
            call("vitta_gemm_f16x3_amax", ptr(a), a.stride(0), ptr(ops.operand_amax(a)), ptr(hi), ptr(lo), ptr(wam), k,
                 ptr(out), out.stride(0), m, n, k, ptr(bias), ptr(residual), ldr, int(act), ptr(aux_out), ptr(row_scale),
                 int(rows_per_group), ptr(am), stream_ptr())
            ops._attach_amax(out, am)
            return out
        call("vitta_gemm_f16x3_ex", ptr(a), a.stride(0), ptr(ops.operand_amax(a)), ptr(hi), ptr(lo), ptr(wam), k, ptr(out),
             out.stride(0), m, n, k, ptr(bias), ptr(residual), ldr, int(act), ptr(aux_out), ptr(row_scale),
             int(rows_per_group), 0, stream_ptr())
        return out
    hi, lo = ops.weight_split(w, mode)
    call("vitta_gemm_tf32x3_ex", ptr(a), a.stride(0), ptr(hi), ptr(lo), k, ptr(out), out.stride(0), m, n, k, ptr(bias),
         ptr(residual), ldr, int(act), ptr(aux_out), ptr(row_scale), int(rows_per_group), 0, stream_ptr())
    return out


def linear_wgrad(x, gy, want_bias=False):
    """dW (N, K) = gy[M, N]^T @ x[M, K] on the split-K tcgen05 weight-gradient kernel (a 1x1 'convolution' over M pixels).
    ``want_bias``: also return db (N,) = column sums of gy -- a by-product of the fp16-split kernel, else vitta_colsum."""
    m, k = x.shape
    n = gy.shape[1]
    f, wdt = 1, m    # pointwise: the kernel walks the rows 32 at a time
    nws = _lib.load().vitta_conv2d_wgrad_ws_floats(f, 1, wdt, k, n, 1, 1, 1, 0)
    if nws <= 0:
        raise _lib.VittaError("linear_wgrad: bad geometry")
    ws = _workspace("wgrad", nws, x.device, False)
    gw = torch.empty(n, k, dtype=torch.float32, device=x.device)
    if ops.gemm_precision() == "f16x3" and x.is_contiguous() and gy.is_contiguous():
        if want_bias:
            db = torch.empty(n, dtype=torch.float32, device=x.device)
            call("vitta_conv2d_wgrad_f16x3_bias", ptr(x), ptr(ops.operand_amax(x)), ptr(gy), ptr(ops.operand_amax(gy)), f, 1,
                 wdt, k, n, 1, 1, 1, 0, ptr(gw), ptr(db), 0, ptr(ws), stream_ptr())
            return gw, db
        call("vitta_conv2d_wgrad_f16x3", ptr(x), ptr(ops.operand_amax(x)), ptr(gy), ptr(ops.operand_amax(gy)), f, 1, wdt, k, n, 1,
             1, 1, 0, ptr(gw), 0, ptr(ws), stream_ptr())
        return gw
    call("vitta_conv2d_wgrad_tf32x3", ptr(x), ptr(gy), f, 1, wdt, k, n, 1, 1, 1, 0, ptr(gw), 0, ptr(ws), stream_ptr())
    return (gw, colsum(gy)) if want_bias else gw


def colsum(g):
    rows, c = g.shape
    n = _lib.load().vitta_colsum_ws_floats(rows, c)
    ws = _workspace(("colsum", c), n, g.device, True)
    out = torch.empty(c, dtype=torch.float32, device=g.device)
    call("vitta_colsum", ptr(g), rows, c, ptr(out), 0, ptr(ws), stream_ptr())
    return out


def row_scale(x, scale, rows_per_group):
    out = torch.empty_like(x)
    if ops.gemm_precision() == "f16x3":       # the scaled gradient is a GEMM operand: range for free
        am = ops.new_amax(x.device)
        call("vitta_row_scale_amax", ptr(x), ptr(scale), x.shape[0], int(rows_per_group), x.shape[1], ptr(out), ptr(am),
             stream_ptr())
        ops._attach_amax(out, am)
        return out
    call("vitta_row_scale", ptr(x), ptr(scale), x.shape[0], int(rows_per_group), x.shape[1], ptr(out), stream_ptr())
    return out


def _gather_struct(gather):
    if gather is None:
        return None, None
    g = _lib.VittaLnGather(*[int(v) for v in gather])
    return g, C.byref(g)


def ln_fwd(x, weight, bias, eps, rows, c, part=None, gather=None):
    y = torch.empty(rows, c, dtype=torch.float32, device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    keep, gp = _gather_struct(gather)
    if ops.gemm_precision() == "f16x3":
        # y feeds fp16-split GEMMs (qkv / fc1 / reduction): its range comes out of this pass
        am = ops.new_amax(x.device)
        call("vitta_ln_fwd_amax", ptr(x), ptr(weight), ptr(bias), float(eps), ptr(y), ptr(mean), ptr(rstd), ptr(part), rows,
             c, gp, ptr(am), stream_ptr())
        ops._attach_amax(y, am)
        return y, mean, rstd
    call("vitta_ln_fwd", ptr(x), ptr(weight), ptr(bias), float(eps), ptr(y), ptr(mean), ptr(rstd), ptr(part), rows, c, gp,
         stream_ptr())
    return y, mean, rstd


def ln_bwd(gy, x, weight, bias, mean, rstd, rows, c, gadd=None, coef=None, gscale=None, gather=None, gx_shape=None):
    dev = gy.device
    n = _lib.load().vitta_ln_bwd_ws_floats(rows, c)
    ws = _workspace(("ln_bwd", c), n, dev, True)
    gx = torch.empty(gx_shape if gx_shape is not None else (rows, c), dtype=torch.float32, device=dev)
    dgamma = torch.zeros(c, dtype=torch.float32, device=dev)
    dbeta = torch.zeros(c, dtype=torch.float32, device=dev)
    ca = cb = cm = gs = None
    if coef is not None:
        ca, cb, cm = coef
        gs = ptr(gscale)
    keep, gp = _gather_struct(gather)
    if ops.gemm_precision() == "f16x3" and gather is None:
        # gx is the residual-stream gradient the previous half-block's GEMMs (proj / fc2 data and weight gradients) take
        am = ops.new_amax(dev)
        call("vitta_ln_bwd_amax", ptr(gy), ptr(x), ptr(weight), ptr(bias), ptr(mean), ptr(rstd), ptr(gadd), ca, cb, cm, gs,
             ptr(gx), ptr(dgamma), ptr(dbeta), ptr(ws), rows, c, gp, ptr(am), stream_ptr())
        ops._attach_amax(gx, am)
        return gx, dgamma, dbeta
    call("vitta_ln_bwd", ptr(gy), ptr(x), ptr(weight), ptr(bias), ptr(mean), ptr(rstd), ptr(gadd), ca, cb, cm, gs, ptr(gx),
         ptr(dgamma), ptr(dbeta), ptr(ws), rows, c, gp, stream_ptr())
    return gx, dgamma, dbeta


def _int3(v):
    return (C.c_int * 3)(int(v[0]), int(v[1]), int(v[2]))


def wmsa3d_fwd(qkv, table, dims, heads, window, shift, scale):
    b, d, h, w = dims
    c = heads * 32
    out = torch.empty(b * d * h * w, c, dtype=torch.float32, device=qkv.device)
    lse = torch.empty(b * d * h * w * heads, dtype=torch.float32, device=qkv.device)
    qam = ops.operand_amax(qkv)     # range of V's fp16 operand split: emitted by the qkv GEMM's epilogue (else one amax pass)
    if ops.gemm_precision() == "f16x3":       # out is the operand of the fp16-split proj GEMM and of its weight gradient
        am = ops.new_amax(qkv.device)
        call("vitta_wmsa3d_fwd_amax", ptr(qkv), ptr(qam), ptr(table), ptr(out), ptr(lse), b, d, h, w, heads, 32,
             _int3(window), _int3(shift), float(scale), ptr(am), stream_ptr())
        ops._attach_amax(out, am)
        return out, lse
    call("vitta_wmsa3d_fwd", ptr(qkv), ptr(qam), ptr(table), ptr(out), ptr(lse), b, d, h, w, heads, 32, _int3(window),
         _int3(shift), float(scale), stream_ptr())
    return out, lse


def wmsa3d_bwd(qkv, table, out, dout, lse, dims, heads, window, shift, scale, impl=0):
    """impl 0: tcgen05 kernels; impl 1: the exact-fp32 FFMA2 kernel (cross-check)."""
    b, d, h, w = dims
    dqkv = torch.empty_like(qkv)
    dtable = torch.zeros_like(table)
    ws = _workspace("wmsa_bwd", b * d * h * w * heads, qkv.device, False)
    # ranges of the fp16 operand splits (impl 0): emitted by the producing GEMM epilogues, else one amax pass each
    qam = ptr(ops.operand_amax(qkv)) if impl == 0 else None
    dam = ptr(ops.operand_amax(dout)) if impl == 0 else None
    if ops.gemm_precision() == "f16x3" and impl == 0:      # dqkv feeds the qkv data / weight gradient GEMMs
        am = ops.new_amax(qkv.device)
        call("vitta_wmsa3d_bwd_amax", ptr(qkv), qam, ptr(table), ptr(out), ptr(dout), dam, ptr(lse), ptr(dqkv), ptr(dtable),
             ptr(ws), b, d, h, w, heads, 32, _int3(window), _int3(shift), float(scale), int(impl), ptr(am), stream_ptr())
        ops._attach_amax(dqkv, am)
        return dqkv, dtable
    call("vitta_wmsa3d_bwd", ptr(qkv), qam, ptr(table), ptr(out), ptr(dout), dam, ptr(lse), ptr(dqkv), ptr(dtable), ptr(ws), b,
         d, h, w, heads, 32, _int3(window), _int3(shift), float(scale), int(impl), stream_ptr())
    return dqkv, dtable


# ----------------------------------------------------------------------------------------------
# K9: LayerNorm (+ statistics tap) returning the normalised rows AND an alias of the input for the shortcut, so
# that the backward kernel adds the shortcut gradient itself
# ----------------------------------------------------------------------------------------------
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, arena, ly, want_alias):
        _chk(x, "layer_norm")
        rows, c = x.shape
        part = None
        if ly is not None:
            part = arena.partial_buffer(ly, rows, c, 1, 1, x.device, ln=True)
        y, mean, rstd = ln_fwd(x, weight, bias, eps, rows, c, part)
        ctx.save_for_backward(x, weight, bias, mean, rstd)
        ctx.meta = (arena, ly, want_alias)
        tok = ops.new_token(x) if ly is not None else None
        alias = x.view_as(x) if want_alias else None
        return y, alias, tok

    @staticmethod
    def backward(ctx, gy, galias, gtok):
        x, weight, bias, mean, rstd = ctx.saved_tensors
        arena, ly, want_alias = ctx.meta
        rows, c = x.shape
        if gy is None:
            gy = torch.zeros_like(x)
        gy = gy.contiguous()
        coef = gs = None
        if ly is not None and gtok is not None:
            coef = arena.coef_ptrs(ly)
            gs = gtok.contiguous()
        gadd = galias.contiguous() if galias is not None else None
        gx, dg, db = ln_bwd(gy, x, weight, bias, mean, rstd, rows, c, gadd, coef, gs)
        return gx, dg, db, None, None, None, None


def layer_norm_rows(x, weight, bias, eps, arena=None, ly=None, want_alias=True):
    y, alias, tok = LayerNormFn.apply(x, weight, bias, eps, arena, ly, want_alias)
    if ly is not None:
        ly.token = tok
    return y, alias


# ----------------------------------------------------------------------------------------------
# K8 + K7: attention half-block   x + DropPath(proj(W-MSA(qkv(y))))
# ----------------------------------------------------------------------------------------------
def padded_token_dims(dims, window):
    """(B, Dp, Hp, Wp) with every extent rounded up to a multiple of the clamped window (get_window_size,
    swin_transformer.py:71-84, then the F.pad of :222-227), or None when the volume already is one."""
    b, d, h, w = dims
    ext = []
    for x, wsz in zip((d, h, w), window):
        ws = x if x <= wsz else wsz
        ext.append(x + (-x) % ws)
    return None if tuple(ext) == (d, h, w) else (b,) + tuple(ext)


def _pad_tokens(t, dims, pdims, fill=None):
    """(B*D*H*W, C) rows -> (B*Dp*Hp*Wp, C): the volume sits at the origin, the added tokens hold ``fill`` (a (C,) vector)
    or zeros -- the reference pads the NORMALISED tokens with zeros (swin_transformer.py:227), so the padding tokens' qkv
    row is the qkv bias."""
    b, d, h, w = dims
    _, dp, hp, wp = pdims
    c = t.shape[1]
    out = t.new_zeros(b, dp, hp, wp, c) if fill is None else fill.detach().to(t.dtype).expand(b, dp, hp, wp, c).contiguous()
    out[:, :d, :h, :w] = t.view(b, d, h, w, c)
    return out.view(-1, c)


def _crop_tokens(t, pdims, dims):
    b, d, h, w = dims
    _, dp, hp, wp = pdims
    return t.view(b, dp, hp, wp, t.shape[1])[:, :d, :h, :w].reshape(-1, t.shape[1])


class SwinAttentionFn(torch.autograd.Function):
    """A token volume that is not a multiple of the window (any resolution other than the 224 x 224 of the reference
    configurations) is zero-padded the way the reference does (:222-227, :246-247): the attention kernels run on the padded
    volume -- padding tokens take part as keys with q = k = v = the qkv bias, the shift mask is that of the padded
    volume -- and the result is cropped.  The padding / cropping copies are torch ops: that path is correct, not tuned."""

    @staticmethod
    def forward(ctx, y, shortcut, wqkv, bqkv, table, wproj, bproj, dims, heads, window, shift, scale, rscale):
        _chk(y, "swin_attention")
        b, d, h, w = dims
        rows = b * d * h * w
        pdims = padded_token_dims(dims, window)
        qkv = gemm(y, wqkv, 0, bias=bqkv, want_amax=pdims is None)   # its range scales V's fp16 split in the attention kernel
        if pdims is None:
            ao, lse = wmsa3d_fwd(qkv, table, dims, heads, window, shift, scale)
            ao_full = ao
        else:
            qkv = _pad_tokens(qkv, dims, pdims, fill=bqkv)
            ao_full, lse = wmsa3d_fwd(qkv, table, pdims, heads, window, shift, scale)
            ao = _crop_tokens(ao_full, pdims, dims)
        out = gemm(ao, wproj, 0, bias=bproj, residual=shortcut, row_scale=rscale, rows_per_group=rows // b)
        ctx.save_for_backward(y, wqkv, table, wproj, qkv, ao, lse, rscale, ao_full if pdims is not None else None)
        ctx.meta = (dims, heads, window, shift, scale, bqkv is not None, bproj is not None, pdims)
        return out

    @staticmethod
    def backward(ctx, g):
        y, wqkv, table, wproj, qkv, ao, lse, rscale, ao_full = ctx.saved_tensors
        dims, heads, window, shift, scale, has_bqkv, has_bproj, pdims = ctx.meta
        b = dims[0]
        g = g.contiguous()
        rpg = g.shape[0] // b
        gs = row_scale(g, rscale, rpg) if rscale is not None else g      # gradient of the branch output
        dao = gemm(gs, wproj, 1, want_amax=True)       # its range scales dO's fp16 split in the attention backward
        if has_bproj:
            dwproj, dbproj = linear_wgrad(ao, gs, want_bias=True)
        else:
            dwproj, dbproj = linear_wgrad(ao, gs), None
        if pdims is None:
            dqkv, dtable = wmsa3d_bwd(qkv, table, ao, dao, lse, dims, heads, window, shift, scale)
            pad_bias = None
        else:
            # cropped outputs carry no gradient; the padding tokens' qkv rows are the bias, so their gradient joins d(bias)
            dqkv_p, dtable = wmsa3d_bwd(qkv, table, ao_full, _pad_tokens(dao, dims, pdims), lse, pdims, heads, window, shift,
                                        scale)
            dqkv = _crop_tokens(dqkv_p, pdims, dims)
            pad_bias = dqkv_p.sum(0) - dqkv.sum(0) if has_bqkv else None
        dy = gemm(dqkv, wqkv, 1)
        if has_bqkv:
            dwqkv, dbqkv = linear_wgrad(y, dqkv, want_bias=True)
            if pad_bias is not None:
                dbqkv = dbqkv + pad_bias
        else:
            dwqkv, dbqkv = linear_wgrad(y, dqkv), None
        return dy, g, dwqkv, dbqkv, dtable, dwproj, dbproj, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------
# K8: MLP half-block   x + DropPath(fc2(GELU(fc1(y))))
# ----------------------------------------------------------------------------------------------
class SwinMlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, shortcut, w1, b1, w2, b2, n_samples, rscale):
        _chk(y, "swin_mlp")
        rows = y.shape[0]
        # fc1's epilogue computes GELU and, from the same erf, GELU' -- saved in place of the pre-activation, so the backward's
        # data-gradient epilogue only multiplies (an erf + exp per element in a GEMM epilogue bounds the K = 96 layers)
        dgelu = torch.empty(rows, w1.shape[0], dtype=torch.float32, device=y.device)
        act = gemm(y, w1, 0, bias=b1, act=4, aux_out=dgelu, want_amax=True)
        out = gemm(act, w2, 0, bias=b2, residual=shortcut, row_scale=rscale, rows_per_group=rows // n_samples)
        ctx.save_for_backward(y, w1, w2, dgelu, act, rscale)
        ctx.meta = (n_samples, b1 is not None, b2 is not None)
        return out

    @staticmethod
    def backward(ctx, g):
        y, w1, w2, dgelu, act, rscale = ctx.saved_tensors
        n_samples, has_b1, has_b2 = ctx.meta
        g = g.contiguous()
        rpg = g.shape[0] // n_samples
        gs = row_scale(g, rscale, rpg) if rscale is not None else g
        dpre = gemm(gs, w2, 1, residual=dgelu, act=5, want_amax=True)    # (g @ W2) * GELU'(pre) in the epilogue
        if has_b2:
            dw2, db2 = linear_wgrad(act, gs, want_bias=True)
        else:
            dw2, db2 = linear_wgrad(act, gs), None
        dy = gemm(dpre, w1, 1)
        if has_b1:
            dw1, db1 = linear_wgrad(y, dpre, want_bias=True)
        else:
            dw1, db1 = linear_wgrad(y, dpre), None
        return dy, g, dw1, db1, dw2, db2, None, None


# ----------------------------------------------------------------------------------------------
# PatchMerging: gather 2x2 neighbours -> LayerNorm(4C) (+ tap) -> Linear(4C -> 2C, no bias)
# ----------------------------------------------------------------------------------------------
class PatchMergeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, nw, nb, wred, eps, dims, arena, ly):
        _chk(x, "patch_merging")
        b, d, h, w = dims
        cin = x.shape[1]
        rows = b * d * ((h + 1) // 2) * ((w + 1) // 2)
        gather = (b, d, h, w, cin)
        part = None
        if ly is not None:
            part = arena.partial_buffer(ly, rows, 4 * cin, 1, 1, x.device, ln=True)
        yn, mean, rstd = ln_fwd(x, nw, nb, eps, rows, 4 * cin, part, gather)
        out = gemm(yn, wred, 0)
        ctx.save_for_backward(x, nw, nb, wred, yn, mean, rstd)
        ctx.meta = (gather, rows, arena, ly)
        tok = ops.new_token(x) if ly is not None else None
        return out, tok

    @staticmethod
    def backward(ctx, g, gtok):
        x, nw, nb, wred, yn, mean, rstd = ctx.saved_tensors
        gather, rows, arena, ly = ctx.meta
        g = g.contiguous()
        dyn = gemm(g, wred, 1)
        dwred = linear_wgrad(yn, g)
        coef = gs = None
        if ly is not None and gtok is not None:
            coef = arena.coef_ptrs(ly)
            gs = gtok.contiguous()
        gx, dg, db = ln_bwd(dyn, x, nw, nb, mean, rstd, rows, 4 * x.shape[1], None, coef, gs, gather, tuple(x.shape))
        return gx, dg, db, dwred, None, None, None, None


# ----------------------------------------------------------------------------------------------
# PatchEmbed3D: Conv3d(kernel = stride = patch) as patchify + GEMM, then LayerNorm
# ----------------------------------------------------------------------------------------------
class PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, video, wconv, bconv, nw, nb, eps, patch):
        _chk(video, "patch_embed")
        b, cin, t, h, w = video.shape
        pt, ph, pw = patch
        if cin != 3 or pw != 4 or t % pt or h % ph or w % pw:
            raise _lib.VittaError("patch_embed: needs 3-channel clips that are multiples of the (pt, ph, 4) patch")
        video = video.contiguous()
        rows = b * (t // pt) * (h // ph) * (w // pw)
        kdim = 3 * pt * ph * pw
        patches = torch.empty(rows, kdim, dtype=torch.float32, device=video.device)
        call("vitta_patchify3d", ptr(video), b, t, h, w, pt, ph, pw, ptr(patches), stream_ptr())
        w2d = wconv.view(wconv.shape[0], kdim)
        tok = gemm(patches, w2d, 0, bias=bconv)
        if nw is None:
            ctx.save_for_backward(patches, w2d)
            ctx.meta = (False, bconv is not None, tuple(wconv.shape))
            return tok
        y, mean, rstd = ln_fwd(tok, nw, nb, eps, rows, tok.shape[1])
        ctx.save_for_backward(patches, w2d, tok, nw, nb, mean, rstd)
        ctx.meta = (True, bconv is not None, tuple(wconv.shape))
        return y

    @staticmethod
    def backward(ctx, g):
        has_norm, has_bias, wshape = ctx.meta
        g = g.contiguous()
        dg = db = None
        if has_norm:
            patches, w2d, tok, nw, nb, mean, rstd = ctx.saved_tensors
            g, dg, db = ln_bwd(g, tok, nw, nb, mean, rstd, tok.shape[0], tok.shape[1])
        else:
            patches, w2d = ctx.saved_tensors
        dw = linear_wgrad(patches, g).view(wshape)
        dbias = colsum(g) if has_bias else None
        return None, dw, dbias, dg, db, None, None


# ----------------------------------------------------------------------------------------------
# token pooling of the head
# ----------------------------------------------------------------------------------------------
class FrameMeanFn(torch.autograd.Function):
    """(frames*rows, C) -> (frames, C) mean over the rows of each frame (AdaptiveAvgPool3d((1,1,1)) of I3DHead)."""

    @staticmethod
    def forward(ctx, x, frames):
        _chk(x, "frame_mean")
        n, c = x.shape
        rows = n // frames
        out = torch.empty(frames, c, dtype=torch.float32, device=x.device)
        call("vitta_frame_mean", ptr(x), frames, rows, c, ptr(out), stream_ptr())
        ctx.meta = (frames, rows, c)
        return out

    @staticmethod
    def backward(ctx, g):
        frames, rows, c = ctx.meta
        g = g.contiguous()
        gx = torch.empty(frames * rows, c, dtype=torch.float32, device=g.device)
        call("vitta_frame_mean_bwd", ptr(g), frames, rows, c, ptr(gx), stream_ptr())
        return gx, None
