"""GPU parity of the Video-Swin kernels (K7 window attention, K8 GEMM epilogues, K9 LayerNorm + statistics) against the
oracle / float64 torch, and of the whole Swin adaptation step against the golden vectors recorded from the unmodified
reference's ``tta_standard``."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import cases

pytestmark = pytest.mark.gpu


def _rnd(*shape, seed=0, scale=1.0, dev="cuda"):
    g = np.random.Generator(np.random.PCG64(seed))
    return torch.from_numpy((g.normal(size=shape) * scale).astype(np.float32)).to(dev)


# ----------------------------------------------------------------------------------------------
# K9 LayerNorm
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rows,c", [(300, 32), (1000, 96), (1003, 128), (2049, 192), (515, 256), (777, 512), (260, 1024),
                                    (130, 1536), (64, 2048)])
def test_layernorm_fwd_bwd_vs_float64(cuda_device, rows, c):
    from vitta_b200 import ops_swin
    x = _rnd(rows, c, seed=1, scale=1.7) + 0.3
    w = _rnd(c, seed=2) * 0.2 + 1.0
    b = _rnd(c, seed=3) * 0.1
    gy = _rnd(rows, c, seed=4)
    gadd = _rnd(rows, c, seed=5)
    y, mean, rstd = ops_swin.ln_fwd(x, w, b, 1e-5, rows, c)
    xd = x.double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.layer_norm(xd, (c,), wd, bd, 1e-5)
    cases.assert_close(y.cpu(), yr.detach().cpu(), 1e-5, 1e-5, "ln y")
    yr.backward(gy.double())
    gx, dg, db = ops_swin.ln_bwd(gy, x, w, b, mean, rstd, rows, c, gadd=gadd)
    cases.assert_close(gx.cpu(), (xd.grad + gadd.double()).cpu(), 1e-4, 2e-5, "ln gx")
    cases.assert_close(dg.cpu(), wd.grad.cpu(), 1e-4, 1e-4, "ln dgamma")
    cases.assert_close(db.cpu(), bd.grad.cpu(), 1e-4, 1e-4, "ln dbeta")


@pytest.mark.parametrize("rows,c", [(5000, 64), (5003, 96), (4097, 192), (12544, 512), (3136, 1024)])
def test_layernorm_fused_statistics_and_hook_gradient(cuda_device, rows, c):
    """LN forward emits the hook's per-channel statistics; LN backward adds the closed-form hook gradient."""
    from vitta_b200 import ops, ops_swin
    x = _rnd(rows, c, seed=1, scale=1.3) + 0.2
    w = _rnd(c, seed=2) * 0.2 + 1.0
    b = _rnd(c, seed=3) * 0.1
    src_m, src_v = _rnd(c, seed=6).cpu().numpy() * 0.1, np.abs(_rnd(c, seed=7).cpu().numpy()) + 0.5
    arena = ops.StatsArena()
    ly = arena.add_layer(c, src_m, src_v, "l1_loss", True, 0.1)
    ly.n_batch = 2
    xg = x.clone().requires_grad_(True)
    wg, bg = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    y, _ = ops_swin.layer_norm_rows(xg, wg, bg, 1e-5, arena, ly, want_alias=False)
    loss = arena.layer_loss(ly) + (y * 0.0).sum()
    loss.backward()
    # reference: torch fp64
    xd = x.double().requires_grad_(True)
    wd, bd = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.layer_norm(xd, (c,), wd, bd, 1e-5)
    m, v = yr.mean(0), yr.var(0, unbiased=False)
    em, ev = 0.1 * m, 0.1 * v
    lr = (torch.from_numpy(src_v).double().to(x.device) - ev).abs().mean() + \
         (torch.from_numpy(src_m).double().to(x.device) - em).abs().mean()
    lr.backward()
    cases.assert_close(arena.vec(arena.batch_mean, ly).cpu(), m.detach().cpu(), 1e-4, 1e-5, "batch mean")
    cases.assert_close(arena.vec(arena.batch_var, ly).cpu(), v.detach().cpu(), 1e-4, 1e-6, "batch var")
    cases.assert_close(float(loss), float(lr), 1e-5, 1e-6, "loss")
    cases.assert_close(xg.grad.cpu(), xd.grad.cpu(), 1e-3, 1e-9, "hook grad through LN")
    cases.assert_close(wg.grad.cpu(), wd.grad.cpu(), 1e-3, 1e-7, "dgamma")


@pytest.mark.parametrize("b,d,h,w,cin", [(2, 2, 8, 6, 32), (1, 3, 7, 5, 64)])
def test_patch_merging_gather_norm(cuda_device, b, d, h, w, cin):
    from vitta_b200 import ops_swin
    from oracle import vitta_oracle as O  # noqa: F401  (checker only)
    x = _rnd(b, d, h, w, cin, seed=1)
    nw = _rnd(4 * cin, seed=2) * 0.2 + 1.0
    nb = _rnd(4 * cin, seed=3) * 0.1
    wred = _rnd(2 * cin, 4 * cin, seed=4) * 0.1
    xg = x.reshape(-1, cin).clone().requires_grad_(True)
    nwg, nbg, wg = (t.clone().requires_grad_(True) for t in (nw, nb, wred))
    out, _ = ops_swin.PatchMergeFn.apply(xg, nwg, nbg, wg, 1e-5, (b, d, h, w), None, None)
    go = _rnd(*out.shape, seed=5)
    out.backward(go)
    # reference: swin_transformer.py:293-312
    xd = x.double().requires_grad_(True)
    xp = F.pad(xd, (0, 0, 0, w % 2, 0, h % 2))
    cat = torch.cat([xp[:, :, 0::2, 0::2, :], xp[:, :, 1::2, 0::2, :], xp[:, :, 0::2, 1::2, :], xp[:, :, 1::2, 1::2, :]], -1)
    nwd, nbd, wd = (t.double().requires_grad_(True) for t in (nw, nb, wred))
    ref = F.linear(F.layer_norm(cat, (4 * cin,), nwd, nbd, 1e-5), wd)
    ref.backward(go.double().view(ref.shape))
    cases.assert_close(out.detach().cpu(), ref.detach().reshape(out.shape).cpu(), 1e-4, 1e-5, "merge out")
    cases.assert_close(xg.grad.cpu(), xd.grad.reshape(-1, cin).cpu(), 1e-4, 1e-5, "merge gx")
    cases.assert_close(wg.grad.cpu(), wd.grad.cpu(), 1e-4, 1e-4, "merge dW")
    cases.assert_close(nwg.grad.cpu(), nwd.grad.cpu(), 1e-4, 1e-4, "merge dgamma")


# ----------------------------------------------------------------------------------------------
# K8 epilogues / helpers
# ----------------------------------------------------------------------------------------------
def test_gemm_extended_epilogue(cuda_device):
    from vitta_b200 import ops_swin
    m, n, k = 1000, 192, 96
    a, w, bias = _rnd(m, k, seed=1), _rnd(n, k, seed=2) * 0.2, _rnd(n, seed=3)
    res = _rnd(m, n, seed=4)
    rs = torch.tensor([0.0, 1.25, 1.25, 0.0], device="cuda")
    pre = torch.empty(m, n, device="cuda")
    out = ops_swin.gemm(a, w, 0, bias=bias, residual=res, act=1, aux_out=pre, row_scale=rs, rows_per_group=250)
    pr = a.double() @ w.double().t() + bias.double()
    ref = F.gelu(pr) * rs.double().repeat_interleave(250)[:, None] + res.double()
    cases.assert_close(pre.cpu(), pr.cpu(), 1e-5, 1e-5, "pre-activation")
    cases.assert_close(out.cpu(), ref.cpu(), 1e-5, 1e-5, "gelu+scale+residual")
    # data gradient with GELU' epilogue: (g @ W) * gelu'(pre)
    g = _rnd(m, n, seed=5)
    prek = _rnd(m, k, seed=6)
    dx = ops_swin.gemm(g, w, 1, residual=prek, act=2)
    pk = prek.double().requires_grad_(True)
    F.gelu(pk).sum().backward()
    cases.assert_close(dx.cpu(), ((g.double() @ w.double()) * pk.grad).cpu(), 1e-5, 1e-5, "dgelu epilogue")
    # GELU with its derivative from one erf (act 4: aux_out = GELU'(pre)), and the multiplying backward epilogue (act 5)
    dg = torch.empty(m, n, device="cuda")
    out4 = ops_swin.gemm(a, w, 0, bias=bias, residual=res, act=4, aux_out=dg, row_scale=rs, rows_per_group=250)
    prg = pr.clone().requires_grad_(True)
    F.gelu(prg).sum().backward()
    cases.assert_close(out4.cpu(), ref.cpu(), 1e-5, 1e-5, "gelu (act 4)")
    cases.assert_close(dg.cpu(), prg.grad.cpu(), 1e-5, 1e-5, "gelu derivative (act 4)")
    dgk = _rnd(m, k, seed=7)
    dx5 = ops_swin.gemm(g, w, 1, residual=dgk, act=5)
    cases.assert_close(dx5.cpu(), ((g.double() @ w.double()) * dgk.double()).cpu(), 1e-5, 1e-5, "operand-multiply epilogue (act 5)")
    # weight gradient through the conv wgrad kernel and bias gradient
    dw = ops_swin.linear_wgrad(a, g)
    cases.assert_close(dw.cpu(), (g.double().t() @ a.double()).cpu(), 1e-5, 1e-4, "linear wgrad")
    cases.assert_close(ops_swin.colsum(g).cpu(), g.double().sum(0).cpu(), 1e-5, 1e-4, "colsum")
    cases.assert_close(ops_swin.row_scale(g, rs, 250).cpu(), (g.double() * rs.double().repeat_interleave(250)[:, None]).cpu(),
                       1e-6, 1e-7, "row scale")


def test_patch_embed_vs_conv3d(cuda_device):
    from vitta_b200 import ops_swin
    v = _rnd(2, 3, 4, 16, 24, seed=1)
    wc, bc = _rnd(32, 3, 2, 4, 4, seed=2) * 0.1, _rnd(32, seed=3) * 0.1
    nw, nb = _rnd(32, seed=4) * 0.2 + 1.0, _rnd(32, seed=5) * 0.1
    wcg, bcg, nwg, nbg = (t.clone().requires_grad_(True) for t in (wc, bc, nw, nb))
    y = ops_swin.PatchEmbedFn.apply(v, wcg, bcg, nwg, nbg, 1e-5, (2, 4, 4))
    go = _rnd(*y.shape, seed=6)
    y.backward(go)
    wd, bd, nwd, nbd = (t.double().requires_grad_(True) for t in (wc, bc, nw, nb))
    t = F.conv3d(v.double(), wd, bd, stride=(2, 4, 4)).flatten(2).transpose(1, 2)
    ref = F.layer_norm(t, (32,), nwd, nbd, 1e-5).reshape(-1, 32)
    ref.backward(go.double())
    cases.assert_close(y.detach().cpu(), ref.detach().cpu(), 1e-4, 1e-5, "patch embed")
    cases.assert_close(wcg.grad.cpu(), wd.grad.cpu(), 1e-4, 1e-4, "patch embed dW")
    cases.assert_close(bcg.grad.cpu(), bd.grad.cpu(), 1e-4, 1e-4, "patch embed db")


# ----------------------------------------------------------------------------------------------
# K7 window attention vs the reference formulation (roll -> partition -> attention -> reverse -> roll)
# ----------------------------------------------------------------------------------------------
def _ref_attention(qkv, table, dims, heads, window, shift, scale):
    """float64 restatement of swin_transformer.py:229-248 + :145-166 using the oracle's helpers."""
    from oracle import vitta_oracle as O
    b, d, h, w = dims
    c = heads * 32
    x = qkv.view(b, d, h, w, 3 * c)
    ws, ss = O.swin_window_and_shift((d, h, w), window, shift)
    mask = None
    if any(s > 0 for s in ss):
        x = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        mask = O.swin_attn_mask(d, h, w, ws, ss).to(x.device).double()
    xw = O.swin_partition(x, ws)
    b_, n, _ = xw.shape
    q3 = xw.reshape(b_, n, 3, heads, 32).permute(2, 0, 3, 1, 4)
    q, k, v = q3[0] * scale, q3[1], q3[2]
    attn = q @ k.transpose(-2, -1)
    ridx = O.swin_rel_index(window).to(x.device)
    bias = table[ridx[:n, :n].reshape(-1)].reshape(n, n, -1)
    attn = attn + bias.permute(2, 0, 1).contiguous().unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.view(b_ // nw, nw, heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    attn = F.softmax(attn, -1)
    o = (attn @ v).transpose(1, 2).reshape(b_, n, c)
    o = O.swin_reverse(o.view(-1, *(ws + (c,))), ws, b, d, h, w)
    if any(s > 0 for s in ss):
        o = torch.roll(o, shifts=ss, dims=(1, 2, 3))
    return o.reshape(-1, c)


ATTN_CASES = [
    # b, d, h, w, heads, window, shift
    (2, 8, 14, 14, 2, (8, 7, 7), (0, 0, 0)),
    (2, 8, 14, 14, 2, (8, 7, 7), (4, 3, 3)),
    (1, 16, 7, 7, 4, (8, 7, 7), (4, 3, 3)),      # temporal-only shift (H, W clamp the shift to 0)
    (2, 4, 7, 7, 1, (8, 7, 7), (4, 3, 3)),       # clamped window (4,7,7): N = 196, relative_position_index[:N,:N] quirk
    (3, 8, 4, 4, 2, (8, 7, 7), (0, 0, 0)),       # clamped window (8,4,4): N = 128
    (2, 2, 4, 4, 2, (8, 7, 7), (0, 0, 0)),       # clamped window (2,4,4): N = 32 = a single 32-key chunk (one PV issuer idle)
    (5, 1, 4, 4, 1, (8, 7, 7), (0, 0, 0)),       # clamped window (1,4,4): N = 16, items alternate the chunk parity
    (2, 8, 7, 14, 3, (8, 7, 7), (4, 3, 3)),      # two windows along W only, three heads (odd number of items per CTA)
]


@pytest.mark.parametrize("b,d,h,w,heads,window,shift", ATTN_CASES)
def test_window_attention_fwd_bwd_vs_reference(cuda_device, b, d, h, w, heads, window, shift):
    from vitta_b200 import ops_swin
    c = heads * 32
    rows = b * d * h * w
    qkv = _rnd(rows, 3 * c, seed=1, scale=1.2)
    nrel = (2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1)
    table = _rnd(nrel, heads, seed=2, scale=0.5)
    scale = 32 ** -0.5
    out, lse = ops_swin.wmsa3d_fwd(qkv, table, (b, d, h, w), heads, window, shift, scale)
    qd = qkv.double().requires_grad_(True)
    td = table.double().requires_grad_(True)
    ref = _ref_attention(qd, td, (b, d, h, w), heads, window, shift, scale)
    cases.assert_close(out.cpu(), ref.detach().cpu(), 2e-5, 2e-6, "attention output")
    go = _rnd(rows, c, seed=3)
    ref.backward(go.double())
    for impl in (0, 1):     # 0: tcgen05 kernels (the product path), 1: exact-fp32 FFMA2 kernel
        dqkv, dtable = ops_swin.wmsa3d_bwd(qkv, table, out, go, lse, (b, d, h, w), heads, window, shift, scale, impl)
        # fp32 sums over N = 392 terms against a float64 reference: the absolute floor scales with the tensor
        cases.assert_close(dqkv.cpu(), qd.grad.cpu(), 1e-4, 2e-5 * float(qd.grad.abs().max()), "dqkv impl %d" % impl)
        cases.assert_close(dtable.cpu(), td.grad.cpu(), 1e-4, 2e-5 * float(td.grad.abs().max()), "dtable impl %d" % impl)


# ----------------------------------------------------------------------------------------------
# one whole block with live DropPath (per-sample factors in the GEMM epilogues, scaled gradients in the backward)
# ----------------------------------------------------------------------------------------------
def test_swin_block_with_drop_path_vs_float64(cuda_device):
    """SwinTransformerBlock3D (shifted) with fixed per-sample DropPath factors against a float64 restatement of
    swin_transformer.py:215-274 built from the oracle's pieces.  Covers row_scale in the proj / fc2 epilogues, the scaled
    branch gradients, the GELU' epilogue and the shortcut-gradient fusion in the LayerNorm backward."""
    from oracle import vitta_oracle as O
    from vitta_b200.models.videoswintransformer_models.swin_transformer import SwinTransformerBlock3D
    b, d, h, w, c, heads = 3, 8, 14, 14, 64, 2
    window, shift = (8, 7, 7), (4, 3, 3)
    torch.manual_seed(3)
    blk = SwinTransformerBlock3D(c, heads, window, shift, drop_path=0.3).to(cuda_device)
    with torch.no_grad():
        for p_ in blk.parameters():
            p_.copy_(torch.randn_like(p_) * (0.3 if p_.dim() == 1 else 0.08))
        blk.norm1.weight.add_(1.0)
        blk.norm2.weight.add_(1.0)
    f1 = torch.tensor([0.0, 1.0 / 0.7, 1.0 / 0.7], device=cuda_device)
    f2 = torch.tensor([1.0 / 0.7, 0.0, 1.0 / 0.7], device=cuda_device)
    seq = iter([f1, f2])
    blk._drop_factors = lambda n, dev: next(seq)          # deterministic DropPath draws
    x = _rnd(b, d, h, w, c, seed=5).requires_grad_(True)
    y = blk(x)
    go = _rnd(*y.shape, seed=6)
    y.backward(go)
    # float64 reference
    sd = {"blk." + k: v.detach().double().cpu().requires_grad_(v.dtype.is_floating_point) for k, v in blk.state_dict().items()
          if "relative_position_index" not in k}
    xd = x.detach().double().cpu().requires_grad_(True)
    facs = iter([f1.double().cpu(), f2.double().cpu()])
    gen = lambda shape, keep: next(facs).view(shape) * keep     # _drop_path multiplies by mask / keep
    ws, ss = O.swin_window_and_shift((d, h, w), window, shift)
    mask = O.swin_attn_mask(d, h, w, ws, ss).double()
    ref = O.swin_block(xd, sd, "blk", heads, window, shift, mask, O.swin_rel_index(window), None, 0.3, gen)
    ref.backward(go.double().cpu())
    cases.assert_close(y.detach().cpu(), ref.detach(), 2e-5, 2e-5, "block output")
    cases.assert_close(x.grad.cpu(), xd.grad, 1e-4, 2e-5 * float(xd.grad.abs().max()), "block input gradient")
    for k in ("attn.qkv.weight", "attn.proj.bias", "mlp.fc1.weight", "mlp.fc2.bias", "norm2.weight",
              "attn.relative_position_bias_table"):
        gp = dict(blk.named_parameters())[k].grad.cpu()
        gr = sd["blk." + k].grad
        cases.assert_close(gp, gr, 2e-4, 3e-5 * float(gr.abs().max()), "grad " + k)


@pytest.mark.parametrize("d,h,w,shift", [(8, 10, 9, (4, 3, 3)), (12, 7, 16, (0, 0, 0)), (10, 14, 14, (4, 3, 3))])
def test_swin_block_on_a_padded_token_volume_vs_float64(cuda_device, d, h, w, shift):
    """Token volumes that are not multiples of the window: the reference zero-pads the normalised tokens
    (swin_transformer.py:222-227), attends on the padded volume (padding tokens are keys with the qkv bias; the shift mask is
    the padded volume's) and crops (:246-247).  Output, input gradient and every parameter gradient -- the qkv bias
    receives the padding tokens' share -- against the float64 restatement."""
    from oracle import vitta_oracle as O
    from vitta_b200 import ops_swin
    from vitta_b200.models.videoswintransformer_models.swin_transformer import SwinTransformerBlock3D
    b, c, heads, window = 2, 64, 2, (8, 7, 7)
    assert ops_swin.padded_token_dims((b, d, h, w), window) is not None
    torch.manual_seed(11)
    blk = SwinTransformerBlock3D(c, heads, window, shift, drop_path=0.0).to(cuda_device)
    with torch.no_grad():
        for p_ in blk.parameters():
            p_.copy_(torch.randn_like(p_) * (0.3 if p_.dim() == 1 else 0.08))
        blk.norm1.weight.add_(1.0)
        blk.norm2.weight.add_(1.0)
    x = _rnd(b, d, h, w, c, seed=7).requires_grad_(True)
    y = blk(x)
    go = _rnd(*y.shape, seed=8)
    y.backward(go)
    sd = {"blk." + k: v.detach().double().cpu().requires_grad_(v.dtype.is_floating_point) for k, v in blk.state_dict().items()
          if "relative_position_index" not in k}
    xd = x.detach().double().cpu().requires_grad_(True)
    ws, ss = O.swin_window_and_shift((d, h, w), window, shift)
    dp, hp, wp = (-(-e // s_) * s_ for e, s_ in zip((d, h, w), ws))
    mask = O.swin_attn_mask(dp, hp, wp, ws, ss).double() if any(ss) else None
    ref = O.swin_block(xd, sd, "blk", heads, window, shift, mask, O.swin_rel_index(window), None, 0.0, None)
    ref.backward(go.double().cpu())
    cases.assert_close(y.detach().cpu(), ref.detach(), 2e-5, 2e-5, "block output")
    cases.assert_close(x.grad.cpu(), xd.grad, 1e-4, 2e-5 * float(xd.grad.abs().max()), "block input gradient")
    for k in ("attn.qkv.weight", "attn.qkv.bias", "attn.proj.weight", "attn.proj.bias", "norm1.weight",
              "attn.relative_position_bias_table"):
        gp = dict(blk.named_parameters())[k].grad.cpu()
        gr = sd["blk." + k].grad
        cases.assert_close(gp, gr, 2e-4, 3e-5 * float(gr.abs().max()), "grad " + k)


# ----------------------------------------------------------------------------------------------
# the whole Swin adaptation step vs the reference golden vectors
# ----------------------------------------------------------------------------------------------
def _build_swin(cfg, dev):
    from vitta_b200 import synth
    from vitta_b200.models.videoswintransformer_models.recognizer3d import Recognizer3D
    model = Recognizer3D(num_classes=cfg["K"], patch_size=(2, 4, 4), window_size=tuple(cfg["window"]), drop_path_rate=0.0,
                         embed_dim=cfg["embed_dim"], depths=cfg["depths"], num_heads=cfg["heads"])
    tmpl = cases.swin_state_template(cfg["K"], cfg["embed_dim"], cfg["depths"], cfg["heads"], tuple(cfg["window"]))
    assert list(model.state_dict().keys()) == list(tmpl.keys())
    sd0 = synth.synth_state_dict(tmpl, seed=1)
    model.load_state_dict(sd0, strict=True)
    model.cls_head.dropout.p = 0.0
    return torch.nn.DataParallel(model.to(dev), device_ids=[0]), sd0


@pytest.mark.parametrize("name", list(cases.SWIN_CASES))
def test_swin_tta_vs_reference_golden(cuda_device, name):
    _run_swin_case(cuda_device, name, cases.SWIN_CASES[name])


@pytest.mark.parametrize("name", list(cases.SWIN_OPTION_CASES))
def test_swin_option_modes_vs_reference_golden(cuda_device, name):
    """--update_only_bn_affine on Video-Swin (Adam over the LayerNorm affine parameters, everything else frozen); MSE
    alignment with AverageMeterTensor statistics (moving_avg=False)."""
    _run_swin_case(cuda_device, name, cases.SWIN_OPTION_CASES[name])


def _run_swin_case(cuda_device, name, cfg):
    import vitta_b200
    from vitta_b200.corpus.basics import OnlineAdapter, compute_statistics
    from vitta_b200.utils import norm_stats_utils as nsu
    from vitta_b200.utils.opts import default_args
    vitta_b200.set_fp32_exact()
    nsu.reset_arenas()
    dev = cuda_device
    g = cases.load_golden(name)
    src_m, src_v = cases.src_stats_from_golden(g)
    model, sd0 = _build_swin(cfg, dev)
    common = dict(arch='videoswintransformer', clip_length=cfg["T"], batch_size=cfg["N"], num_classes=cfg["K"],
                  input_size=cfg["res"], window_size=tuple(cfg["window"]), patch_size=(2, 4, 4), num_clips=1)
    # source statistics through our compute_statistics (fused LN statistics, eval forward)
    clean = cases.case_inputs(cfg, "swin", "clean", 2, 100)

    class DS(torch.utils.data.Dataset):
        def __init__(self, x):
            self.x = x

        def __len__(self):
            return self.x.shape[0]

        def __getitem__(self, i):
            return self.x[i], 0
    a2 = default_args(stat_type='spatiotemp', result_dir=None, **common)
    a2.dataset_factory = lambda a, split, dataset_type: DS(torch.cat(clean, 0))
    om, ov = compute_statistics(model, a2)
    assert len(om) == len(src_m)
    for i in range(len(om)):
        cases.assert_close(om[i], src_m[i], 2e-4, 2e-5 * float(np.abs(src_m[i]).max()) + 1e-6, "src_mean/%d" % i)
        cases.assert_close(ov[i], src_v[i], 2e-4, 2e-5 * float(np.abs(src_v[i]).max()) + 1e-6, "src_var/%d" % i)

    args = default_args(n_augmented_views=cfg["M"], if_pred_consistency=cfg["consis"], reg_type=cfg["reg_type"],
                        lr=cfg["lr"], moving_avg=cfg.get("moving_avg", True), momentum_mvg=cfg["momentum_mvg"],
                        lambda_pred_consis=cfg["lambda_consis"], chosen_blocks=cfg["chosen"],
                        if_sample_tta_aug_views=cfg.get("sample_views", True),
                        update_only_bn_affine=cfg.get("bn_affine", False), **common)
    ad = OnlineAdapter(model, args, (src_m, src_v))
    assert len(ad.stat_reg_hooks) == int(g["n_hooks"])
    tta_in, eval_in = cases.tta_inputs(cfg, "swin")
    for s in range(cfg["steps"]):
        r = ad.adapt(tta_in[s].to(dev))
        cases.assert_close(r["loss_reg"].cpu(), g["step%d/loss_reg" % s], 1e-4, 1e-6, "loss_reg step %d" % s)
        if cfg["consis"]:
            cases.assert_close(r["loss_consis"].cpu(), g["step%d/loss_consis" % s], 1e-3, 1e-7, "loss_consis")
        for h, hook in enumerate(ad.stat_reg_hooks):
            cases.assert_close(hook.r_feature.detach().cpu(), g["step%d/r_feature/%d" % (s, h)], 1e-4, 1e-6,
                               "r_feature %d" % h)
            em, ev = g["step%d/ema_mean/%d" % (s, h)], g["step%d/ema_var/%d" % (s, h)]
            scale = float((np.abs(em) + np.sqrt(np.abs(ev))).max())
            cases.assert_close(hook.ema_mean.cpu(), em, 1e-4, 3e-5 * scale + 1e-7, "ema_mean %d" % h)
            cases.assert_close(hook.ema_var.cpu(), ev, 1e-4, 1e-5 * float(np.abs(ev).max()) + 1e-7, "ema_var %d" % h)
        ad.hooks_off()
        ev = ad.evaluate(eval_in[s].to(dev))
        want = g["step%d/eval_logits" % s]
        cases.assert_close(ev.cpu(), want, 2e-4, 2e-5 * float(np.abs(want).max()), "eval logits step %d" % s)
        ad.hooks_on()
    new_sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in ad.model.state_dict().items()}
    names = [str(n) for n in g["delta_names"]]
    ref = g["delta_norm_sum"]
    for i, n in enumerate(names):
        d = (new_sd[n].detach().cpu() - sd0[n]).double()
        cases.assert_close(float(d.norm()), ref[i, 0], 1e-2, 1e-9, "delta norm " + n)
    for k in g.files:
        if k.startswith("delta/"):
            n = k[len("delta/"):]
            d = (new_sd[n].detach().cpu() - sd0[n]).reshape(-1)[:4096]
            scale = float(np.abs(g[k]).max()) + 1e-12
            ulp = 1.2e-7 * float(sd0[n].abs().max())
            cases.assert_close(d, g[k], 5e-3, 2e-3 * scale + ulp, k)


def test_swin_state_dict_keys_match_reference_layout():
    from vitta_b200.models.videoswintransformer_models.recognizer3d import Recognizer3D
    m = Recognizer3D(num_classes=101, patch_size=(2, 4, 4), window_size=(8, 7, 7), drop_path_rate=0.2)
    tmpl = cases.swin_state_template(101, 128, [2, 2, 18, 2], [4, 8, 16, 32])
    assert list(m.state_dict().keys()) == list(tmpl.keys())
    sd = m.state_dict()
    assert all(tuple(sd[k].shape) == tuple(v.shape) for k, v in tmpl.items())
    from oracle import vitta_oracle as O
    ln_names = [n for n, mod in m.named_modules() if isinstance(mod, torch.nn.LayerNorm)]
    assert ln_names == O.swin_norm_layers()


def test_producers_emit_the_operand_range_of_their_outputs(cuda_device):
    """Under the fp16 operand split every producer of a GEMM operand on the Swin path (LayerNorm forward / backward,
    DropPath row scaling, the fc1 / GELU' GEMM epilogues, window attention forward / backward) emits max|output| into a
    pooled device scalar attached to the output tensor: it must equal the true maximum (it is an exact max of the stored
    fp32 values), so no consumer ever needs the standalone vitta_amax_f32 pass."""
    from vitta_b200 import ops, ops_swin
    before = ops.gemm_precision()
    ops.set_gemm_precision("f16x3")
    try:
        ops.reset_amax_pool()

        def check(t, what):
            ent = getattr(t, "_vitta_amax", None)
            assert ent is not None, what + ": no range attached"
            assert float(ent[0]) == float(t.abs().max()), (what, float(ent[0]), float(t.abs().max()))
        rows, c = 1500, 96
        x = _rnd(rows, c, seed=1, scale=1.7) + 0.3
        w, b = _rnd(c, seed=2) * 0.2 + 1.0, _rnd(c, seed=3) * 0.1
        y, mean, rstd = ops_swin.ln_fwd(x, w, b, 1e-5, rows, c)
        check(y, "ln_fwd")
        gx, _, _ = ops_swin.ln_bwd(_rnd(rows, c, seed=4) * 1e-6, x, w, b, mean, rstd, rows, c, gadd=_rnd(rows, c, seed=5) * 1e-6)
        check(gx, "ln_bwd")
        sc = torch.tensor([0.0, 1.25, 1.25], device=cuda_device)
        check(ops_swin.row_scale(x, sc, rows // 3), "row_scale")
        w1 = torch.nn.Parameter(_rnd(4 * c, c, seed=6) * 0.1)
        pre = torch.empty(rows, 4 * c, device=cuda_device)
        act = ops_swin.gemm(y, w1, 0, bias=_rnd(4 * c, seed=7) * 0.1, act=1, aux_out=pre, want_amax=True)
        check(act, "gemm epilogue (GELU)")
        heads, dims, window = 3, (1, 4, 7, 7), (4, 7, 7)
        qkv = _rnd(4 * 7 * 7, 3 * heads * 32, seed=8)
        table = _rnd((2 * 4 - 1) * 13 * 13, heads, seed=9) * 0.2
        out, lse = ops_swin.wmsa3d_fwd(qkv, table, dims, heads, window, (0, 0, 0), 32 ** -0.5)
        check(out, "wmsa3d_fwd")
        dqkv, _ = ops_swin.wmsa3d_bwd(qkv, table, out, _rnd(*out.shape, seed=10), lse, dims, heads, window, (0, 0, 0), 32 ** -0.5)
        check(dqkv, "wmsa3d_bwd")
    finally:
        ops.set_gemm_precision(before)


@pytest.mark.parametrize("m,k,n", [(5000, 96, 288), (1568, 384, 96), (3136, 128, 512), (100, 64, 40)])
def test_linear_wgrad_returns_the_bias_gradient(cuda_device, m, k, n):
    """The fp16-split weight-gradient kernel adds up dY over the pixels while it transposes the dY tiles: db must equal
    the column sums of gy (and dW is unchanged by the option)."""
    from vitta_b200 import ops, ops_swin
    before = ops.gemm_precision()
    ops.set_gemm_precision("f16x3")
    try:
        x, gy = _rnd(m, k, seed=1), _rnd(m, n, seed=2) * 1e-3
        gw0 = ops_swin.linear_wgrad(x, gy)
        gw, db = ops_swin.linear_wgrad(x, gy, want_bias=True)
        assert torch.equal(gw, gw0)
        want = gy.double().sum(0)
        cases.assert_close(db.cpu(), want.cpu(), 1e-5, 1e-6 * float(gy.abs().sum(0).max()), "bias gradient")
    finally:
        ops.set_gemm_precision(before)
