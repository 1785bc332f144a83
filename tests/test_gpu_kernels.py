"""GPU parity: every kernel of the C-ABI (through the Python operators that bind it) against the CPU oracle and the
golden vectors produced by the unmodified reference.  fp32; tolerance rtol 1e-4 (BASELINE.json north_star) plus an
absolute floor scaled to the data."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import cases
from oracle import vitta_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture(scope="module")
def units():
    return cases.load_golden("units")


@pytest.fixture(autouse=True)
def _fp32(cuda_device):
    import vitta_b200
    vitta_b200.set_fp32_exact()
    from vitta_b200.utils import norm_stats_utils as nsu
    nsu.reset_arenas()
    nsu.set_process_group(None)


def _identity_bn(mod):
    """The unit protocols below hand the hook `output == input`.  For an eval-mode BatchNorm the hook recomputes the
    output from the module input in backward (in-place-ReLU safety), so the module must really be the identity:
    running_var = 1 - eps makes weight * rsqrt(running_var + eps) == 1 (a default BN scales by 1/sqrt(1 + 1e-5))."""
    if isinstance(mod, (nn.BatchNorm2d, nn.BatchNorm3d)):
        with torch.no_grad():
            mod.running_var.fill_(1.0 - mod.eps)
    return mod


def _mod(kind, c):
    return _identity_bn({"bn2d": nn.BatchNorm2d(c), "bn3d": nn.BatchNorm3d(c), "ln": nn.LayerNorm(c)}[kind])


@pytest.mark.parametrize("kind", ["bn2d", "bn3d", "ln"])
@pytest.mark.parametrize("reg", ["l1_loss", "mse_loss", "kld"])
@pytest.mark.parametrize("mavg", [1, 0])
def test_align_hook_vs_reference_golden(units, cuda_device, kind, reg, mavg):
    """Generic hook path (K1 + K2 + K3) on the exact tensors the reference hook saw; r_feature, EMA and the
    gradient w.r.t. the feature must match the reference's autograd."""
    from vitta_b200.utils.norm_stats_utils import CombineNormStatsRegHook_onereg
    key = "hook/%s/%s/%d" % (kind, reg, mavg)
    mod = _mod(kind, 6).to(cuda_device).eval()
    hook = CombineNormStatsRegHook_onereg(
        mod, clip_len=4, spatiotemp_stats_clean_tuple=(units[key + "/src_mean"], units[key + "/src_var"]),
        reg_type=reg, moving_avg=bool(mavg), momentum=0.1 if reg != "kld" else 0.9, stat_type_list=["spatiotemp"],
        reduce_dim=True, before_norm=False, if_sample_tta_aug_views=True, n_augmented_views=2)
    for s in range(3):
        feat = torch.from_numpy(units["%s/s%d/feat" % (key, s)]).to(cuda_device).requires_grad_(True)
        out = feat * 1.0
        hook.hook_fn(mod, (out,), out)   # same protocol as make_golden: the "layer output" is feat*1
        r = hook.r_feature
        r.backward()
        cases.assert_close(r.detach().cpu(), units["%s/s%d/r" % (key, s)], RTOL, 1e-6, key + "/r")
        cases.assert_close(hook.ema_mean.cpu(), units["%s/s%d/ema_mean" % (key, s)], RTOL, 1e-6, "ema_mean")
        cases.assert_close(hook.ema_var.cpu(), units["%s/s%d/ema_var" % (key, s)], RTOL, 1e-6, "ema_var")
        g = units["%s/s%d/grad" % (key, s)]
        cases.assert_close(feat.grad.cpu(), g, RTOL, 1e-6 * float(np.abs(g).max()), "grad")


@pytest.mark.parametrize("kind", ["bn2d", "bn3d", "ln"])
@pytest.mark.parametrize("st", ["spatiotemp", "temp", "temp_v2", "spatial"])
def test_stat_hook_vs_reference_golden(units, cuda_device, kind, st):
    from vitta_b200.utils.norm_stats_utils import ComputeNormStatsHook
    key = "stat/%s/%s" % (kind, st)
    mod = _mod(kind, 6).to(cuda_device).eval()
    hook = ComputeNormStatsHook(mod, clip_len=4, stat_type=st, before_norm=False, batch_size=2)
    feat = torch.from_numpy(units[key + "/feat"]).to(cuda_device)
    with torch.no_grad():
        hook.hook_fn(mod, (feat,), feat)
    cases.assert_close(hook.batch_mean.cpu(), units[key + "/mean"], RTOL, 1e-6, key + "/mean")
    cases.assert_close(hook.batch_var.cpu(), units[key + "/var"], RTOL, 1e-6, key + "/var")


@pytest.mark.parametrize("running", [1, 0])
def test_bns_hook_vs_reference_golden(units, cuda_device, running):
    from vitta_b200.utils.BNS_utils import BNFeatureHook
    key = "bns/%d" % running
    mod = nn.BatchNorm2d(6).to(cuda_device).eval()
    mod.running_mean.copy_(torch.from_numpy(units[key + "/running_mean"]))
    mod.running_var.copy_(torch.from_numpy(units[key + "/running_var"]))
    hook = BNFeatureHook(mod, reg_type="l1_loss", running_manner=bool(running), use_src_stat_in_reg=True, momentum=0.1)
    for s in range(2):
        x = torch.from_numpy(units["%s/s%d/x" % (key, s)]).to(cuda_device).requires_grad_(True)
        xin = x * 1.0
        hook.hook_fn(mod, (xin,), None)
        r = hook.r_feature
        r.backward()
        cases.assert_close(r.detach().cpu(), units["%s/s%d/r" % (key, s)], RTOL, 1e-6, key)
        g = units["%s/s%d/grad" % (key, s)]
        cases.assert_close(x.grad.cpu(), g, RTOL, 1e-6 * float(np.abs(g).max()), key)


def test_bns_hook_live_running_statistics_target(cuda_device):
    """use_src_stat_in_reg=False (utils/BNS_utils.py:61-62): the target is the layer's running statistics at hook time,
    which move when the BN layer runs in train mode.  Checked against the same arithmetic in float64."""
    from vitta_b200.utils.BNS_utils import BNFeatureHook
    g = torch.Generator().manual_seed(4)
    mod = nn.BatchNorm2d(8).to(cuda_device).train()
    mod.running_mean.copy_(torch.randn(8, generator=g))
    mod.running_var.copy_(torch.rand(8, generator=g) + 0.5)
    hook = BNFeatureHook(mod, reg_type="l1_loss", running_manner=True, use_src_stat_in_reg=False, momentum=0.1)
    ema_m = torch.zeros(8, dtype=torch.float64)
    ema_v = torch.zeros(8, dtype=torch.float64)
    for s in range(2):
        x = (torch.randn(6, 8, 5, 5, generator=g) * 2 + 1).to(cuda_device)
        mod(x)                                         # updates running_mean / running_var, then fires the hook
        xd = x.double().cpu()
        bm, bv = xd.mean((0, 2, 3)), xd.permute(1, 0, 2, 3).reshape(8, -1).var(1, unbiased=False)
        ema_m, ema_v = 0.1 * bm + 0.9 * ema_m, 0.1 * bv + 0.9 * ema_v
        want = ((mod.running_var.double().cpu() - ema_v).abs().mean()
                + (mod.running_mean.double().cpu() - ema_m).abs().mean())
        cases.assert_close(hook.r_feature.detach().cpu(), want.float().numpy(), RTOL, 1e-6, "live target step %d" % s)


@pytest.mark.parametrize("shape", ["2_2_101", "3_4_17", "1_2_400"])
def test_pred_consis_vs_reference_golden(units, cuda_device, shape):
    from vitta_b200.utils.pred_consistency_utils import compute_pred_consis
    key = "consis/" + shape
    p = torch.from_numpy(units[key + "/preds"]).to(cuda_device).requires_grad_(True)
    loss = compute_pred_consis(p)
    (loss * 1.0).backward()
    cases.assert_close(loss.detach().cpu(), units[key + "/loss"], RTOL, 1e-6, key)
    g = units[key + "/grad"]
    cases.assert_close(p.grad.cpu(), g, 2e-4, 2e-6 * float(np.abs(g).max()), key)


@pytest.mark.parametrize("shape", ["2_8_16_5_7", "1_16_32_7_7"])
def test_tam_vs_reference_golden(units, cuda_device, shape):
    """TAM module of ours (K5 stencil + torch G/L) vs the reference module's forward/backward."""
    from vitta_b200 import synth
    from vitta_b200.models.tanet_models.temporal_module import TAM
    key = "tam/" + shape
    n, t, c, h, w = (int(v) for v in shape.split("_"))
    tam = TAM(c, t)
    tam.load_state_dict(synth.synth_state_dict(tam.state_dict(), seed=5))
    tam = tam.to(cuda_device).train()
    for m in tam.modules():
        if isinstance(m, nn.BatchNorm1d):
            m.eval()
    x = torch.from_numpy(units[key + "/x"]).to(cuda_device).requires_grad_(True)
    y = tam(x)
    cases.assert_close(y.detach().cpu(), units[key + "/y"], RTOL, 1e-6, key + "/y")
    y.backward(torch.from_numpy(units[key + "/go"]).to(cuda_device))
    gx = units[key + "/gx"]
    cases.assert_close(x.grad.cpu(), gx, RTOL, 2e-6 * float(np.abs(gx).max()), key + "/gx")
    for pn, p in tam.named_parameters():
        gp = units["%s/gp/%s" % (key, pn)]
        cases.assert_close(p.grad.cpu(), gp, 3e-4, 1e-5 * float(np.abs(gp).max()) + 1e-7, pn)


# ------------------------------------------------------------------------------------------------
# K1/K2 at real layer shapes, both layouts, against the oracle (torch-CPU restatement)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape,layout", [
    ((16, 256, 14, 14), "nchw"), ((16, 256, 14, 14), "nhwc"), ((16, 2048, 7, 7), "nchw"), ((16, 2048, 7, 7), "nhwc"),
    ((32, 64, 9, 5), "nhwc"), ((8, 24, 3, 3), "nhwc"), ((8, 6, 3, 3), "nchw"), ((16, 1024, 1, 1), "nchw"),
])
def test_stats_kernels_vs_oracle(cuda_device, shape, layout):
    from vitta_b200.utils.norm_stats_utils import CombineNormStatsRegHook_onereg
    g = torch.Generator().manual_seed(1)
    f, c, h, w = shape
    T = 8
    mod = _identity_bn(nn.BatchNorm2d(c)).to(cuda_device).eval()
    src_m = torch.randn(c, generator=g) * 0.3
    src_v = torch.rand(c, generator=g) + 0.5
    hook = CombineNormStatsRegHook_onereg(mod, clip_len=T, spatiotemp_stats_clean_tuple=(src_m.numpy(), src_v.numpy()),
                                          reg_type="l1_loss", moving_avg=True, momentum=0.1,
                                          stat_type_list=["spatiotemp"], before_norm=False,
                                          if_sample_tta_aug_views=True, n_augmented_views=1)
    tap = O.AlignTap("bn2d", T, src_m, src_v, "l1_loss", True, 0.1)
    for s in range(2):
        feat = (torch.randn(shape, generator=g) * (1.0 + s) + 3.0 * torch.randn(1, c, 1, 1, generator=g))
        fo = feat.clone().requires_grad_(True)
        tap(None, fo * 1.0)
        tap.r_feature.backward()
        fg = feat.to(cuda_device)
        if layout == "nhwc":
            fg = fg.contiguous(memory_format=torch.channels_last)
        fg.requires_grad_(True)
        out = fg * 1.0
        hook.hook_fn(mod, (out,), out)
        r = hook.r_feature
        r.backward()
        cases.assert_close(hook.batch_mean.cpu(), tap.batch_mean.detach(), RTOL, 1e-5, "mean")
        cases.assert_close(hook.batch_var.cpu(), tap.batch_var.detach(), RTOL, 1e-6, "var")
        cases.assert_close(r.detach().cpu(), tap.r_feature.detach(), RTOL, 1e-6, "r")
        gmax = float(fo.grad.abs().max())
        cases.assert_close(fg.grad.cpu(), fo.grad, RTOL, 2e-6 * gmax, "grad")


def test_generic_hook_survives_inplace_relu(cuda_device):
    """torchvision-style Bottleneck: BN output overwritten by ReLU(inplace=True) right after the hook."""
    from vitta_b200.utils.norm_stats_utils import CombineNormStatsRegHook_onereg
    g = torch.Generator().manual_seed(3)
    c, T = 32, 4
    conv = nn.Conv2d(8, c, 3, padding=1, bias=False)
    bn = nn.BatchNorm2d(c)
    with torch.no_grad():
        bn.running_mean.normal_(0, 0.2, generator=g)
        bn.running_var.uniform_(0.5, 2.0, generator=g)
        bn.weight.uniform_(0.5, 1.5, generator=g)
        bn.bias.normal_(0, 0.2, generator=g)
    net_cpu = nn.Sequential(conv, bn, nn.ReLU(inplace=True)).eval()
    import copy
    net_gpu = copy.deepcopy(net_cpu).to(cuda_device)
    src_m, src_v = torch.randn(c, generator=g) * 0.2, torch.rand(c, generator=g) + 0.5
    hook = CombineNormStatsRegHook_onereg(net_gpu[1], clip_len=T, spatiotemp_stats_clean_tuple=(src_m.numpy(), src_v.numpy()),
                                          reg_type="mse_loss", moving_avg=True, momentum=0.1,
                                          stat_type_list=["spatiotemp"], before_norm=False)
    tap = O.AlignTap("bn2d", T, src_m, src_v, "mse_loss", True, 0.1)
    x = torch.randn(8, 8, 10, 10, generator=g)
    # oracle
    y = net_cpu[1](net_cpu[0](x))
    tap(None, y)
    z = F.relu(y)
    (tap.r_feature + z.mean()).backward()
    # ours
    zz = net_gpu(x.to(cuda_device))
    (hook.r_feature + zz.mean()).backward()
    for (n1, p1), (n2, p2) in zip(net_cpu.named_parameters(), net_gpu.named_parameters()):
        cases.assert_close(p2.grad.cpu(), p1.grad, 2e-4, 1e-6 * float(p1.grad.abs().max()), n1)


# ------------------------------------------------------------------------------------------------
# K4 fused BN/act/stats/pool fwd+bwd vs the oracle ops
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("res_mode", ["none", "raw", "bn"])
@pytest.mark.parametrize("shape", [(16, 64, 12, 12), (8, 256, 7, 7), (8, 24, 5, 3)])
@pytest.mark.parametrize("pool", [False, True])
def test_bn_act_vs_oracle(cuda_device, res_mode, shape, pool):
    from vitta_b200.nn import StatsBatchNorm2d, norm_act
    from vitta_b200.utils.norm_stats_utils import CombineNormStatsRegHook_onereg
    g = torch.Generator().manual_seed(5)
    f, c, h, w = shape
    T = 4

    def mkbn():
        bn = StatsBatchNorm2d(c)
        with torch.no_grad():
            bn.running_mean.normal_(0, 0.3, generator=g)
            bn.running_var.uniform_(0.5, 2.0, generator=g)
            bn.weight.uniform_(0.5, 1.5, generator=g)
            bn.bias.normal_(0, 0.3, generator=g)
        return bn.eval()
    import copy
    bn1, bn2 = mkbn(), mkbn()
    bn1g, bn2g = copy.deepcopy(bn1).to(cuda_device), copy.deepcopy(bn2).to(cuda_device)
    src = [(torch.randn(c, generator=g) * 0.3, torch.rand(c, generator=g) + 0.5) for _ in range(2)]
    hooks = [CombineNormStatsRegHook_onereg(m, clip_len=T, spatiotemp_stats_clean_tuple=(s[0].numpy(), s[1].numpy()),
                                            reg_type="l1_loss", moving_avg=True, momentum=0.1,
                                            stat_type_list=["spatiotemp"], before_norm=False)
             for m, s in zip((bn1g, bn2g), src)]
    taps = [O.AlignTap("bn2d", T, s[0], s[1], "l1_loss", True, 0.1) for s in src]
    for step in range(2):
        x = torch.randn(shape, generator=g) * 1.5
        r = torch.randn(shape, generator=g)
        go = torch.randn(shape, generator=g)
        gp = torch.randn(f, c, generator=g)
        # oracle
        xo, ro = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
        y = F.batch_norm(xo, bn1.running_mean, bn1.running_var, bn1.weight, bn1.bias, False, 0.0, bn1.eps)
        taps[0](None, y)
        total = taps[0].r_feature
        if res_mode == "raw":
            y = y + ro
        elif res_mode == "bn":
            y2 = F.batch_norm(ro, bn2.running_mean, bn2.running_var, bn2.weight, bn2.bias, False, 0.0, bn2.eps)
            taps[1](None, y2)
            total = total + 0.5 * taps[1].r_feature
            y = y + y2
        out = F.relu(y)
        po = out.mean((2, 3))
        for p in list(bn1.parameters()) + list(bn2.parameters()):
            p.grad = None
        (total + (out * go).sum() + ((po * gp).sum() if pool else 0.0)).backward()
        # ours
        xg = x.to(cuda_device).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        rg = r.to(cuda_device).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        for p in list(bn1g.parameters()) + list(bn2g.parameters()):
            p.grad = None
        outg, pg = norm_act(bn1g, xg, True, T, res=None if res_mode == "none" else rg,
                            res_bn=bn2g if res_mode == "bn" else None, want_pool=pool)
        tot = hooks[0].r_feature
        if res_mode == "bn":
            tot = tot + 0.5 * hooks[1].r_feature
        (tot + (outg * go.to(cuda_device)).sum() + ((pg * gp.to(cuda_device)).sum() if pool else 0.0)).backward()
        cases.assert_close(outg.detach().cpu(), out.detach(), RTOL, 1e-5, "out")
        if pool:
            cases.assert_close(pg.detach().cpu(), po.detach(), RTOL, 1e-5, "pool")
        cases.assert_close(tot.detach().cpu(), total.detach(), RTOL, 1e-6, "loss")
        cases.assert_close(xg.grad.cpu(), xo.grad, RTOL, 2e-6 * float(xo.grad.abs().max()), "gx")
        if res_mode != "none":
            cases.assert_close(rg.grad.cpu(), ro.grad, RTOL, 2e-6 * float(ro.grad.abs().max()), "gres")
        pairs = [(bn1g, bn1)] + ([(bn2g, bn2)] if res_mode == "bn" else [])
        for mg, mo in pairs:
            cases.assert_close(mg.weight.grad.cpu(), mo.weight.grad, 2e-4, 2e-5 * float(mo.weight.grad.abs().max()), "gw")
            cases.assert_close(mg.bias.grad.cpu(), mo.bias.grad, 2e-4, 2e-5 * float(mo.bias.grad.abs().max()), "gb")


def test_fused_sgd_vs_torch(cuda_device):
    from vitta_b200.ops import FusedSGD
    g = torch.Generator().manual_seed(9)
    shapes = [(64, 3, 7, 7), (5,), (1000, 33), (4097,), (3, 3, 3), (256, 64, 1, 1)]
    ref = [torch.randn(s, generator=g).requires_grad_(True) for s in shapes]
    ours = [p.detach().clone().to(cuda_device).requires_grad_(True) for p in ref]
    o1 = torch.optim.SGD(ref, lr=0.01, momentum=0.9, weight_decay=5e-4)
    o2 = FusedSGD(ours, lr=0.01, momentum=0.9, weight_decay=5e-4)
    for step in range(4):
        for i, (a, b) in enumerate(zip(ref, ours)):
            if i == 1 and step < 2:      # a parameter without gradient for the first steps: must be skipped
                a.grad = None
                b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g)
            a.grad = gr.clone()
            b.grad = gr.to(cuda_device)
        o1.step()
        o2.step()
        for a, b in zip(ref, ours):
            cases.assert_close(b.detach().cpu(), a.detach(), 1e-6, 1e-7, "param step %d" % step)


def test_library_fails_loudly_without_cuda_tensor(cuda_device):
    from vitta_b200 import _lib
    from vitta_b200.utils.pred_consistency_utils import compute_pred_consis
    with pytest.raises(_lib.VittaError):
        compute_pred_consis(torch.randn(2, 2, 5))


@pytest.mark.parametrize("f,h,w", [(3, 64, 64), (2, 224, 224), (5, 16, 40)])
def test_stem_conv_on_tcgen05_vs_float64(cuda_device, f, h, w):
    """conv1 (64x3x7x7, stride 2, pad 3) through vitta_stem_pack + the overlapping-row TMA operand of
    vitta_stem_conv_tf32x3, against torch's float64 convolution; the weight gradient (library path) against autograd."""
    import torch.nn.functional as F
    from vitta_b200 import ops
    g = torch.Generator().manual_seed(31)
    x = torch.randn(f, 3, h, w, generator=g).to(cuda_device)
    wt = (torch.randn(64, 3, 7, 7, generator=g) / 12.0).to(cuda_device).requires_grad_(True)
    y = ops.StemConvFn.apply(x, wt)
    assert y.shape == (f, 64, h // 2, w // 2) and y.is_contiguous(memory_format=torch.channels_last)
    w64 = wt.detach().double().requires_grad_(True)
    ref = F.conv2d(x.double(), w64, None, 2, 3)
    scale = float(ref.abs().max())
    assert float((y.detach().double() - ref).abs().max()) < 2e-6 * scale * 4
    go = torch.randn(y.shape, generator=g).to(cuda_device)
    y.backward(go)
    ref.backward(go.double())
    # the weight gradient is our fp32 FFMA kernel (exact products): tight bound, relative to the largest entry
    assert float((wt.grad.double() - w64.grad).abs().max()) < 2e-5 * float(w64.grad.abs().max())


@pytest.mark.parametrize("f,c,h,w", [(2, 64, 16, 16), (3, 64, 15, 22), (4, 64, 112, 112), (2, 128, 9, 8)])
def test_bn_relu_pool_one_pass_vs_torch(cuda_device, f, c, h, w):
    """vitta_bn_relu_pool_fwd / _bwd against BatchNorm2d(eval) -> ReLU -> MaxPool2d(3, 2, 1) of torch: output, and the
    gradients w.r.t. the input and the BN affine parameters (the argmax rule is PyTorch's: first maximum in scan order)."""
    import torch.nn as nn
    import torch.nn.functional as F
    from vitta_b200 import ops
    g = torch.Generator().manual_seed(41)
    x = torch.randn(f, c, h, w, generator=g).to(cuda_device)
    bn = nn.BatchNorm2d(c).to(cuda_device).eval()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(c, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(c, generator=g) * 0.3)
        bn.running_mean.copy_(torch.randn(c, generator=g) * 0.2)
        bn.running_var.copy_(torch.rand(c, generator=g) + 0.5)
    x1 = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    out = ops.BnReluPoolFn.apply(x1, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    x2 = x.clone().requires_grad_(True)
    bn2 = nn.BatchNorm2d(c).to(cuda_device).eval()
    bn2.load_state_dict(bn.state_dict())
    ref = F.max_pool2d(F.relu(bn2(x2)), 3, 2, 1)
    assert out.shape == ref.shape
    assert float((out - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    go = torch.randn(ref.shape, generator=g).to(cuda_device)
    out.backward(go)
    ref.backward(go)
    assert float((x1.grad - x2.grad).abs().max()) <= 1e-5 * float(x2.grad.abs().max())
    for a, b in ((bn.weight.grad, bn2.weight.grad), (bn.bias.grad, bn2.bias.grad)):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max()) + 1e-6


@pytest.mark.parametrize("n,t,c,hw", [(2, 8, 64, 5), (3, 16, 128, 4), (1, 16, 512, 2), (2, 6, 24, 3)])
def test_tam_gate_kernels_vs_torch_modules(cuda_device, n, t, c, hw):
    """K5b (G and L branches of the TAM in 3 + 7 launches, eval-mode BatchNorm1d) against the same TAM evaluated through
    its torch modules (a foreign forward hook forces that path): output and every gradient."""
    import copy
    from vitta_b200.models.tanet_models.temporal_module import TAM
    g = torch.Generator().manual_seed(51)
    tam = TAM(in_channels=c, n_segment=t).to(cuda_device)
    with torch.no_grad():
        for m in tam.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.copy_(torch.rand(m.num_features, generator=g) + 0.5)
                m.bias.copy_(torch.randn(m.num_features, generator=g) * 0.2)
                m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
    tam.train()
    for m in tam.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.eval()
    ref = copy.deepcopy(tam)
    ref.G[0].register_forward_hook(lambda *a: None)       # foreign hook -> module-by-module path
    assert tam._gate_fusable(torch.empty(1, device=cuda_device), t, c) and not ref._gate_fusable(torch.empty(1, device=cuda_device), t, c)
    x = torch.randn(n * t, c, hw, hw, generator=g).to(cuda_device)
    x1 = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    x2 = x.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y1, y2 = tam(x1), ref(x2)
    assert float((y1 - y2).abs().max()) <= 2e-5 * float(y2.abs().max())
    go = torch.randn(y2.shape, generator=g).to(cuda_device).contiguous(memory_format=torch.channels_last)
    y1.backward(go)
    y2.backward(go)
    assert float((x1.grad - x2.grad).abs().max()) <= 5e-5 * float(x2.grad.abs().max())
    for (k, p1), (_, p2) in zip(tam.named_parameters(), ref.named_parameters()):
        assert p1.grad is not None and p2.grad is not None, k
        err, scale = float((p1.grad - p2.grad).abs().max()), float(p2.grad.abs().max())
        assert err <= 1e-4 * scale + 1e-7, (k, err, scale)
