"""TEST INFRASTRUCTURE -- CPU restatement (the "oracle") of ViTTA's test-time-adaptation inner loop.

This file is NOT part of the product.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it, and only as the checker
(or as the timed CPU stand-in for the reference, which cannot travel to the GPU box).  The product
path (``vitta_b200``) never imports it and fails loudly when its CUDA library is missing.

It restates, in plain torch-CPU fp32 and driven by a flat ``state_dict`` (no nn.Module tree), the
algorithm of the reference wlin-at/ViTTA @ c8e01fa.  Every function cites the reference lines it
follows.  The op sequence deliberately mirrors the reference's (permute -> contiguous -> mean/var,
grouped conv2d for the TAM, materialised attention) so that timing it on host cores is a fair
stand-in for the reference's own PyTorch-CPU path.

Parity pin: ``oracle/make_golden.py`` runs the *unmodified reference* (``oracle/ref_harness.py``) in
the build container on seeded synthetic inputs and commits its outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this restatement against those vectors.  The reference ships
no tests or golden vectors of its own, so these generated vectors are the pin.
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # torch.nn.BatchNorm default used by torchvision resnet50 and the TAM (temporal_module.py:29,37)


# --------------------------------------------------------------------------------------------
# a2-a6: statistics hooks
# --------------------------------------------------------------------------------------------
def feature_to_ncthw(feature, kind, clip_len=None):
    """Reshape a hooked norm-layer output to (N, C, T, H, W).

    bn2d: utils/norm_stats_utils.py:188-193,208-214   (N*T, C, H, W) -> view -> permute(0,2,1,3,4)
    bn3d: :195-197                                     already (N, C, T, H, W)
    ln:   :222-236                                     (B, D, H, W, C) -> permute(0,4,1,2,3)
    """
    if kind == "bn2d":
        nt, c, h, w = feature.shape
        return feature.view(nt // clip_len, clip_len, c, h, w).permute(0, 2, 1, 3, 4).contiguous()
    if kind == "bn3d":
        return feature
    if kind == "ln":
        assert feature.dim() == 5
        return feature.permute(0, 4, 1, 2, 3).contiguous()
    raise ValueError(kind)


def spatiotemp_stats(x):
    """Per-channel mean and *biased* variance over (N,T,H,W): norm_stats_utils.py:242-243 (= :93-95)."""
    c = x.shape[1]
    mean = x.mean((0, 2, 3, 4))
    var = x.permute(1, 0, 2, 3, 4).contiguous().view(c, -1).var(1, unbiased=False)
    return mean, var


def other_stats(x, stat_type):
    """ComputeNormStatsHook's remaining stat types: norm_stats_utils.py:81-98."""
    n, c, t, h, w = x.shape
    if stat_type == "temp":
        return x.mean((0, 2)), x.permute(1, 3, 4, 0, 2).contiguous().view(c, h, w, -1).var(-1, unbiased=False)
    if stat_type == "temp_v2":
        y = x.mean((3, 4))
        return y.mean((0, 2)), y.permute(1, 0, 2).contiguous().view(c, -1).var(1, unbiased=False)
    if stat_type == "spatial":
        return x.mean((0, 3, 4)), x.permute(1, 2, 0, 3, 4).contiguous().view(c, t, -1).var(-1, unbiased=False)
    if stat_type == "spatiotemp":
        return spatiotemp_stats(x)
    raise ValueError(stat_type)


def kld(mean_true, mean_pred, var_true, var_pred):
    """norm_stats_utils.py:8-16."""
    v = 0.5 * torch.log(var_pred / var_true) + (var_true + (mean_true - mean_pred) ** 2) / (2 * var_pred) - 0.5
    return v.sum()


def regularization(mean_true, mean_pred, var_true, var_pred, reg_type):
    """compute_regularization, norm_stats_utils.py:531-542 (mean reduction over C for l1/mse)."""
    if reg_type == "mse_loss":
        return F.mse_loss(var_pred, var_true) + F.mse_loss(mean_pred, mean_true)
    if reg_type == "l1_loss":
        return F.l1_loss(var_pred, var_true) + F.l1_loss(mean_pred, mean_true)
    if reg_type == "kld":
        return kld(mean_true, mean_pred, var_true, var_pred)
    raise ValueError(reg_type)


class EmaMeter:
    """MovingAverageTensor, utils/utils_.py:204-211: starts at scalar 0, history detached, no bias correction."""

    def __init__(self, momentum):
        self.momentum = momentum
        self.avg = torch.tensor(0.0)

    def update(self, val, n=None):
        self.avg = self.momentum * val + (1.0 - self.momentum) * self.avg.detach()


class MeanMeter:
    """AverageMeterTensor, utils/utils_.py:190-202: weighted running mean with detached sum."""

    def __init__(self):
        self.sum = torch.tensor(0.0)
        self.count = 0
        self.avg = torch.tensor(0.0)

    def update(self, val, n=1):
        self.sum = self.sum.detach() + val * n
        self.count += n
        self.avg = self.sum / self.count


class AlignTap:
    """State + forward rule of CombineNormStatsRegHook_onereg (norm_stats_utils.py:103-258) for one layer."""

    def __init__(self, kind, clip_len, src_mean, src_var, reg_type="l1_loss", moving_avg=True, momentum=0.1,
                 before_norm=False):
        self.kind, self.clip_len, self.reg_type = kind, clip_len, reg_type
        self.before_norm = before_norm
        self.src_mean = None if src_mean is None else torch.as_tensor(src_mean, dtype=torch.float32)
        self.src_var = None if src_var is None else torch.as_tensor(src_var, dtype=torch.float32)
        self.moving_avg = moving_avg
        mk = (lambda: EmaMeter(momentum)) if moving_avg else MeanMeter
        self.mean_meter, self.var_meter = mk(), mk()
        self.r_feature = torch.tensor(0.0)
        self.batch_mean = self.batch_var = None

    def __call__(self, norm_input, norm_output):
        feature = norm_input if self.before_norm else norm_output
        self.r_feature = torch.tensor(0.0)
        if self.kind == "bn1d":
            return  # :158-183 -- with stat_type_list == ['spatiotemp'] BatchNorm1d hooks contribute nothing
        x = feature_to_ncthw(feature, self.kind, self.clip_len)
        mean, var = spatiotemp_stats(x)
        self.batch_mean, self.batch_var = mean, var
        n = x.shape[0]
        self.mean_meter.update(mean, n)
        self.var_meter.update(var, n)
        self.r_feature = self.r_feature + regularization(self.src_mean, self.mean_meter.avg, self.src_var,
                                                         self.var_meter.avg, self.reg_type)


class StatTap:
    """ComputeNormStatsHook (norm_stats_utils.py:18-101) for one layer."""

    def __init__(self, kind, clip_len, stat_type="spatiotemp", before_norm=False):
        self.kind, self.clip_len, self.stat_type, self.before_norm = kind, clip_len, stat_type, before_norm
        self.batch_mean = self.batch_var = None

    def __call__(self, norm_input, norm_output):
        feature = norm_input if self.before_norm else norm_output
        if self.kind == "bn1d":
            return
        x = feature_to_ncthw(feature, self.kind, self.clip_len)
        self.batch_mean, self.batch_var = other_stats(x, self.stat_type)


class BnsTap:
    """BNFeatureHook (utils/BNS_utils.py:19-77): statistics of the BN *input* per frame batch vs the BN
    running statistics, optional EMA from zeros."""

    def __init__(self, running_mean, running_var, reg_type="l1_loss", running_manner=True, momentum=0.1):
        self.src_mean, self.src_var = running_mean.detach().clone(), running_var.detach().clone()
        self.reg_type, self.running_manner, self.momentum = reg_type, running_manner, momentum
        self.mean = torch.zeros_like(running_mean)
        self.var = torch.zeros_like(running_var)
        self.r_feature = torch.tensor(0.0)

    def __call__(self, norm_input, norm_output):
        x = norm_input
        c = x.shape[1]
        dims = [d for d in range(x.dim()) if d != 1]
        bm = x.mean(dims)
        bv = x.transpose(0, 1).contiguous().view(c, -1).var(1, unbiased=False)
        if self.running_manner:
            self.mean = self.momentum * bm + (1 - self.momentum) * self.mean.detach()
            self.var = self.momentum * bv + (1 - self.momentum) * self.var.detach()
        else:
            self.mean, self.var = bm, bv
        self.r_feature = regularization(self.src_mean, self.mean, self.src_var, self.var, self.reg_type)


# --------------------------------------------------------------------------------------------
# a9: prediction consistency
# --------------------------------------------------------------------------------------------
def pred_consistency(preds):
    """compute_pred_consis, utils/pred_consistency_utils.py:15-31.  preds (B, V, K) logits."""
    b, v, k = preds.shape
    p = [F.softmax(preds[:, i, :], dim=1) for i in range(v)]
    avg = torch.stack(p, 0).mean(0)  # NOT detached (:24-25)
    return sum(F.l1_loss(p[i], avg, reduction="sum") for i in range(v)) / v


def accuracy(output, target, topk=(1,)):
    """utils/utils_.py:224-237."""
    maxk = max(topk)
    _, pred = output.topk(maxk, 1, True, True)
    correct = pred.t().eq(target.view(1, -1).expand(maxk, -1))
    return [correct[:k].reshape(-1).float().sum(0) * (100.0 / target.size(0)) for k in topk]


# --------------------------------------------------------------------------------------------
# a10-a13: TANet-R50 (TSN + TAM), functional over a state dict
# --------------------------------------------------------------------------------------------
RESNET50_STAGES = ((64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2))  # torchvision resnet50: width, blocks, stride


def tanet_norm_layers():
    """[(name, kind)] of every BatchNorm in ``named_modules()`` order -- the index<->layer map of the
    source-statistics .npy lists (choose_layers, utils/BNS_utils.py:245-259)."""
    out = [("base_model.bn1", "bn2d")]
    for li, (_, nblk, _) in enumerate(RESNET50_STAGES, 1):
        for b in range(nblk):
            p = "base_model.layer%d.%d" % (li, b)
            out += [(p + ".net.bn1", "bn2d"), (p + ".net.bn2", "bn2d"), (p + ".net.bn3", "bn2d")]
            if b == 0:
                out.append((p + ".net.downsample.1", "bn2d"))
            out += [(p + ".tam.G.1", "bn1d"), (p + ".tam.L.1", "bn1d")]
    return out


def _bn(x, sd, p, taps, bn_training=False):
    """BatchNorm in eval mode (fix_BNS, corpus/basics.py:606-611) followed by the layer's forward hook."""
    if bn_training:
        y = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                         True, 0.1, BN_EPS)
    else:
        y = F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                         False, 0.0, BN_EPS)
    if taps is not None and p in taps:
        taps[p](x, y)
    return y


def tam_forward(x, sd, p, t, taps=None, bn_training=False):
    """TAM.forward, models/tanet_models/temporal_module.py:43-65.  x: (N*T, C, H, W)."""
    nt, c, h, w = x.shape
    n = nt // t
    new_x = x.view(n, t, c, h, w).permute(0, 2, 1, 3, 4).contiguous()
    pooled = F.adaptive_avg_pool2d(new_x.view(n * c, t, h, w), (1, 1)).view(-1, t)
    g = F.linear(pooled, sd[p + ".G.0.weight"])
    g = F.relu(_bn(g, sd, p + ".G.1", taps, bn_training))
    kern = F.softmax(F.linear(g, sd[p + ".G.3.weight"]), -1).view(n * c, 1, -1, 1)
    a = F.conv1d(pooled.view(n, c, t), sd[p + ".L.0.weight"], None, 1, 1)
    a = F.relu(_bn(a, sd, p + ".L.1", taps, bn_training))
    act = torch.sigmoid(F.conv1d(a, sd[p + ".L.3.weight"])).view(n, c, t, 1, 1)
    new_x = new_x * act
    out = F.conv2d(new_x.view(1, n * c, t, h * w), kern, None, (1, 1), (1, 0), groups=n * c)
    return out.view(n, c, t, h, w).permute(0, 2, 1, 3, 4).contiguous().view(nt, c, h, w)


def temporal_bottleneck(x, sd, p, t, stride, has_ds, taps, bn_training=False):
    """TemporalBottleneck.forward, temporal_module.py:85-106 (torchvision Bottleneck v1.5: stride on conv2)."""
    q = p + ".net"
    out = F.conv2d(x, sd[q + ".conv1.weight"])
    out = F.relu(_bn(out, sd, q + ".bn1", taps, bn_training))
    out = tam_forward(out, sd, p + ".tam", t, taps, bn_training)
    out = F.conv2d(out, sd[q + ".conv2.weight"], None, stride, 1)
    out = F.relu(_bn(out, sd, q + ".bn2", taps, bn_training))
    out = F.conv2d(out, sd[q + ".conv3.weight"])
    out = _bn(out, sd, q + ".bn3", taps, bn_training)
    if has_ds:
        idt = F.conv2d(x, sd[q + ".downsample.0.weight"], None, stride)
        idt = _bn(idt, sd, q + ".downsample.1", taps, bn_training)
    else:
        idt = x
    return F.relu(out + idt)


def tanet_forward(sd, x, t, taps=None, dropout_p=0.0, bn_training=False):
    """TSN.forward, models/tanet_models/tanet.py:308-333.  x: (N', T, 3, H, W) -> logits (N', K).
    ``dropout_p`` > 0 reproduces the live Dropout(0.8) of the adaptation forward (basics.py:606)."""
    x = x.reshape((-1, 3) + tuple(x.shape[-2:]))
    x = F.conv2d(x, sd["base_model.conv1.weight"], None, 2, 3)
    x = F.relu(_bn(x, sd, "base_model.bn1", taps, bn_training))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, (_, nblk, stride) in enumerate(RESNET50_STAGES, 1):
        for b in range(nblk):
            x = temporal_bottleneck(x, sd, "base_model.layer%d.%d" % (li, b), t, stride if b == 0 else 1,
                                    b == 0, taps, bn_training)
    x = F.adaptive_avg_pool2d(x, 1).flatten(1)
    if dropout_p > 0:
        x = F.dropout(x, dropout_p, True)
    x = F.linear(x, sd["new_fc.weight"], sd["new_fc.bias"])
    x = x.view((-1, t) + tuple(x.shape[1:]))
    return x.mean(1)  # SegmentAvg_static, basic_ops.py:38-51


# --------------------------------------------------------------------------------------------
# a14-a16: Video Swin (Recognizer3D), functional over a state dict
# --------------------------------------------------------------------------------------------
def swin_window_and_shift(x_size, window, shift):
    """get_window_size, swin_transformer.py:71-84: clamp the window (and zero the shift) per dimension."""
    ws, ss = list(window), list(shift)
    for i in range(3):
        if x_size[i] <= window[i]:
            ws[i] = x_size[i]
            ss[i] = 0
    return tuple(ws), tuple(ss)


def swin_partition(x, ws):
    """window_partition, swin_transformer.py:38-50."""
    b, d, h, w, c = x.shape
    x = x.view(b, d // ws[0], ws[0], h // ws[1], ws[1], w // ws[2], ws[2], c)
    return x.permute(0, 1, 3, 5, 2, 4, 6, 7).contiguous().view(-1, ws[0] * ws[1] * ws[2], c)


def swin_reverse(win, ws, b, d, h, w):
    """window_reverse, swin_transformer.py:53-66."""
    x = win.view(b, d // ws[0], h // ws[1], w // ws[2], ws[0], ws[1], ws[2], -1)
    return x.permute(0, 1, 4, 2, 5, 3, 6, 7).contiguous().view(b, d, h, w, -1)


def swin_attn_mask(dp, hp, wp, ws, ss):
    """compute_mask, swin_transformer.py:316-329: 27 region ids -> 0 / -100 additive mask per window."""
    img = torch.zeros((1, dp, hp, wp, 1))
    cnt = 0
    for d in (slice(-ws[0]), slice(-ws[0], -ss[0]), slice(-ss[0], None)):
        for h in (slice(-ws[1]), slice(-ws[1], -ss[1]), slice(-ss[1], None)):
            for w in (slice(-ws[2]), slice(-ws[2], -ss[2]), slice(-ss[2], None)):
                img[:, d, h, w, :] = cnt
                cnt += 1
    mw = swin_partition(img, ws).squeeze(-1)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def swin_rel_index(window):
    """relative_position_index buffer, swin_transformer.py:113-125."""
    coords = torch.stack(torch.meshgrid(*[torch.arange(s) for s in window], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window[0] - 1
    rel[:, :, 1] += window[1] - 1
    rel[:, :, 2] += window[2] - 1
    rel[:, :, 0] *= (2 * window[1] - 1) * (2 * window[2] - 1)
    rel[:, :, 1] *= 2 * window[2] - 1
    return rel.sum(-1)


def swin_window_attention(xw, sd, p, heads, rel_index, mask):
    """WindowAttention3D.forward, swin_transformer.py:138-169.  xw: (B_, N, C)."""
    b_, n, c = xw.shape
    qkv = F.linear(xw, sd[p + ".qkv.weight"], sd[p + ".qkv.bias"]).reshape(b_, n, 3, heads, c // heads)
    qkv = qkv.permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * ((c // heads) ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[p + ".relative_position_bias_table"][rel_index[:n, :n].reshape(-1)].reshape(n, n, -1)
    attn = attn + bias.permute(2, 0, 1).contiguous().unsqueeze(0)
    if mask is not None:
        nw = mask.shape[0]
        attn = attn.view(b_ // nw, nw, heads, n, n) + mask.unsqueeze(1).unsqueeze(0)
        attn = attn.view(-1, heads, n, n)
    attn = F.softmax(attn, -1)
    x = (attn @ v).transpose(1, 2).reshape(b_, n, c)
    return F.linear(x, sd[p + ".proj.weight"], sd[p + ".proj.bias"])


def _ln(x, sd, p, taps):
    y = F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)
    if taps is not None and p in taps:
        taps[p](x, y)
    return y


def _drop_path(x, p, gen):
    """timm 0.6.7 DropPath (call site swin_transformer.py:210): per-sample Bernoulli(keep)/keep."""
    if p == 0.0 or gen is None:
        return x
    keep = 1.0 - p
    mask = gen((x.shape[0],) + (1,) * (x.dim() - 1), keep)
    return x * (mask / keep)


def swin_block(x, sd, p, heads, window, shift, mask_matrix, rel_index, taps, dp_rate=0.0, dp_gen=None):
    """SwinTransformerBlock3D.forward/forward_part1/forward_part2, swin_transformer.py:215-274."""
    b, d, h, w, c = x.shape
    ws, ss = swin_window_and_shift((d, h, w), window, shift)
    shortcut = x
    x = _ln(x, sd, p + ".norm1", taps)
    pad_d = (ws[0] - d % ws[0]) % ws[0]
    pad_b = (ws[1] - h % ws[1]) % ws[1]
    pad_r = (ws[2] - w % ws[2]) % ws[2]
    x = F.pad(x, (0, 0, 0, pad_r, 0, pad_b, 0, pad_d))
    _, dp, hp, wp, _ = x.shape
    if any(s > 0 for s in ss):
        x = torch.roll(x, shifts=(-ss[0], -ss[1], -ss[2]), dims=(1, 2, 3))
        mask = mask_matrix
    else:
        mask = None
    aw = swin_window_attention(swin_partition(x, ws), sd, p + ".attn", heads, rel_index, mask)
    x = swin_reverse(aw.view(-1, *(ws + (c,))), ws, b, dp, hp, wp)
    if any(s > 0 for s in ss):
        x = torch.roll(x, shifts=ss, dims=(1, 2, 3))
    if pad_d > 0 or pad_r > 0 or pad_b > 0:
        x = x[:, :d, :h, :w, :].contiguous()
    x = shortcut + _drop_path(x, dp_rate, dp_gen)
    y = _ln(x, sd, p + ".norm2", taps)
    y = F.linear(F.gelu(F.linear(y, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])),
                 sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
    return x + _drop_path(y, dp_rate, dp_gen)


def swin_norm_layers(depths=(2, 2, 18, 2), prefix="backbone."):
    """LayerNorm names in ``named_modules()`` order (choose_layers; the reference then drops entry 0,
    corpus/basics.py:541-543)."""
    out = [prefix + "patch_embed.norm"]
    for i, dep in enumerate(depths):
        for b in range(dep):
            out += ["%slayers.%d.blocks.%d.norm1" % (prefix, i, b), "%slayers.%d.blocks.%d.norm2" % (prefix, i, b)]
        if i < len(depths) - 1:
            out.append("%slayers.%d.downsample.norm" % (prefix, i))
    out.append(prefix + "norm")
    return out


def swin_forward(sd, x, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32), window=(8, 7, 7), patch=(2, 4, 4),
                 taps=None, drop_path_rate=0.0, dp_gen=None, dropout_p=0.0, prefix="backbone."):
    """Recognizer3D.forward (recognizer3d.py:95-115) = SwinTransformer3D.forward (swin_transformer.py:650-663)
    + I3DHead.forward (i3d_head.py:58-77).  x: (N, V, 3, T, H, W) -> (vid (N,K), view (N,V,K))."""
    n, v = x.shape[:2]
    x = x.reshape((-1,) + tuple(x.shape[2:]))
    # PatchEmbed3D, :440-456 (inputs here are always multiples of the patch, so no padding branch)
    x = F.conv3d(x, sd[prefix + "patch_embed.proj.weight"], sd[prefix + "patch_embed.proj.bias"], patch)
    bsz, c, d, h, w = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = _ln(x, sd, prefix + "patch_embed.norm", taps)
    x = x.transpose(1, 2).reshape(bsz, c, d, h, w)
    nblk = sum(depths)
    dpr = [r.item() for r in torch.linspace(0, drop_path_rate, nblk)]
    shift_full = tuple(s // 2 for s in window)
    bi = 0
    for i, dep in enumerate(depths):
        # BasicLayer.forward, :392-413
        bsz, c, d, h, w = x.shape
        ws, ss = swin_window_and_shift((d, h, w), window, shift_full)
        x = x.permute(0, 2, 3, 4, 1).contiguous()
        dp_, hp_, wp_ = (int(math.ceil(s / k)) * k for s, k in zip((d, h, w), ws))
        mask = swin_attn_mask(dp_, hp_, wp_, ws, ss) if any(s > 0 for s in ss) else None
        rel_index = swin_rel_index(window)
        for b in range(dep):
            x = swin_block(x, sd, "%slayers.%d.blocks.%d" % (prefix, i, b), heads[i], window,
                           (0, 0, 0) if b % 2 == 0 else shift_full, mask, rel_index, taps, dpr[bi], dp_gen)
            bi += 1
        if i < len(depths) - 1:
            # PatchMerging.forward, :293-312
            p = "%slayers.%d.downsample" % (prefix, i)
            if h % 2 == 1 or w % 2 == 1:
                x = F.pad(x, (0, 0, 0, w % 2, 0, h % 2))
            x = torch.cat([x[:, :, 0::2, 0::2, :], x[:, :, 1::2, 0::2, :], x[:, :, 0::2, 1::2, :],
                           x[:, :, 1::2, 1::2, :]], -1)
            x = _ln(x, sd, p + ".norm", taps)
            x = F.linear(x, sd[p + ".reduction.weight"])
        x = x.permute(0, 4, 1, 2, 3).contiguous()
    x = x.permute(0, 2, 3, 4, 1)
    x = _ln(x.contiguous(), sd, prefix + "norm", taps)
    x = x.permute(0, 4, 1, 2, 3)
    feat = x.mean((2, 3, 4))  # AdaptiveAvgPool3d(1)
    if dropout_p > 0:
        feat = F.dropout(feat, dropout_p, True)
    score = F.linear(feat, sd["cls_head.fc_cls.weight"], sd["cls_head.fc_cls.bias"]).view(n, v, -1)
    return score.mean(1), score


# --------------------------------------------------------------------------------------------
# a1 / a17: the adaptation step
# --------------------------------------------------------------------------------------------
class TTAState:
    """Everything ``tta_standard`` (corpus/basics.py:403-747) keeps between steps in tta_online mode:
    trainable tensors, one SGD optimiser over *all* parameters (:559-560) and one AlignTap per norm
    layer whose name contains a chosen block (:571-587)."""

    def __init__(self, sd, arch, clip_len, src_means, src_vars, chosen_blocks, reg_type="l1_loss",
                 moving_avg=True, momentum_mvg=0.1, lr=5e-5, momentum=0.9, weight_decay=5e-4,
                 swin_cfg=None, name_prefix="", update_only_bn_affine=False, stat_reg="mean_var",
                 running_manner=True, momentum_bns=0.1, before_norm=False):
        self.arch, self.clip_len = arch, clip_len
        self.swin_cfg = swin_cfg or {}
        self.sd = {}
        if update_only_bn_affine:
            # corpus/basics.py:547-557 + utils/BNS_utils.py:262-288: everything frozen except the norm layers
            # (TANet: BatchNorm1d/2d/3d, Swin: every LayerNorm incl. patch_embed.norm); Adam over their weight / bias
            if arch == "tanet":
                norm_names = [n for n, _ in tanet_norm_layers()]
            else:
                norm_names = swin_norm_layers((self.swin_cfg or {}).get("depths", (2, 2, 18, 2)))
            trainable = {n + leaf for n in norm_names for leaf in (".weight", ".bias")}
        params = []
        for k, v in sd.items():
            v = v.detach().clone()
            if v.is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var")):
                if not update_only_bn_affine or k in trainable:
                    v.requires_grad_(True)
                    params.append(v)
            self.sd[k] = v
        self.params = params
        if update_only_bn_affine:
            self.opt = torch.optim.Adam(params, lr=lr, betas=(0.9, 0.999), weight_decay=0.)
        else:
            self.opt = torch.optim.SGD(params, lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.taps = {}
        if stat_reg == "BNS":
            # corpus/basics.py:588-600: a BNFeatureHook on EVERY BatchNorm (1d/2d/3d) of the chosen blocks, target = the
            # layer's own running statistics as they are when the hook is created
            assert arch == "tanet", "BNS is defined on BatchNorm layers"
            for name, kind in tanet_norm_layers():
                if any(b in name_prefix + name for b in chosen_blocks):
                    self.taps[name] = BnsTap(self.sd[name + ".running_mean"], self.sd[name + ".running_var"], reg_type,
                                             running_manner, momentum_bns)
                    self.taps[name].kind = "bns"
        elif arch == "tanet":
            layers = tanet_norm_layers()
            it = iter(range(len(src_means)))
            for name, kind in layers:
                idx = None if kind == "bn1d" else next(it)  # basics.py:488-498: None placeholders at BN1d
                full = name_prefix + name
                if any(b in full for b in chosen_blocks):
                    self.taps[name] = AlignTap(kind, clip_len, None if idx is None else src_means[idx],
                                               None if idx is None else src_vars[idx], reg_type, moving_avg,
                                               momentum_mvg, before_norm)
        else:
            names = swin_norm_layers(self.swin_cfg.get("depths", (2, 2, 18, 2)))[1:]
            assert len(names) == len(src_means)
            for i, name in enumerate(names):
                full = name_prefix + name
                if any(b in full for b in chosen_blocks):
                    self.taps[name] = AlignTap("ln", clip_len, src_means[i], src_vars[i], reg_type, moving_avg,
                                               momentum_mvg, before_norm)

    def forward(self, inp, taps, dropout_p=0.0, drop_path_rate=0.0, dp_gen=None):
        if self.arch == "tanet":
            return tanet_forward(self.sd, inp, self.clip_len, taps, dropout_p)
        return swin_forward(self.sd, inp, taps=taps, dropout_p=dropout_p, drop_path_rate=drop_path_rate,
                            dp_gen=dp_gen, **self.swin_cfg)

    def adapt_step(self, inp, n_videos, n_views, if_pred_consistency, lambda_feature_reg=1.0,
                   lambda_pred_consis=0.1, dropout_p=0.0, drop_path_rate=0.0, dp_gen=None):
        """Loop body corpus/basics.py:606-677.  inp: TANet (N*M, T, 3, H, W); Swin (N, M, 3, T, H, W)."""
        loss_consis = None
        if self.arch == "tanet":
            logits = self.forward(inp, self.taps, dropout_p)
            view_logits = logits.reshape(n_videos, n_views, -1)
            if if_pred_consistency:
                loss_consis = pred_consistency(view_logits)
            out = view_logits.mean(1)
        else:
            out, view_logits = self.forward(inp, self.taps, dropout_p, drop_path_rate, dp_gen)
            if if_pred_consistency:
                loss_consis = pred_consistency(view_logits)
        loss_reg = torch.tensor(0.0)
        for tap in self.taps.values():
            loss_reg = loss_reg + tap.r_feature
        if if_pred_consistency:
            loss = lambda_feature_reg * loss_reg + lambda_pred_consis * loss_consis
        else:
            loss = loss_reg  # :667 -- lambda_feature_reg is NOT applied in this branch
        self.opt.zero_grad()
        loss.backward()
        self.opt.step()
        return {"logits": out.detach(), "view_logits": view_logits.detach(), "loss_reg": loss_reg.detach(),
                "loss_consis": None if loss_consis is None else loss_consis.detach(), "loss": loss.detach()}

    @torch.no_grad()
    def eval_forward(self, inp):
        """Per-step clean evaluation forward, corpus/basics.py:691-713 (hooks removed, model.eval())."""
        if self.arch == "tanet":
            return tanet_forward(self.sd, inp, self.clip_len, None, 0.0)
        return swin_forward(self.sd, inp, taps=None, **self.swin_cfg)[0]


def collect_source_stats(sd, arch, clip_len, batches, swin_cfg=None, before_norm=False):
    """compute_statistics, corpus/basics.py:220-307: model.eval(), ComputeNormStatsHook on every
    BN2d/BN3d (TANet) or LN[1:] (Swin); the per-batch mean and *per-batch biased variance* are averaged
    with AverageMeter(n=batch) (:298-304) -- i.e. NOT a global variance."""
    swin_cfg = swin_cfg or {}
    if arch == "tanet":
        names = [(n, k) for n, k in tanet_norm_layers() if k != "bn1d"]
    else:
        names = [(n, "ln") for n in swin_norm_layers(swin_cfg.get("depths", (2, 2, 18, 2)))[1:]]
    taps = {n: StatTap(k, clip_len, "spatiotemp", before_norm) for n, k in names}
    sm = [0.0] * len(names)
    sv = [0.0] * len(names)
    cnt = 0
    with torch.no_grad():
        for inp in batches:
            if arch == "tanet":
                tanet_forward(sd, inp, clip_len, taps)
                bz = inp.shape[0]
            else:
                swin_forward(sd, inp, taps=taps, **swin_cfg)
                bz = inp.shape[0]
            for i, (n, _) in enumerate(names):
                sm[i] = sm[i] + taps[n].batch_mean * bz
                sv[i] = sv[i] + taps[n].batch_var * bz
            cnt += bz
    return [(m / cnt).numpy() for m in sm], [(v / cnt).numpy() for v in sv]
