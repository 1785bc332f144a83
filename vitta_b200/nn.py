"""Norm modules that can host a statistics *tap* (see utils/norm_stats_utils.py) and the fused call path."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class StatsBatchNorm2d(nn.BatchNorm2d):
    """nn.BatchNorm2d (same parameters / buffers / state-dict keys).  Called as a module it behaves exactly
    like its base class, so foreign forward hooks keep working; the vitta_b200 models call
    :func:`norm_act` instead, which runs the fused sm_100a kernel when that is legal."""
    _vitta_fused = True
    _vitta_tap = None


class StatsLayerNorm(nn.LayerNorm):
    """nn.LayerNorm (same parameters / state-dict keys) that can host a statistics tap.  The vitta_b200 Swin modules
    call :func:`layer_norm` instead of the module, which runs the fused sm_100a kernel (K9) when that is legal."""
    _vitta_fused = True
    _vitta_tap = None


def layer_norm(ln, x, want_alias=True):
    """x: (..., C) contiguous tokens -> (normalised rows (rows, C), alias of x as rows for the shortcut).
    Fused path: LayerNorm + the statistics partials of an attached tap in one pass; its backward adds the hook
    gradient and the shortcut gradient.  With foreign forward hooks on the module the module itself is called so that
    every hook observes the reference's tensors (``(B, D, H, W, C)`` output)."""
    from . import ops_swin
    c = x.shape[-1]
    if (isinstance(ln, StatsLayerNorm) and not ln._forward_hooks and not ln._forward_pre_hooks and x.is_cuda
            and ln.elementwise_affine):
        arena = ly = None
        tap = ln._vitta_tap
        if tap is not None:
            arena, ly = tap.tap_target()
            tap.note_batch(x.shape[0])
        rows = x.reshape(-1, c)
        if not rows.is_contiguous():
            rows = rows.contiguous()
        return ops_swin.layer_norm_rows(rows, ln.weight, ln.bias, ln.eps, arena, ly, want_alias)
    y = ln(x)
    return y.reshape(-1, c), (x.reshape(-1, c) if want_alias else None)


def _fusable(bn):
    return (bn is not None and isinstance(bn, StatsBatchNorm2d) and not bn.training and not bn._forward_hooks
            and not bn._forward_pre_hooks)


def norm_act(bn, x, relu, clip_len, res=None, res_bn=None, want_pool=False):
    """out = relu?(bn(x) + [res | res_bn(res)]); returns (out, per-frame mean of out or None).

    Fused path (K4): eval-mode BatchNorm without foreign hooks -- one kernel reads x (and res) once, writes
    out once, and emits the statistics partials of whatever taps are attached to ``bn`` / ``res_bn``.
    Otherwise (BatchNorm in training mode, i.e. ``fix_BNS=False``, or third-party forward hooks present) the
    modules are called one by one so that every hook observes the reference's tensors."""
    if _fusable(bn) and (res_bn is None or _fusable(res_bn)) and x.is_cuda:
        arena = ly_main = ly_res = None
        tap = bn._vitta_tap
        n_clips = x.shape[0] // clip_len if clip_len else x.shape[0]
        if tap is not None:
            arena, ly_main = tap.tap_target()
            tap.note_batch(n_clips)
        if res_bn is not None and res_bn._vitta_tap is not None:
            a2, ly_res = res_bn._vitta_tap.tap_target()
            res_bn._vitta_tap.note_batch(n_clips)
            if arena is not None and a2 is not arena:
                raise RuntimeError("taps of one block belong to different arenas")
            arena = a2
        x = x.contiguous(memory_format=torch.channels_last)
        if res is not None:
            res = res.contiguous(memory_format=torch.channels_last)
        return ops.bn_act(x, bn, relu, res, res_bn, arena, ly_main, ly_res, want_pool)
    y = bn(x)
    if res is not None:
        y = y + (res_bn(res) if res_bn is not None else res)
    if relu:
        y = F.relu(y)
    pool = F.adaptive_avg_pool2d(y, 1).flatten(1) if want_pool else None
    return y, pool


def conv2d_shortcut(conv, x):
    """(conv(x), alias of x): for the first convolution of a residual block whose input also feeds the shortcut.  The
    alias' gradient is added inside the data-gradient kernel of ``conv`` (no separate accumulation pass)."""
    if not x.is_cuda:
        raise ops._lib.VittaError("vitta_b200 models need CUDA tensors; there is no CPU path")
    kh, kw = conv.kernel_size
    if (conv.bias is None and conv.groups == 1 and conv.dilation == (1, 1) and conv.in_channels % 4 == 0
            and conv.stride == (1, 1) and conv.padding[0] == conv.padding[1] and kh == kw
            and conv.padding_mode == 'zeros' and not conv._forward_hooks and not conv._forward_pre_hooks
            and torch.is_grad_enabled() and x.requires_grad):
        x = x.contiguous(memory_format=torch.channels_last)
        return ops.conv2d_shortcut(x, conv.weight, 1, conv.padding[0])
    return conv2d(conv, x), x


def conv2d(conv, x):
    """nn.Conv2d call path of the vitta_b200 models: bias-free, ungrouped, undilated convolutions whose input has a
    multiple of 4 channels run on the tcgen05 3xTF32 implicit-GEMM kernel (K6); the 3-channel stem convolution is
    the one layer left to the library (its rows are not 16-byte aligned for TMA)."""
    if not x.is_cuda:
        raise ops._lib.VittaError("vitta_b200 models need CUDA tensors; there is no CPU path")
    kh, kw = conv.kernel_size
    if (conv.bias is None and conv.groups == 1 and conv.dilation == (1, 1) and conv.in_channels % 4 == 0
            and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1] and kh == kw
            and conv.padding_mode == 'zeros' and not conv._forward_hooks and not conv._forward_pre_hooks):
        x = x.contiguous(memory_format=torch.channels_last)
        return ops.conv2d(x, conv.weight, conv.stride[0], conv.padding[0])
    return conv(x)


def stem(conv, bn, maxpool, x, clip_len):
    """conv1 -> bn1 -> relu -> maxpool of the ResNet trunk.  Fast path: the 7x7/2 convolution on the tcgen05 kernel and
    BN + ReLU + MaxPool(3, 2, 1) as ONE pass (no 411 MB un-pooled activation).  It is taken when the BatchNorm is in eval
    mode without a statistics tap or foreign hooks (the default ViTTA configuration aligns layer3 / layer4 only); otherwise
    the layers run one by one so that every hook observes the reference's tensors."""
    if ops.stem_conv_supported(conv, x) and x.is_cuda:
        y = ops.StemConvFn.apply(x, conv.weight)
    else:
        y = conv2d(conv, x.contiguous(memory_format=torch.channels_last))
    if (_fusable(bn) and bn._vitta_tap is None and isinstance(maxpool, nn.MaxPool2d) and maxpool.kernel_size == 3
            and maxpool.stride == 2 and maxpool.padding == 1 and maxpool.dilation == 1 and not maxpool.ceil_mode
            and not maxpool.return_indices and not maxpool._forward_hooks and y.shape[1] % 4 == 0
            and 256 % (y.shape[1] // 4) == 0 and y.is_cuda):
        y = y.contiguous(memory_format=torch.channels_last)
        return ops.BnReluPoolFn.apply(y, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    y, _ = norm_act(bn, y, True, clip_len)
    return maxpool(y)
