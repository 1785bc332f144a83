"""TAM and TemporalBottleneck -- mirror of the reference's ``models/tanet_models/temporal_module.py`` with the
heavy tensor work moved into sm_100a kernels (K4 fused norm/act/stats/pool, K5 temporal stencil).

Module / parameter names are the reference's (``net.conv1 .. net.bn3``, ``net.downsample.{0,1}``, ``tam.G.{0,1,3}``,
``tam.L.{0,1,3}``) so reference checkpoints load unchanged and ``choose_layers`` enumerates the norm layers in
the same order."""
import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ...nn import StatsBatchNorm2d, conv2d, conv2d_shortcut, norm_act


class TAM(nn.Module):
    """Temporal Adaptive Module (reference :12-65).

    G (global branch): per (video, channel) a softmax-normalised 3-tap temporal kernel from the spatially
    pooled T-vector.  L (local branch): a sigmoid gate per (video, channel, frame).  Output:
        out[n,t,c,:] = sum_k K[n,c,k] * L[n,c,t+k-1] * x[n,t+k-1,c,:]      (zero padded in t)
    The pooled input normally arrives from the preceding fused norm kernel (``pooled``); the tiny G/L networks
    run as ordinary torch modules; the stencil over the (N, T, HW, C) activation is kernel K5."""

    def __init__(self, in_channels, n_segment, kernel_size=3, stride=1, padding=1):
        super().__init__()
        if kernel_size != 3 or stride != 1 or padding != 1:
            raise NotImplementedError("the sm_100a stencil implements the reference's only configuration "
                                      "(kernel 3, stride 1, padding 1; temporal_module.py:134-139)")
        self.in_channels = in_channels
        self.n_segment = n_segment
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.G = nn.Sequential(
            nn.Linear(n_segment, n_segment * 2, bias=False),
            nn.BatchNorm1d(n_segment * 2), nn.ReLU(inplace=True),
            nn.Linear(n_segment * 2, kernel_size, bias=False), nn.Softmax(-1))
        self.L = nn.Sequential(
            nn.Conv1d(in_channels, in_channels // 4, kernel_size, stride=1, padding=kernel_size // 2, bias=False),
            nn.BatchNorm1d(in_channels // 4), nn.ReLU(inplace=True),
            nn.Conv1d(in_channels // 4, in_channels, 1, bias=False), nn.Sigmoid())

    def forward(self, x, pooled=None):
        nt, c, h, w = x.shape
        t = self.n_segment
        n = nt // t
        if pooled is None:
            pooled = F.adaptive_avg_pool2d(x, 1).flatten(1)          # (N*T, C)
        if self._gate_fusable(x, t, c):
            # both gate networks in 2 launches (K5b); BatchNorm1d hooks of the alignment driver contribute exact zeros
            # (reference utils/norm_stats_utils.py:158-183) and are marked as fired without running the modules
            g_bn, l_bn = self.G[1], self.L[1]
            for bn in (g_bn, l_bn):
                for fn in bn._forward_hooks.values():
                    fn.__self__.mark_bn1d_fired(x.device)
            kern, act = ops.TamGateFn.apply(pooled, self.G[0].weight, g_bn.weight, g_bn.bias, g_bn.running_mean,
                                            g_bn.running_var, self.G[3].weight, self.L[0].weight, l_bn.weight, l_bn.bias,
                                            l_bn.running_mean, l_bn.running_var, self.L[3].weight.view(c, c // 4), t,
                                            g_bn.eps, l_bn.eps)
            x = x.contiguous(memory_format=torch.channels_last)
            return ops.TamStencilFn.apply(x, kern, act, t)
        p_ntc = pooled.view(n, t, c)
        # G: per (video, channel) a softmax-normalised 3-tap kernel from the pooled T-vector
        kern = self.G(p_ntc.permute(0, 2, 1).reshape(n * c, t)).view(n, c, self.kernel_size)
        kern = kern.permute(0, 2, 1).contiguous()                                                    # (N, 3, C)
        act = self._local_gate(p_ntc, n, t, c)                                                       # (N, T, C)
        x = x.contiguous(memory_format=torch.channels_last)
        return ops.TamStencilFn.apply(x, kern, act, t)

    def _gate_fusable(self, x, t, c):
        """K5b applies when both BatchNorm1d layers are in eval mode (``fix_BNS``, the ViTTA default), carry affine
        parameters, and no module of the branches has hooks other than the alignment driver's no-op BatchNorm1d hooks."""
        if not x.is_cuda or t < 2 or t > 16 or c % 4 != 0:
            return False
        from ...utils.norm_stats_utils import CombineNormStatsRegHook_onereg
        for seq in (self.G, self.L):
            for m in seq:
                if m._forward_pre_hooks:
                    return False
                for fn in m._forward_hooks.values():
                    owner = getattr(fn, '__self__', None)
                    if not (isinstance(owner, CombineNormStatsRegHook_onereg) and owner._is_bn1d):
                        return False
        for bn in (self.G[1], self.L[1]):
            if bn.training or not bn.affine or not bn.track_running_stats:
                return False
        return True

    def _local_gate(self, p_ntc, n, t, c):
        """L branch (reference :35-41,52-55) on the (N, T, C) pooled tensor.  The k=3 temporal Conv1d is evaluated as ONE
        linear map over the three time-shifted copies (rows = (video, frame)), the k=1 Conv1d as a plain linear map:
        identical arithmetic, but small GEMMs instead of cuDNN's conv1d paths with their NCHW<->NHWC conversions.  The
        modules (and therefore parameter names, BatchNorm1d hooks and train/eval behaviour) are the reference's."""
        conv_a, bn, relu, conv_b, gate = self.L
        if (conv_a._forward_hooks or conv_b._forward_hooks or conv_a._forward_pre_hooks or conv_b._forward_pre_hooks):
            return self.L(p_ntc.permute(0, 2, 1).contiguous()).permute(0, 2, 1).contiguous()
        z = F.pad(p_ntc, (0, 0, 1, 1))                                                   # zero padding in t
        x3 = torch.cat([z[:, :-2], z[:, 1:-1], z[:, 2:]], dim=2).reshape(n * t, 3 * c)    # taps t-1, t, t+1
        w_a = conv_a.weight.permute(0, 2, 1).reshape(conv_a.out_channels, 3 * c)          # [out][tap][in]
        hdn = relu(bn(F.linear(x3, w_a)))                                                # BatchNorm1d on (rows, C/4)
        out = gate(F.linear(hdn, conv_b.weight.view(c, conv_b.in_channels)))
        return out.view(n, t, c)


class Bottleneck(nn.Module):
    """torchvision ResNet v1.5 bottleneck container (stride on the 3x3 conv); only holds the layers."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = StatsBatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = StatsBatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = StatsBatchNorm2d(planes * self.expansion)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class TemporalBottleneck(nn.Module):
    """conv1-bn1-relu-TAM-conv2-bn2-relu-conv3-bn3 (+downsample) -add-relu (reference :68-106)."""

    def __init__(self, net, n_segment=8, t_kernel_size=3, t_stride=1, t_padding=1):
        super().__init__()
        self.net = net
        assert isinstance(net, Bottleneck)
        self.n_segment = n_segment
        self.tam = TAM(in_channels=net.conv1.out_channels, n_segment=n_segment, kernel_size=t_kernel_size,
                       stride=t_stride, padding=t_padding)

    def infer_pairs(self):
        """(convolution, BatchNorm) pairs of this block in the order the inference forward uses them."""
        net = self.net
        pairs = [(net.conv1, net.bn1), (net.conv2, net.bn2), (net.conv3, net.bn3)]
        if net.downsample is not None:
            pairs.append((net.downsample[0], net.downsample[1]))
        return pairs

    def forward_infer(self, x, folds, want_pool=False):
        """Inference forward with every eval-mode BatchNorm folded into its convolution (K6 epilogue: + bias, + shortcut,
        ReLU): no norm pass, no statistics, nothing saved for a backward.  Same arithmetic up to the association
        conv(x, k*W) + b' vs (conv(x, W) - mean) * k + beta."""
        net = self.net
        out = ops.conv2d_folded(x, net.conv1, folds, True)
        out = self.tam(out, ops.frame_mean_cl(out))
        out = ops.conv2d_folded(out, net.conv2, folds, True)
        idt = x if net.downsample is None else ops.conv2d_folded(x, net.downsample[0], folds, False)
        out = ops.conv2d_folded(out, net.conv3, folds, True, residual=idt)
        return out, (ops.frame_mean_cl(out) if want_pool else None)

    def forward(self, x, want_pool=False):
        net, t = self.net, self.n_segment
        out, x = conv2d_shortcut(net.conv1, x)                              # x: alias for the shortcut path
        out, pooled = norm_act(net.bn1, out, True, t, want_pool=True)       # BN + stats + ReLU + HW-pool: 1 pass
        out = self.tam(out, pooled)
        out = conv2d(net.conv2, out)
        out, _ = norm_act(net.bn2, out, True, t)
        out = conv2d(net.conv3, out)
        if net.downsample is not None:
            idt = conv2d(net.downsample[0], x)
            return norm_act(net.bn3, out, True, t, res=idt, res_bn=net.downsample[1], want_pool=want_pool)
        return norm_act(net.bn3, out, True, t, res=x, want_pool=want_pool)


def make_temporal_modeling(net, n_segment=8, t_kernel_size=3, t_stride=1, t_padding=1):
    """Wrap every Bottleneck of layer1..layer4 in a TemporalBottleneck (reference :109-140)."""
    for name in ("layer1", "layer2", "layer3", "layer4"):
        stage = getattr(net, name)
        setattr(net, name, nn.Sequential(*[TemporalBottleneck(b, n_segment, t_kernel_size, t_stride, t_padding)
                                           for b in stage.children()]))
