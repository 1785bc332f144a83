"""Video Swin Transformer -- mirror of the reference's ``models/videoswintransformer_models/swin_transformer.py``.

Same class names, constructor arguments, attribute / parameter names and state-dict keys
(``patch_embed.proj``, ``layers.i.blocks.j.{norm1,attn.{relative_position_bias_table,relative_position_index,qkv,proj},
norm2,mlp.{fc1,fc2}}``, ``layers.i.downsample.{reduction,norm}``, ``norm``), so reference checkpoints load unchanged and
``choose_layers`` enumerates the LayerNorms in the same order.  The arithmetic is not here: activations stay one
``(B, D, H, W, C)`` token matrix from patch embedding to the head, and each half-block is one fused operator of
``vitta_b200.ops_swin`` (K7 window attention, K8 GEMM epilogues, K9 LayerNorm + statistics).  roll / window_partition /
window_reverse / compute_mask of the reference (:38-66, :229-248, :316-329) do not exist as tensors -- they are index
maps inside the attention kernel.
"""
import numpy as np
import torch
import torch.nn as nn

from ... import ops_swin
from ..._lib import VittaError
from ...nn import StatsLayerNorm, layer_norm


class DropPath(nn.Module):
    """Stochastic depth per sample (timm 0.6.7 ``DropPath``, call site reference :210).  Holds only the rate: the
    per-sample factor it draws is applied inside the GEMM epilogue that also adds the shortcut."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.scale_by_keep = scale_by_keep

    def sample(self, n, device):
        """(n,) factors, or None when inactive."""
        if self.drop_prob == 0.0 or not self.training:
            return None
        keep = 1.0 - self.drop_prob
        r = torch.empty(n, dtype=torch.float32, device=device).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            r.div_(keep)
        return r

    def forward(self, x):
        r = self.sample(x.shape[0], x.device)
        return x if r is None else x * r.view((-1,) + (1,) * (x.dim() - 1))


class Mlp(nn.Module):
    """fc1 -> GELU -> fc2 (reference :17-35).  Parameter container; SwinTransformerBlock3D runs it fused."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU or drop != 0.:
            raise NotImplementedError("the fused MLP implements the reference's configuration (exact GELU, drop 0)")
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)


def get_window_size(x_size, window_size, shift_size=None):
    """reference :71-84."""
    use_window_size = list(window_size)
    use_shift_size = list(shift_size) if shift_size is not None else None
    for i in range(len(x_size)):
        if x_size[i] <= window_size[i]:
            use_window_size[i] = x_size[i]
            if shift_size is not None:
                use_shift_size[i] = 0
    if shift_size is None:
        return tuple(use_window_size)
    return tuple(use_window_size), tuple(use_shift_size)


def relative_position_index(window_size):
    """The buffer of reference :113-125."""
    coords = torch.stack(torch.meshgrid(*[torch.arange(s) for s in window_size], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window_size[0] - 1
    rel[:, :, 1] += window_size[1] - 1
    rel[:, :, 2] += window_size[2] - 1
    rel[:, :, 0] *= (2 * window_size[1] - 1) * (2 * window_size[2] - 1)
    rel[:, :, 1] *= 2 * window_size[2] - 1
    return rel.sum(-1)


class WindowAttention3D(nn.Module):
    """Parameter container of the window attention (reference :87-169): ``relative_position_bias_table``,
    ``relative_position_index`` (buffer, kept for state-dict compatibility; the kernel recomputes the index from
    coordinates), ``qkv``, ``proj``."""

    def __init__(self, dim, window_size, num_heads, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        if attn_drop != 0. or proj_drop != 0.:
            raise NotImplementedError("attention / projection dropout are 0 in every ViTTA configuration")
        self.dim = dim
        self.window_size = tuple(window_size)
        self.num_heads = num_heads
        head_dim = dim // num_heads
        if head_dim != 32 or dim % num_heads:
            raise NotImplementedError("the sm_100a window-attention kernel is specialised for head_dim 32 "
                                      "(Swin-T/S/B/L all use 32)")
        self.scale = qk_scale or head_dim ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros(
            (2 * window_size[0] - 1) * (2 * window_size[1] - 1) * (2 * window_size[2] - 1), num_heads))
        self.register_buffer("relative_position_index", relative_position_index(self.window_size))
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=.02)
        self.softmax = nn.Softmax(dim=-1)


class SwinTransformerBlock3D(nn.Module):
    """x + DropPath(W-MSA(LN(x))), then x + DropPath(MLP(LN(x)))  (reference :172-274)."""

    def __init__(self, dim, num_heads, window_size=(2, 7, 7), shift_size=(0, 0, 0), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop=0., attn_drop=0., drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm,
                 use_checkpoint=False):
        super().__init__()
        self.dim = dim
        self.num_heads = num_heads
        self.window_size = tuple(window_size)
        self.shift_size = tuple(shift_size)
        self.mlp_ratio = mlp_ratio
        self.use_checkpoint = use_checkpoint
        assert 0 <= self.shift_size[0] < self.window_size[0], "shift_size must in 0-window_size"
        assert 0 <= self.shift_size[1] < self.window_size[1], "shift_size must in 0-window_size"
        assert 0 <= self.shift_size[2] < self.window_size[2], "shift_size must in 0-window_size"
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("norm_layer must be nn.LayerNorm")
        self.norm1 = StatsLayerNorm(dim)
        self.attn = WindowAttention3D(dim, window_size=self.window_size, num_heads=num_heads, qkv_bias=qkv_bias,
                                      qk_scale=qk_scale, attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = StatsLayerNorm(dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def _drop_factors(self, n, device):
        return self.drop_path.sample(n, device) if isinstance(self.drop_path, DropPath) else None

    def forward(self, x, mask_matrix=None):
        """x: (B, D, H, W, C) -> same shape.  ``mask_matrix`` is accepted for signature compatibility and ignored (the
        kernel derives the shift mask from region ids)."""
        b, d, h, w, c = x.shape
        a = self.attn
        y1, short1 = layer_norm(self.norm1, x)
        x1 = ops_swin.SwinAttentionFn.apply(y1, short1, a.qkv.weight, a.qkv.bias, a.relative_position_bias_table,
                                            a.proj.weight, a.proj.bias, (b, d, h, w), self.num_heads, self.window_size,
                                            self.shift_size, a.scale, self._drop_factors(b, x.device))
        y2, short2 = layer_norm(self.norm2, x1.view(b, d, h, w, c))
        m = self.mlp
        x2 = ops_swin.SwinMlpFn.apply(y2, short2, m.fc1.weight, m.fc1.bias, m.fc2.weight, m.fc2.bias, b,
                                      self._drop_factors(b, x.device))
        return x2.view(b, d, h, w, c)


class PatchMerging(nn.Module):
    """2x2 spatial merge -> LayerNorm(4C) -> Linear(4C, 2C, no bias)  (reference :277-312), one fused operator."""

    def __init__(self, dim, norm_layer=nn.LayerNorm):
        super().__init__()
        self.dim = dim
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = StatsLayerNorm(4 * dim)

    def forward(self, x):
        b, d, h, w, c = x.shape
        n = self.norm
        if n._forward_hooks or n._forward_pre_hooks:
            raise VittaError("PatchMerging.norm carries foreign forward hooks; attach vitta_b200 hook classes instead")
        arena = ly = None
        if n._vitta_tap is not None:
            arena, ly = n._vitta_tap.tap_target()
            n._vitta_tap.note_batch(b)
        rows = x.reshape(-1, c)
        out, tok = ops_swin.PatchMergeFn.apply(rows, n.weight, n.bias, self.reduction.weight, n.eps, (b, d, h, w), arena, ly)
        if ly is not None:
            ly.token = tok
        return out.view(b, d, (h + 1) // 2, (w + 1) // 2, 2 * c)


class BasicLayer(nn.Module):
    """One stage (reference :332-413).  Input and output are channels-last token volumes (B, D, H, W, C)."""

    def __init__(self, dim, depth, num_heads, window_size=(1, 7, 7), mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0.,
                 attn_drop=0., drop_path=0., norm_layer=nn.LayerNorm, downsample=None, use_checkpoint=False):
        super().__init__()
        self.window_size = tuple(window_size)
        self.shift_size = tuple(i // 2 for i in window_size)
        self.depth = depth
        self.use_checkpoint = use_checkpoint
        self.blocks = nn.ModuleList([
            SwinTransformerBlock3D(dim=dim, num_heads=num_heads, window_size=window_size,
                                   shift_size=(0, 0, 0) if (i % 2 == 0) else self.shift_size, mlp_ratio=mlp_ratio,
                                   qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop, attn_drop=attn_drop,
                                   drop_path=drop_path[i] if isinstance(drop_path, list) else drop_path,
                                   norm_layer=norm_layer, use_checkpoint=use_checkpoint)
            for i in range(depth)])
        self.downsample = downsample
        if self.downsample is not None:
            self.downsample = downsample(dim=dim, norm_layer=norm_layer)

    def forward(self, x):
        for blk in self.blocks:
            x = blk(x, None)
        if self.downsample is not None:
            x = self.downsample(x)
        return x


class PatchEmbed3D(nn.Module):
    """Conv3d(kernel = stride = patch) + LayerNorm (reference :416-456) as patchify -> GEMM -> LayerNorm.
    Returns the channels-last token volume (B, D, H/ph, W/pw, C)."""

    def __init__(self, patch_size=(2, 4, 4), in_chans=3, embed_dim=96, norm_layer=None):
        super().__init__()
        self.patch_size = tuple(patch_size)
        self.in_chans = in_chans
        self.embed_dim = embed_dim
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = StatsLayerNorm(embed_dim) if norm_layer is not None else None

    def forward(self, x):
        b, _, t, h, w = x.shape
        pt, ph, pw = self.patch_size
        if t % pt or h % ph or w % pw:
            x = nn.functional.pad(x, (0, (pw - w % pw) % pw, 0, (ph - h % ph) % ph, 0, (pt - t % pt) % pt))  # :440-446
            b, _, t, h, w = x.shape
        n = self.norm
        if n is not None and (n._forward_hooks or n._forward_pre_hooks or n._vitta_tap is not None):
            raise VittaError("patch_embed.norm cannot be hooked (the reference skips it too, corpus/basics.py:541-543)")
        y = ops_swin.PatchEmbedFn.apply(x, self.proj.weight, self.proj.bias, None if n is None else n.weight,
                                        None if n is None else n.bias, 1e-5 if n is None else n.eps, self.patch_size)
        return y.view(b, t // pt, h // ph, w // pw, self.embed_dim)


class SwinTransformer3D(nn.Module):
    """Backbone (reference :459-668).  ``forward`` returns (B, C, D, H, W) like the reference -- as a permuted VIEW of
    the channels-last token volume, which the vitta_b200 head consumes without a copy."""

    def __init__(self, pretrained=None, pretrained2d=True, patch_size=(4, 4, 4), in_chans=3, embed_dim=96,
                 depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24], window_size=(2, 7, 7), mlp_ratio=4., qkv_bias=True,
                 qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.2, norm_layer=nn.LayerNorm,
                 patch_norm=False, frozen_stages=-1, use_checkpoint=False):
        super().__init__()
        if drop_rate != 0. or attn_drop_rate != 0.:
            raise NotImplementedError("drop_rate / attn_drop_rate are 0 in every ViTTA configuration")
        self.pretrained = pretrained
        self.pretrained2d = pretrained2d
        self.num_layers = len(depths)
        self.embed_dim = embed_dim
        self.patch_norm = patch_norm
        self.frozen_stages = frozen_stages
        self.window_size = tuple(window_size)
        self.patch_size = tuple(patch_size)
        self.patch_embed = PatchEmbed3D(patch_size=patch_size, in_chans=in_chans, embed_dim=embed_dim,
                                        norm_layer=norm_layer if self.patch_norm else None)
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, sum(depths))]
        self.layers = nn.ModuleList()
        for i_layer in range(self.num_layers):
            self.layers.append(BasicLayer(
                dim=int(embed_dim * 2 ** i_layer), depth=depths[i_layer], num_heads=num_heads[i_layer],
                window_size=window_size, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, drop=drop_rate,
                attn_drop=attn_drop_rate, drop_path=dpr[sum(depths[:i_layer]):sum(depths[:i_layer + 1])],
                norm_layer=norm_layer, downsample=PatchMerging if i_layer < self.num_layers - 1 else None,
                use_checkpoint=use_checkpoint))
        self.num_features = int(embed_dim * 2 ** (self.num_layers - 1))
        self.norm = StatsLayerNorm(self.num_features)
        self.apply(self._init_weights)
        self._freeze_stages()

    @staticmethod
    def _init_weights(m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _freeze_stages(self):
        if self.frozen_stages >= 0:
            self.patch_embed.eval()
            for param in self.patch_embed.parameters():
                param.requires_grad = False
        if self.frozen_stages >= 1:
            self.pos_drop.eval()
            for i in range(0, self.frozen_stages):
                m = self.layers[i]
                m.eval()
                for param in m.parameters():
                    param.requires_grad = False

    def init_weights(self, pretrained=None):
        self.apply(self._init_weights)

    def forward_tokens(self, x):
        """(B, 3, T, H, W) -> final-norm tokens (B, D, H', W', C)."""
        x = self.patch_embed(x)
        for layer in self.layers:
            x = layer(x)
        b, d, h, w, c = x.shape
        y, _ = layer_norm(self.norm, x, want_alias=False)
        return y.view(b, d, h, w, c)

    def forward(self, x):
        return self.forward_tokens(x).permute(0, 4, 1, 2, 3)

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        return self
