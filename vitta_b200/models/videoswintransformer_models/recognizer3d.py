"""Recognizer3D -- mirror of the reference's ``models/videoswintransformer_models/recognizer3d.py`` (:43-115):
Video-Swin backbone + I3DHead, ``(N, V, 3, T, H, W) -> (video scores (N, K), per-view scores (N, V, K))``.

The reference hard-codes Swin-B (embed 128, depths [2,2,18,2], heads [4,8,16,32], head input 1024; :53-55,67).  Those
stay the defaults; ``embed_dim`` / ``depths`` / ``num_heads`` are exposed as extra keyword arguments so that Swin-T
(BASELINE.json configs[2]) and small test models can be built with the same class."""
import torch.nn as nn
import torch.nn.functional as F

from .i3d_head import I3DHead
from .swin_transformer import SwinTransformer3D


class Recognizer3D(nn.Module):
    def __init__(self, num_classes=None, patch_size=None, window_size=None, drop_path_rate=None, embed_dim=128,
                 depths=(2, 2, 18, 2), num_heads=(4, 8, 16, 32)):
        super().__init__()
        self.pretrained = None
        self.pretrained2d = True
        self.patch_size = patch_size
        self.in_chans = 3
        self.embed_dim = embed_dim
        self.depths = list(depths)
        self.num_heads = list(num_heads)
        self.window_size = window_size
        self.mlp_ratio = 4.0
        self.qkv_bias = True
        self.qk_scale = None
        self.drop_rate = 0.
        self.attn_drop_rate = 0.
        self.drop_path_rate = drop_path_rate
        self.patch_norm = True
        self.num_classes = num_classes
        self.in_channels = embed_dim * 2 ** (len(self.depths) - 1)
        self.spatial_type = 'avg'
        self.dropout_ratio = 0.5
        self.score_type = 'score'
        self.backbone = SwinTransformer3D(
            pretrained=self.pretrained, pretrained2d=self.pretrained2d, patch_size=self.patch_size,
            in_chans=self.in_chans, embed_dim=self.embed_dim, depths=self.depths, num_heads=self.num_heads,
            window_size=self.window_size, mlp_ratio=self.mlp_ratio, qkv_bias=self.qkv_bias, qk_scale=self.qk_scale,
            drop_rate=self.drop_rate, attn_drop_rate=self.attn_drop_rate, drop_path_rate=self.drop_path_rate,
            patch_norm=self.patch_norm)
        self.cls_head = I3DHead(num_classes=self.num_classes, in_channels=self.in_channels,
                                spatial_type=self.spatial_type, dropout_ratio=self.dropout_ratio)

    def forward(self, x):
        """x: (batch, n_views, C, T, H, W)."""
        n_views = x.shape[1]
        x = x.reshape((-1,) + x.shape[2:])
        feat = self.backbone(x)
        cls_score = self.cls_head(feat)
        return self.average_clips(cls_score, num_segs=n_views)

    def average_clips(self, cls_score, num_segs=1):
        bz = cls_score.shape[0]
        cls_score = cls_score.view(bz // num_segs, num_segs, -1)
        if self.score_type == 'prob':
            return F.softmax(cls_score, dim=2).mean(dim=2)
        if self.score_type == 'score':
            return cls_score.mean(dim=1), cls_score
        raise NotImplementedError
