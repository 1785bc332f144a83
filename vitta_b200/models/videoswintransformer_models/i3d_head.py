"""Classification head -- mirror of the reference's ``models/videoswintransformer_models/i3d_head.py``:
AdaptiveAvgPool3d((1,1,1)) -> Dropout(0.5) -> Linear(in_channels, num_classes)  (reference :58-77)."""
import torch
import torch.nn as nn

from ... import ops_swin


class I3DHead(nn.Module):
    def __init__(self, num_classes, in_channels, spatial_type='avg', dropout_ratio=0.5, init_std=0.01):
        super().__init__()
        self.num_classes = num_classes
        self.in_channels = in_channels
        self.spatial_type = spatial_type
        self.dropout_ratio = dropout_ratio
        self.init_std = init_std
        self.dropout = nn.Dropout(p=self.dropout_ratio) if self.dropout_ratio != 0 else None
        self.fc_cls = nn.Linear(self.in_channels, self.num_classes)
        self.avg_pool = nn.AdaptiveAvgPool3d((1, 1, 1)) if self.spatial_type == 'avg' else None
        self.init_weights()

    def init_weights(self):
        nn.init.normal_(self.fc_cls.weight, 0, self.init_std)
        nn.init.constant_(self.fc_cls.bias, 0)

    def forward(self, x):
        """x: (N, C, D, H, W).  When it is the permuted view of a channels-last token volume (what the vitta_b200
        backbone returns) the pooling is one pass of the frame-mean kernel over the (N*D*H*W, C) rows."""
        if self.avg_pool is not None:
            n, c = x.shape[0], x.shape[1]
            tokens = x.permute(0, 2, 3, 4, 1)
            if x.is_cuda and tokens.is_contiguous() and c % 4 == 0:
                x = ops_swin.FrameMeanFn.apply(tokens.reshape(-1, c), n)
            else:
                x = self.avg_pool(x)
        if self.dropout is not None:
            x = self.dropout(x)
        x = x.view(x.shape[0], -1)
        return self.fc_cls(x)
