"""TSN wrapper around TANet-R50 -- mirror of the reference's ``models/tanet_models/tanet.py``.

Same constructor arguments, attribute names and state-dict keys (``base_model.*``, ``new_fc.*``).  The ResNet-50
trunk is built here (the reference pulls torchvision's and its ImageNet weights, tanet.py:129, which needs a
network); activations are kept channels-last and every norm/activation/TAM step is a fused sm_100a kernel."""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.init import constant_, normal_

from ...nn import StatsBatchNorm2d, conv2d, norm_act, stem
from .basic_ops import ConsensusModule
from .temporal_module import Bottleneck, TemporalBottleneck, make_temporal_modeling


class ResNet50Trunk(nn.Module):
    """conv1/bn1/relu/maxpool/layer1-4/avgpool/fc with torchvision's names and registration order."""

    def __init__(self, layers=(3, 4, 6, 3), num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = StatsBatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._stage(64, layers[0], 1)
        self.layer2 = self._stage(128, layers[1], 2)
        self.layer3 = self._stage(256, layers[2], 2)
        self.layer4 = self._stage(512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(512 * Bottleneck.expansion, num_classes)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
        self.n_segment = None

    def _stage(self, planes, blocks, stride):
        ds = None
        if stride != 1 or self.inplanes != planes * Bottleneck.expansion:
            ds = nn.Sequential(nn.Conv2d(self.inplanes, planes * Bottleneck.expansion, 1, stride, bias=False),
                               StatsBatchNorm2d(planes * Bottleneck.expansion))
        blks = [Bottleneck(self.inplanes, planes, stride, ds)]
        self.inplanes = planes * Bottleneck.expansion
        blks += [Bottleneck(self.inplanes, planes) for _ in range(1, blocks)]
        return nn.Sequential(*blks)

    def _inference_folds(self, x):
        """OPT-IN (VITTA_INFER_FOLD=1; measured on the B200 the folded forward is still 1.2 ms SLOWER than the layer-by-layer
        one at the benchmark shape, 10.2 vs 9.0 ms: the bias / ReLU / range work it adds to the GEMM epilogue sits on the
        critical path of the store-bound layer1 / layer2 convolutions -- DESIGN.md section 9).
        The BN-folded operand set when this forward may use it: no autograd (the per-step evaluation / source-only
        validation), the fp16 operand split, every BatchNorm2d of the residual stages in eval mode without statistics
        taps or foreign hooks, and plain bias-free convolutions.  None otherwise (the layer-by-layer path runs)."""
        from ... import ops
        from ...nn import _fusable
        if torch.is_grad_enabled() or not x.is_cuda or ops.gemm_precision() != "f16x3" or \
                os.environ.get("VITTA_INFER_FOLD", "0") != "1":
            return None
        blocks = [b for st in (self.layer1, self.layer2, self.layer3, self.layer4) for b in st]
        if not all(isinstance(b, TemporalBottleneck) for b in blocks):
            return None
        pairs = [p for b in blocks for p in b.infer_pairs()]
        for conv, bn in pairs:
            if not (_fusable(bn) and bn._vitta_tap is None and conv.bias is None and conv.groups == 1
                    and conv.dilation == (1, 1) and conv.padding_mode == 'zeros' and not conv._forward_hooks
                    and not conv._forward_pre_hooks and (conv.in_channels * conv.kernel_size[0] * conv.kernel_size[1]) % 8 == 0
                    and conv.stride[0] == conv.stride[1] and conv.padding[0] == conv.padding[1]
                    and conv.kernel_size[0] == conv.kernel_size[1] and conv.weight.is_contiguous()):
                return None
        f = getattr(self, "_folds", None)
        if f is None or list(f.slot) != [id(c) for c, _ in pairs] or f.bias.device != x.device:   # (deepcopy keeps old ids)
            f = ops.FoldedConvs(pairs)
            object.__setattr__(self, "_folds", f)        # plain attribute: not a module / parameter / buffer
        f.refresh()
        return f

    def forward(self, x):
        t = self.n_segment
        x = stem(self.conv1, self.bn1, self.maxpool, x, t)
        pooled = None
        stages = (self.layer1, self.layer2, self.layer3, self.layer4)
        folds = self._inference_folds(x)
        for si, stage in enumerate(stages):
            n = len(stage)
            for bi, blk in enumerate(stage):
                last = si == len(stages) - 1 and bi == n - 1
                if isinstance(blk, TemporalBottleneck):
                    if folds is not None:
                        x, pooled = blk.forward_infer(x, folds, want_pool=last)
                    else:
                        x, pooled = blk(x, want_pool=last)
                else:
                    raise NotImplementedError("TSN(tam=False) trunk is not part of the ViTTA path")
        if pooled is None:
            pooled = self.avgpool(x).flatten(1)
        return self.fc(pooled)


class TSN(nn.Module):
    """Reference signature: TSN(num_class, num_segments, modality, base_model='resnet101', new_length=None,
    consensus_type='avg', before_softmax=True, dropout=0.8, img_feature_dim=256, crop_num=1, partial_bn=True,
    print_spec=True, pretrain='imagenet', tam=False, fc_lr5=False, non_local=False)  (tanet.py:16-33)."""

    def __init__(self, num_class, num_segments, modality, base_model='resnet101', new_length=None, consensus_type='avg',
                 before_softmax=True, dropout=0.8, img_feature_dim=256, crop_num=1, partial_bn=True, print_spec=True,
                 pretrain='imagenet', tam=False, fc_lr5=False, non_local=False):
        super().__init__()
        if modality != 'RGB':
            raise NotImplementedError("only the RGB modality is on the ViTTA path (utils/opts.py:16)")
        if base_model != 'resnet50' or not tam or non_local:
            raise NotImplementedError("ViTTA builds TSN(base_model='resnet50', tam=True, non_local=False) "
                                      "(corpus/basics.py:1463-1474)")
        if not before_softmax and consensus_type != 'avg':
            raise ValueError("Only avg consensus can be used after Softmax")
        self.modality = modality
        self.num_segments = num_segments
        self.reshape = True
        self.before_softmax = before_softmax
        self.dropout = dropout
        self.crop_num = crop_num
        self.consensus_type = consensus_type
        self.img_feature_dim = img_feature_dim
        self.pretrain = pretrain
        self.tam = tam
        self.base_model_name = base_model
        self.fc_lr5 = fc_lr5
        self.non_local = non_local
        self.new_length = 1 if new_length is None else new_length

        self.base_model = ResNet50Trunk()
        make_temporal_modeling(self.base_model, num_segments, t_kernel_size=3, t_stride=1, t_padding=1)
        self.base_model.n_segment = num_segments
        self.base_model.last_layer_name = 'fc'
        self.input_size = 224
        self.input_mean = [0.485, 0.456, 0.406]
        self.input_std = [0.229, 0.224, 0.225]

        feature_dim = self.base_model.fc.in_features
        if self.dropout == 0:
            self.base_model.fc = nn.Linear(feature_dim, num_class)
            self.new_fc = None
            normal_(self.base_model.fc.weight, 0, 0.001)
            constant_(self.base_model.fc.bias, 0)
        else:
            self.base_model.fc = nn.Dropout(p=self.dropout)
            out_dim = self.img_feature_dim if consensus_type in ['TRN', 'TRNmultiscale'] else num_class
            self.new_fc = nn.Linear(feature_dim, out_dim)
            normal_(self.new_fc.weight, 0, 0.001)
            constant_(self.new_fc.bias, 0)
        self.consensus = ConsensusModule(consensus_type)
        if not self.before_softmax:
            self.softmax = nn.Softmax()
        self._enable_pbn = partial_bn

    def train(self, mode=True):
        """Like the reference (tanet.py:182-198): with partial_bn every BatchNorm2d but the first is put in
        eval mode and its affine parameters frozen.  (The reference forgets ``return self``; we return it.)"""
        super().train(mode)
        if self._enable_pbn and mode:
            count = 0
            for m in self.base_model.modules():
                if isinstance(m, nn.BatchNorm2d):
                    count += 1
                    if count >= 2:
                        m.eval()
                        m.weight.requires_grad = False
                        m.bias.requires_grad = False
        return self

    def partialBN(self, enable):
        self._enable_pbn = enable

    def forward(self, input, no_reshape=False):
        """input (N', T*3, H, W) or (N', T, 3, H, W) -> (N', num_class)   (tanet.py:308-333)."""
        if not no_reshape:
            sample_len = 3 * self.new_length
            input = input.view((-1, sample_len) + input.size()[-2:])
        base_out = self.base_model(input)
        if self.dropout > 0:
            base_out = self.new_fc(base_out)
        if not self.before_softmax:
            base_out = self.softmax(base_out)
        base_out = base_out.view((-1, self.num_segments) + base_out.size()[1:])
        return self.consensus(base_out).squeeze(1)
