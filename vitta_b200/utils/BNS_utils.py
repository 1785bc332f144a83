"""Layer selection and the BNS-variant hook -- mirror of the reference's ``utils/BNS_utils.py``."""
import torch
import torch.nn as nn

from .. import ops
from .._lib import VittaError
from . import norm_stats_utils as nsu


def choose_layers(model, candidate_layers):
    """[(name, module)] of every module that is an instance of one of ``candidate_layers``, in
    ``named_modules()`` order -- the order that indexes the source-statistics lists (reference :245-259)."""
    kinds = tuple(candidate_layers)
    return [(name, mod) for name, mod in model.named_modules() if isinstance(mod, kinds)]


def freeze_except_bn(model, bn_condidiate_layers):
    """train() the model, freeze everything, re-enable the given norm layer types (reference :262-276)."""
    model.train()
    model.requires_grad_(False)
    kinds = tuple(bn_condidiate_layers)
    for mod in model.modules():
        if isinstance(mod, kinds):
            mod.requires_grad_(True)
    return model


def collect_bn_params(model, bn_candidate_layers):
    """Affine parameters (weight, bias) of the given norm layer types and their names (reference :278-288)."""
    kinds = tuple(bn_candidate_layers)
    params, names = [], []
    for mod_name, mod in model.named_modules():
        if isinstance(mod, kinds):
            for leaf, p in mod.named_parameters():
                if leaf in ('weight', 'bias'):
                    params.append(p)
                    names.append(f"{mod_name}.{leaf}")
    return params, names


class BNFeatureHook(nsu._TapBase):
    """``--stat_reg BNS`` hook (reference :19-77): statistics of the BatchNorm *input* per frame batch vs the
    layer's running statistics (snapshotted at construction), optional EMA from zeros."""

    def __init__(self, module, reg_type='l2norm', running_manner=False, use_src_stat_in_reg=True, momentum=0.1):
        if not isinstance(module, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
            raise VittaError("BNFeatureHook needs a BatchNorm module")
        self.reg_type = reg_type
        self.running_manner = running_manner
        self.use_src_stat_in_reg = use_src_stat_in_reg
        self.momentum = momentum
        self.before_norm = True
        self.source_mean = module.running_mean.data.clone()
        self.source_var = module.running_var.data.clone()
        self._arena = nsu._align_gen.arena_for_new_hook()
        # running_manner False: the "meter" is the batch statistic itself (w_new = 1, w_old = 0)
        self._layer = self._arena.add_layer(module.num_features, self.source_mean, self.source_var, reg_type, True,
                                            momentum if running_manner else 1.0)
        nsu._align_gen.hooks.append(self)
        self._attach(module)

    def hook_fn(self, module, input, output):
        x = input[0]
        if not x.is_contiguous():
            x = x.contiguous()
        c = x.shape[1]
        if x.dim() == 2:                       # (B, C): TAM G branch, reference :43-45
            O, I = x.shape[0], 1
        else:                                  # (B, C, ...) : reduce over everything but C
            O, I = x.shape[0], x[0, 0].numel()
        ly, arena = self._layer, self._arena
        if torch.is_grad_enabled() and x.requires_grad:
            ly.token = ops.StatsTapFn.apply(x, x, arena, ly, O, c, I, 1, None, None)
        else:
            ly.token = None
            arena.record(ly, x, O, c, I, 1)
        if not self.use_src_stat_in_reg:
            # reference :61-62: the target is the layer's running statistics as they are NOW (they move when the BN layer
            # is in train mode, i.e. --fix_BNS False); with use_src_stat_in_reg they stay the construction-time snapshot
            arena.set_source(ly, module.running_mean, module.running_var)

    @property
    def r_feature(self):
        return self._arena.layer_loss(self._layer)

    @property
    def mean(self):
        self._arena.finalize()
        return self._arena.vec(self._arena.ema_mean, self._layer)

    @property
    def var(self):
        self._arena.finalize()
        return self._arena.vec(self._arena.ema_var, self._layer)
