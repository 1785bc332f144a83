"""Feature-statistics hooks -- host-side mirror of the reference's ``utils/norm_stats_utils.py``.

Same class names, constructor signatures and attributes (``r_feature``, ``batch_mean``/``batch_var``,
``close()``, ``add_hook_back(module)``) as the reference, so ``corpus.basics.tta_standard`` drives them
unchanged (reference corpus/basics.py:575-586,660-661,683-684,721-727).  The arithmetic is not here: the
hooks only describe the feature layout to the sm_100a kernels (K1 statistics partials, K2 merge + EMA +
loss + backward coefficients, K3/K4 closed-form backward) via :class:`vitta_b200.ops.StatsArena`.

Two attachment modes
  * fused: the module is one of ours (``StatsBatchNorm2d`` / ``StatsLayerNorm`` inside the vitta_b200
    models).  The hook registers itself as the module's *tap*; the model's fused norm kernel emits the
    statistics partials while it writes the activation, and its backward kernel adds a_c + b_c*y.
  * generic: any ``nn.BatchNorm2d/3d`` / ``nn.LayerNorm`` (e.g. the reference's own model classes).  A
    regular forward hook runs the standalone K1 kernel on the output and K3 in backward.
"""
import torch
import torch.nn as nn

from .. import ops
from .._lib import VittaError


# ----------------------------------------------------------------------------------------------
# arena generations: hooks built by independent constructor calls share one StatsArena per "generation"
# ----------------------------------------------------------------------------------------------
class _Generation:
    def __init__(self):
        self.arena = None
        self.hooks = []

    def arena_for_new_hook(self):
        """A new generation starts when a hook is constructed after the previous generation has already
        run (finalized at least once) and all of its hooks are closed -- the per-sample re-initialisation of
        ``tta_standard`` mode (reference corpus/basics.py:519-530)."""
        if self.arena is None or (self.arena.finalize_calls > 0 and all(h._closed for h in self.hooks)):
            self.arena = ops.StatsArena(process_group=_process_group)
            self.hooks = []
        return self.arena


_align_gen = _Generation()
_zero_scalars = {}     # device -> shared constant 0-dim zero (never written)
_stat_gen = _Generation()
_process_group = None


def set_process_group(pg):
    """Multi-GPU: statistics of hooks created from now on are merged over this group (collective C1)."""
    global _process_group
    _process_group = pg


def reset_arenas():
    global _align_gen, _stat_gen
    _align_gen, _stat_gen = _Generation(), _Generation()


def new_align_generation():
    """The alignment hooks constructed from now on get an arena of their own.  ``OnlineAdapter`` calls this before it
    builds its hooks, so that two adapters alive at the same time (bench.py runs several workloads in one process; a
    sharded and an unsharded adapter are compared in the multi-GPU parity check) never share statistics buffers --
    the implicit rule of :class:`_Generation` only separates generations whose hooks were closed."""
    global _align_gen
    _align_gen = _Generation()


def _is_fused_module(module):
    return getattr(module, "_vitta_fused", False)


def describe_feature(module, feature, clip_len):
    """-> (tensor, O, C, I, n_clips): the (outer, channel, inner) description of include/vitta_b200.h K1.
    Mirrors the reshapes of the reference (norm_stats_utils.py:188-197, 208-236) without moving any data."""
    if isinstance(module, nn.BatchNorm2d):
        if feature.dim() != 4:
            raise VittaError("BatchNorm2d feature must be (N*T, C, H, W)")
        nt, c, h, w = feature.shape
        if clip_len is None or nt % clip_len != 0:
            raise VittaError("BatchNorm2d feature: N*T=%d is not a multiple of clip_len=%r" % (nt, clip_len))
        n = nt // clip_len
        if h * w > 1 and c > 1 and feature.is_contiguous(memory_format=torch.channels_last):
            return feature, nt * h * w, c, 1, n
        if not feature.is_contiguous():
            feature = feature.contiguous()
        return feature, nt, c, h * w, n
    if isinstance(module, nn.BatchNorm3d):
        if feature.dim() != 5:
            raise VittaError("BatchNorm3d feature must be (N, C, T, H, W)")
        n, c, t, h, w = feature.shape
        if c > 1 and t * h * w > 1 and feature.is_contiguous(memory_format=torch.channels_last_3d):
            return feature, n * t * h * w, c, 1, n
        if not feature.is_contiguous():
            feature = feature.contiguous()
        return feature, n, c, t * h * w, n
    if isinstance(module, nn.LayerNorm):
        assert feature.dim() == 5, "LayerNorm feature must be (B, D, H, W, C)"  # reference :222,231
        if not feature.is_contiguous():
            feature = feature.contiguous()
        c = feature.shape[-1]
        return feature, feature.numel() // c, c, 1, feature.shape[0]
    raise Exception(f'undefined module {module}')


class _TapBase:
    """Attachment logic shared by all hook classes."""

    def _attach(self, module):
        self._closed = False
        if _is_fused_module(module) and not getattr(self, "before_norm", False):
            module._vitta_tap = self
            self._fused_module = module
            self.hook = None
        else:
            self._fused_module = None
            self.hook = module.register_forward_hook(self.hook_fn)

    def close(self):
        if self._fused_module is not None:
            if getattr(self._fused_module, "_vitta_tap", None) is self:
                self._fused_module._vitta_tap = None
        elif self.hook is not None:
            self.hook.remove()
        self._closed = True

    def add_hook_back(self, module):
        self._attach(module)


class CombineNormStatsRegHook_onereg(_TapBase):
    """Alignment hook actually used by ViTTA (reference utils/norm_stats_utils.py:103-258).

    Per forward: mu_c, sigma^2_c over (N', T, H, W) -> meter update (EMA from 0 with detached history, or
    running weighted mean) -> ``r_feature`` = reg(source, meter.avg).  Multiple views are simply part of
    the N' axis ("onereg")."""

    def __init__(self, module, clip_len=None, spatiotemp_stats_clean_tuple=None, reg_type='mse_loss', moving_avg=None,
                 momentum=0.1, stat_type_list=None, reduce_dim=True, before_norm=None, if_sample_tta_aug_views=None,
                 n_augmented_views=None):
        assert stat_type_list == ['spatiotemp']          # reference :131
        self.clip_len = clip_len
        self.reg_type = reg_type
        self.moving_avg = moving_avg
        self.momentum = momentum
        self.stat_type_list = stat_type_list
        self.reduce_dim = reduce_dim
        self.before_norm = before_norm
        self.if_sample_tta_aug_views = if_sample_tta_aug_views
        self.n_augmented_views = n_augmented_views
        self.source_mean_spatiotemp, self.source_var_spatiotemp = spatiotemp_stats_clean_tuple
        self._is_bn1d = isinstance(module, nn.BatchNorm1d)
        self._layer = None
        self._arena = None
        self._zero = None
        if not self._is_bn1d:
            if self.source_mean_spatiotemp is None:
                raise VittaError("source statistics are required for BatchNorm2d/3d and LayerNorm hooks")
            channels = int(torch.as_tensor(self.source_mean_spatiotemp).numel())
            self._arena = _align_gen.arena_for_new_hook()
            self._layer = self._arena.add_layer(channels, self.source_mean_spatiotemp, self.source_var_spatiotemp,
                                                reg_type, bool(moving_avg), momentum)
            _align_gen.hooks.append(self)
        self._attach(module)

    # fused modules call this to learn where to put their partials
    def tap_target(self):
        return self._arena, self._layer

    def note_batch(self, n_clips):
        self._layer.n_batch = int(n_clips)

    def mark_bn1d_fired(self, device):
        """Called by the fused TAM gate kernels instead of running the BatchNorm1d module: this hook contributes an exact
        zero for stat_type_list == ['spatiotemp'] (reference :158-183), so all it needs is to have 'seen' a forward."""
        z = _zero_scalars.get(device)
        if z is None:
            z = _zero_scalars[device] = torch.zeros((), dtype=torch.float32, device=device)
        self._zero = z

    def hook_fn(self, module, input, output):
        feature = input[0] if self.before_norm else output
        if self._is_bn1d:
            # reference :158-183: BatchNorm1d contributes nothing for stat_type_list == ['spatiotemp']
            self._zero = torch.zeros((), dtype=torch.float32, device=feature.device)
            return
        feat, O, C, I, n_clips = describe_feature(module, feature, self.clip_len)
        self._layer.n_batch = n_clips
        ly, arena = self._layer, self._arena
        if torch.is_grad_enabled() and feature.requires_grad:
            saved, yscale, yshift = feat, None, None
            if (not self.before_norm and isinstance(module, (nn.BatchNorm2d, nn.BatchNorm3d)) and not module.training
                    and input[0].shape == feat.shape and input[0].stride() == feat.stride()):
                # eval-mode BN: recompute y from the BN input in backward (in-place ReLU safe)
                with torch.no_grad():
                    yscale = (module.weight * torch.rsqrt(module.running_var + module.eps)).contiguous()
                    yshift = (module.bias - module.running_mean * yscale).contiguous()
                saved = input[0]
            ly.token = ops.StatsTapFn.apply(feat, saved, arena, ly, O, C, I, 1, yscale, yshift)
        else:
            ly.token = None
            arena.record(ly, feat, O, C, I, 1)

    @property
    def r_feature(self):
        if self._is_bn1d:
            if self._zero is None:
                raise AttributeError("r_feature is only available after a forward pass")
            return self._zero
        if self._layer.geom_key is None:
            raise AttributeError("r_feature is only available after a forward pass")
        return self._arena.layer_loss(self._layer)

    # read-outs used by tests / logging (the reference keeps these inside its meter objects)
    @property
    def ema_mean(self):
        self._arena.finalize()
        return self._arena.vec(self._arena.ema_mean, self._layer)

    @property
    def ema_var(self):
        self._arena.finalize()
        return self._arena.vec(self._arena.ema_var, self._layer)

    @property
    def batch_mean(self):
        self._arena.finalize()
        return self._arena.vec(self._arena.batch_mean, self._layer)

    @property
    def batch_var(self):
        self._arena.finalize()
        return self._arena.vec(self._arena.batch_var, self._layer)


class ComputeNormStatsHook(_TapBase):
    """Source-statistics collection hook (reference utils/norm_stats_utils.py:18-101): per forward exposes
    ``batch_mean`` / ``batch_var``.  'spatiotemp' (the only type the shipped scripts use) runs fused or via
    the standalone K1 kernel; 'spatial' / 'temp' / 'temp_v2' re-describe the layout for the same kernel."""

    def __init__(self, module, clip_len=None, stat_type=None, before_norm=None, batch_size=None):
        self.clip_len = clip_len
        self.stat_type = stat_type
        self.before_norm = before_norm
        self.batch_size = batch_size
        self._is_bn1d = isinstance(module, nn.BatchNorm1d)
        self._arena = None
        self._layer = None
        self._result = None
        self._post = None
        self._module_for_c = module
        self._attach_stat(module)

    def _attach_stat(self, module):
        if self.stat_type == 'spatiotemp' and not self._is_bn1d:
            self._attach(module)
        else:
            self._closed = False
            self._fused_module = None
            self.hook = module.register_forward_hook(self.hook_fn)

    def _ensure_layer(self, channels, tag="main"):
        if not hasattr(self, "_layers"):
            self._layers = {}
        key = (tag, channels)
        if key not in self._layers:
            if self._arena is None:
                self._arena = _stat_gen.arena_for_new_hook()
                _stat_gen.hooks.append(self)
            self._layers[key] = self._arena.add_layer(channels, None, None, 'l1_loss', True, 0.0)
        self._layer = self._layers[key]

    def tap_target(self):
        ch = self._module_for_c.num_features if hasattr(self._module_for_c, "num_features") else \
            self._module_for_c.normalized_shape[-1]
        self._ensure_layer(int(ch))
        self._post = None
        return self._arena, self._layer

    def note_batch(self, n_clips):
        pass

    def hook_fn(self, module, input, output):
        feature = input[0] if self.before_norm else output
        feature = feature.detach()
        if self._is_bn1d:
            # reference :31-53 (temp statistics on the TAM's BatchNorm1d outputs)
            assert self.stat_type in ['temp', 'temp_v2']
            f = feature.contiguous()
            if f.dim() == 2:       # (N*C, T) -> per-t statistics over N*C
                self._run(f, f.shape[0], f.shape[1], 1, None)
            else:                  # (N, C, T) -> per-c statistics over (N, T)
                self._run(f, f.shape[0], f.shape[1], f.shape[2], None)
            return
        if self.stat_type == 'spatiotemp':
            feat, O, C, I, _ = describe_feature(module, feature, self.clip_len)
            self._run(feat, O, C, I, None)
            return
        # remaining stat types: bring the feature to (N, T, C, H, W) order once, then re-describe
        if isinstance(module, nn.BatchNorm2d):
            nt, c, h, w = feature.shape
            t = self.clip_len
            x = feature.reshape(nt // t, t, c, h, w)                       # n t c h w (logical)
        elif isinstance(module, nn.BatchNorm3d):
            x = feature.permute(0, 2, 1, 3, 4)                             # n c t h w -> n t c h w
        elif isinstance(module, nn.LayerNorm):
            assert feature.dim() == 5
            x = feature.permute(0, 1, 4, 2, 3)                             # b t h w c -> n t c h w
        else:
            raise Exception(f'undefined module {module}')
        n, t, c, h, w = x.shape
        if self.stat_type == 'spatial':      # (C, T) statistics over (N, H, W)   reference :96-98
            x = x.contiguous()
            self._run(x, n, t * c, h * w, lambda v: v.view(t, c).t().contiguous())
        elif self.stat_type == 'temp':       # (C, H, W) statistics over (N, T)   reference :83-87
            x = x.contiguous()
            self._run(x, n * t, c * h * w, 1, lambda v: v.view(c, h, w))
        elif self.stat_type == 'temp_v2':    # spatial mean first, then (C,) statistics over (N, T)   :88-91
            x = x.contiguous()
            self._run(x, 1, n * t * c, h * w, None, tag="pool")
            pooled = self._arena.vec(self._arena.batch_mean, self._layer_after_finalize()).clone()   # (n*t*c)
            self._run(pooled.view(n * t, c), n * t, c, 1, None)
        else:
            raise VittaError("unknown stat_type %r" % (self.stat_type,))

    def _layer_after_finalize(self):
        self._arena.finalize()
        return self._layer

    def _run(self, feat, O, C, I, post, tag="main"):
        self._ensure_layer(C, tag)
        self._arena.record(self._layer, feat, O, C, I, 1)
        self._post = post

    def _get(self, which):
        if self._layer is None or self._layer.geom_key is None:
            raise AttributeError("statistics are only available after a forward pass")
        self._arena.finalize()
        v = self._arena.vec(getattr(self._arena, which), self._layer)
        return self._post(v) if self._post is not None else v

    @property
    def batch_mean(self):
        return self._get("batch_mean")

    @property
    def batch_var(self):
        return self._get("batch_var")


def compute_kld(mean_true, mean_pred, var_true, var_pred):
    """KL( N(mean_true, var_true) || N(mean_pred, var_pred) ) summed over channels (reference :8-16).
    Small-vector helper kept for API compatibility; the hooks evaluate it inside the finalize kernel."""
    ratio = torch.log(var_pred / var_true)
    return (0.5 * ratio + (var_true + (mean_true - mean_pred).pow(2)) / (2.0 * var_pred) - 0.5).sum()


def compute_regularization(mean_true, mean_pred, var_true, var_pred, reg_type):
    """Same contract as the reference's free function (:531-542) for callers that hold the vectors
    themselves (C-sized tensors; the hooks do this inside K2 instead)."""
    mean_pred = mean_pred.to(mean_true.device)
    var_pred = var_pred.to(var_true.device)
    if reg_type == 'mse_loss':
        return (var_true - var_pred).pow(2).mean() + (mean_true - mean_pred).pow(2).mean()
    if reg_type == 'l1_loss':
        return (var_true - var_pred).abs().mean() + (mean_true - mean_pred).abs().mean()
    if reg_type == 'kld':
        return compute_kld(mean_true, mean_pred, var_true, var_pred)
    raise VittaError("unknown reg_type %r" % (reg_type,))


class NormStatsRegHook():
    """Deprecated in the reference as well (its constructor raises, :545-552)."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError('args.stat_type of str  is deprecated, use list instead.')
