"""TEST INFRASTRUCTURE -- loader for the *real* reference (wlin-at/ViTTA) on CPU.

Only usable where ``/root/reference`` exists (the build container).  It is used by
``oracle/make_golden.py`` to pin the oracle restatement (``oracle/vitta_oracle.py``)
against outputs of the unmodified reference code.  Nothing in the product package, the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` imports this file: the GPU box has no
``/root/reference``.

What the shims do:
  * stub packages that are not installed and are only needed at import/init time:
    timm, mmcv, mmaction, decord, tensorboardX  (DropPath is the one runtime op; it is
    restated here as timm 0.6.7 defines it: per-sample Bernoulli(keep)/keep);
  * ``torchvision.models.resnet50(True)`` would download ImageNet weights
    (models/tanet_models/tanet.py:129) -> forced to ``weights=None``;
  * CPU only: ``.cuda()`` becomes a no-op and ``torch.device("cuda:0")`` inside
    utils/norm_stats_utils.py:141 resolves to the CPU;
  * ``sys.argv`` is cleaned before baselines/setup_baseline.py:9 parses it at import time.
"""
import importlib
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("VITTA_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "corpus"))


class _DropPath(nn.Module):
    """timm==0.6.7 ``DropPath`` semantics (requirements.txt:60; call site swin_transformer.py:210)."""

    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1.0 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = x.new_empty(shape).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _install_stubs():
    if "timm" not in sys.modules:
        _mod("timm")
        _mod("timm.models", create_model=lambda *a, **k: None)
        _mod("timm.models.layers", DropPath=_DropPath, trunc_normal_=nn.init.trunc_normal_,
             drop_path=None, to_2tuple=lambda x: (x, x))
        _mod("timm.models.registry", register_model=lambda f: f)
    if "mmcv" not in sys.modules:
        _mod("mmcv")
        _mod("mmcv.runner", load_checkpoint=lambda *a, **k: None)

        def normal_init(module, mean=0, std=1, bias=0):
            nn.init.normal_(module.weight, mean, std)
            if getattr(module, "bias", None) is not None:
                nn.init.constant_(module.bias, bias)

        _mod("mmcv.cnn", normal_init=normal_init)
        _mod("mmcv.fileio", FileClient=object)
        _mod("mmcv.parallel", DataContainer=object)
    if "mmaction" not in sys.modules:
        _mod("mmaction")
        _mod("mmaction.utils", get_root_logger=lambda *a, **k: None)
    if "decord" not in sys.modules:
        _mod("decord", VideoReader=object, cpu=lambda *a, **k: None)
    if "tensorboardX" not in sys.modules:
        class SummaryWriter:  # created at main_eval.py:85, never written by tta_standard
            def __init__(self, *a, **k):
                pass

            def add_scalar(self, *a, **k):
                pass

            def close(self):
                pass
        _mod("tensorboardX", SummaryWriter=SummaryWriter)


class _TorchCpuProxy:
    """Stands in for the ``torch`` name inside reference modules that pin ``cuda:0``."""

    def __getattr__(self, name):
        return getattr(torch, name)

    @staticmethod
    def device(*a, **k):
        return torch.device("cpu")

    @staticmethod
    def tensor(*a, **k):
        return torch.tensor(*a, **k)


_LOADED = {}


def load_reference():
    """Import the reference's modules (CPU) and return them in a dict."""
    if _LOADED:
        return _LOADED
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    # .cuda() -> no-op on this GPU-less box
    torch.Tensor.cuda = lambda self, *a, **k: self
    nn.Module.cuda = lambda self, *a, **k: self
    import torchvision

    _orig_r50 = torchvision.models.resnet50

    def _r50(*a, **k):
        return _orig_r50(weights=None)

    torchvision.models.resnet50 = _r50

    # The reference's top-level package names (utils, models, corpus, ...) would collide
    # with nothing in this repo (ours live under vitta_b200.*), so a plain sys.path entry works.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    argv, sys.argv = sys.argv, [sys.argv[0]]
    try:
        names = [
            "utils.opts", "utils.utils_", "utils.norm_stats_utils", "utils.BNS_utils",
            "utils.pred_consistency_utils", "models.tanet_models.tanet",
            "models.tanet_models.temporal_module", "models.tanet_models.basic_ops",
            "models.videoswintransformer_models.swin_transformer",
            "models.videoswintransformer_models.recognizer3d",
            "models.videoswintransformer_models.i3d_head", "corpus.basics", "corpus.main_eval",
        ]
        for n in names:
            _LOADED[n] = importlib.import_module(n)
    finally:
        sys.argv = argv
    proxy = _TorchCpuProxy()
    _LOADED["utils.norm_stats_utils"].torch = proxy
    return _LOADED


def build_reference_tsn(num_class, num_segments, dropout=0.8):
    """Reference TSN exactly as corpus/basics.py:1463-1474 builds it."""
    ref = load_reference()
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref["models.tanet_models.tanet"].TSN(
            num_class, num_segments, "RGB", base_model="resnet50", consensus_type="avg",
            img_feature_dim=256, tam=True, non_local=False, partial_bn=False, dropout=dropout)
    return model


def build_reference_swin(num_classes, patch_size=(2, 4, 4), window_size=(8, 7, 7), drop_path_rate=0.2,
                         embed_dim=None, depths=None, num_heads=None):
    """Reference Recognizer3D (corpus/basics.py:1489).  ``embed_dim/depths/num_heads`` re-parametrise
    the hard-coded Swin-B (recognizer3d.py:53-55) to e.g. Swin-T or a tiny test model by rebuilding
    the backbone with the reference's own SwinTransformer3D class."""
    ref = load_reference()
    rec = ref["models.videoswintransformer_models.recognizer3d"]
    swin = ref["models.videoswintransformer_models.swin_transformer"]
    head = ref["models.videoswintransformer_models.i3d_head"]
    if embed_dim is None:
        return rec.Recognizer3D(num_classes=num_classes, patch_size=patch_size, window_size=window_size,
                                drop_path_rate=drop_path_rate)
    model = rec.Recognizer3D.__new__(rec.Recognizer3D)
    nn.Module.__init__(model)
    model.score_type = "score"
    model.backbone = swin.SwinTransformer3D(
        pretrained=None, pretrained2d=True, patch_size=patch_size, in_chans=3, embed_dim=embed_dim,
        depths=depths, num_heads=num_heads, window_size=window_size, mlp_ratio=4.0, qkv_bias=True,
        qk_scale=None, drop_rate=0.0, attn_drop_rate=0.0, drop_path_rate=drop_path_rate, patch_norm=True)
    model.cls_head = head.I3DHead(num_classes=num_classes, in_channels=embed_dim * 2 ** (len(depths) - 1),
                                  spatial_type="avg", dropout_ratio=0.5)
    return model
