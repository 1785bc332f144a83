"""Temporal view sampling + on-GPU frame preprocessing -- the step just before the adaptation hot path (SURVEY.md
section 8f rank 3).

``sample_tta_view_indices`` mirrors the deterministic sampling styles of the reference's
``models/tanet_models/video_dataset.py::_sample_tta_augmented_views`` (:159-196; ``uniform_equidist`` is the one ViTTA
uses, utils/opts.py ``--tta_view_sample_style_list``) including the final ``+1`` and the clamp to the last frame (:328).
``views_to_device`` turns decoded uint8 frames into the loader tensors of ``corpus.basics`` with one kernel
(``vitta_gather_normalize_u8``) instead of the PIL / numpy pipeline.  Decoding and resizing stay outside (decord /
PIL are out of scope): frames must already be at the target scale.
"""
import ctypes as C

import numpy as np
import torch

from .. import _lib, synth
from .._lib import call, ptr, stream_ptr

DETERMINISTIC_STYLES = ("uniform", "dense", "uniform_equidist", "dense_equidist")


def sample_tta_view_indices(num_frames, num_segments, n_views=2, style="uniform_equidist", new_length=1):
    """Frame indices (0-based into the decoded video) of all views, concatenated view after view: (n_views * T,)."""
    t = int(num_segments)
    if style == "uniform":            # middle frame of each of T equal segments, one view
        tick = (num_frames - new_length + 1) / float(t)
        offs = [int(tick / 2.0 + tick * x) for x in range(t)]
    elif style == "dense":            # T frames with stride 64 // T from the centre of the video, one view
        stride = 64 // t
        pos = max(1, 1 + num_frames - stride * t)
        start = pos // 2
        offs = [(i * stride + start) % num_frames for i in range(t)]
    elif style == "uniform_equidist":  # n_views equidistant phases inside the first segment
        tick = (num_frames - new_length + 1) / float(t)
        starts = np.linspace(0, tick - 1, num=n_views, dtype=int).tolist()
        offs = [int(s + tick * x) % num_frames for s in starts for x in range(t)]
    elif style == "dense_equidist":
        stride = 64 // t
        pos = max(1, 1 + num_frames - stride * t)
        starts = np.linspace(0, pos - 1, num=n_views, dtype=int).tolist()
        offs = [(i * stride + s) % num_frames for s in starts for i in range(t)]
    else:
        raise NotImplementedError("style %r: only the deterministic styles %s are mirrored" % (style, DETERMINISTIC_STYLES))
    idx = np.asarray(offs, dtype=np.int64) + 1                  # the reference's 1-based offsets ...
    return np.minimum(idx, num_frames - 1)                      # ... used as 0-based indices, clamped (:328)


def views_to_device(frames_u8, indices, clip_len, arch="tanet", crop=None, mean=synth.INPUT_MEAN, std=synth.INPUT_STD):
    """frames_u8: (F, H, W, 3) uint8 CUDA tensor; indices: (V*T,) ints.  Returns the loader tensor of ONE video:
    TANet ``(V*T*3, h, w)`` or Swin ``(V, 3, T, h, w)``, normalised fp32.  crop = (y, x, h, w) or None (whole frame)."""
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise _lib.VittaError("views_to_device: frames must be a (F, H, W, 3) uint8 CUDA tensor; there is no CPU path")
    frames_u8 = frames_u8.contiguous()
    f, h, w, _ = frames_u8.shape
    y0, x0, oh, ow = crop if crop is not None else (0, 0, h, w)
    idx = torch.as_tensor(np.asarray(indices, dtype=np.int32)).to(frames_u8.device)
    n = idx.numel()
    v = n // clip_len
    layout = 0 if arch == "tanet" else 1
    out = torch.empty((n * 3, oh, ow) if layout == 0 else (v, 3, clip_len, oh, ow), dtype=torch.float32,
                      device=frames_u8.device)
    m3 = (C.c_float * 3)(*[float(x) for x in mean])
    s3 = (C.c_float * 3)(*[float(x) for x in std])
    call("vitta_gather_normalize_u8", ptr(frames_u8), f, h, w, ptr(idx), n, int(y0), int(x0), int(oh), int(ow), m3, s3,
         layout, int(clip_len), ptr(out), stream_ptr())
    return out
