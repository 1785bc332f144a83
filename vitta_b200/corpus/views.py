"""Temporal view sampling + on-GPU frame preprocessing -- the step just before the adaptation hot path (SURVEY.md
section 8f rank 3).

``sample_tta_view_indices`` mirrors the deterministic sampling styles of the reference's
``models/tanet_models/video_dataset.py::_sample_tta_augmented_views`` (:159-196; ``uniform_equidist`` is the one ViTTA
uses, utils/opts.py ``--tta_view_sample_style_list``) including the final ``+1`` and the clamp to the last frame (:328).
``views_to_device`` turns decoded uint8 frames into the loader tensors of ``corpus.basics`` with one kernel
(``vitta_gather_normalize_u8``) instead of the PIL / numpy pipeline.  Decoding and resizing stay outside (decord /
PIL are out of scope): frames must already be at the target scale.

Spatial side of the views: ``sample_multiscale_crop`` mirrors ``SubgroupWise_MultiScaleCrop_TANet._sample_crop_size``
(models/tanet_models/transforms.py:325-385; one random multi-scale crop box per temporal view, the reference's default with
``--if_spatial_rand_cropping``, utils/opts.py:85, corpus/basics.py:1238-1245) and ``views_to_device(..., boxes=...)`` crops
and resizes on the GPU with Pillow's exact 8-bit bilinear arithmetic (``vitta_gather_crop_resize_normalize_u8``).
"""
import ctypes as C
import random as _random

import numpy as np
import torch

from .. import _lib, synth
from .._lib import call, ptr, stream_ptr

DETERMINISTIC_STYLES = ("uniform", "dense", "uniform_equidist", "dense_equidist")
RANDOM_STYLES = ("uniform_rand", "dense_rand", "random")        # one view per call, drawn from numpy's legacy generator


def sample_tta_view_indices(num_frames, num_segments, n_views=2, style="uniform_equidist", new_length=1, np_rng=np.random):
    """Frame indices (0-based into the decoded video) of all views, concatenated view after view: (n_views * T,).
    The random styles (video_dataset.py:197-229) give ONE view per call (the reference lists a style once per view) and
    draw from ``np_rng`` -- the ``numpy.random`` module like the reference, or a ``RandomState`` -- with the same calls in
    the same order, so ``np.random.seed(s)`` reproduces the reference's indices."""
    t = int(num_segments)
    if style in RANDOM_STYLES:
        if style == "uniform_rand":       # one random frame from each of T equal segments
            avg = (num_frames - new_length + 1) // t
            if avg > 0:
                offs = np.multiply(list(range(t)), avg) + np_rng.randint(avg, size=t)
            elif num_frames > t:          # too short to segment: T sorted draws with replacement
                offs = np.sort(np_rng.randint(num_frames - new_length + 1, size=t))
            else:
                offs = np.zeros((t,))
            idx = np.asarray(offs).astype(np.int64) + 1
        elif style == "dense_rand":       # stride 64 // T from a random start
            stride = 64 // t
            pos = max(1, 1 + num_frames - stride * t)
            start = 0 if pos == 1 else int(np_rng.randint(0, pos - 1))
            idx = np.asarray([(i * stride + start) % num_frames for i in range(t)], dtype=np.int64) + 1
        else:                             # 'random': T distinct frames, sorted -- and NO +1 in the reference (:221-229)
            if num_frames >= t:
                idx = np.sort(np_rng.choice(num_frames, size=t, replace=False)).astype(np.int64)
            else:
                idx = np.asarray(list(range(num_frames)) + [num_frames - 1] * (t - num_frames), dtype=np.int64)
        return np.minimum(idx, num_frames - 1)
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
        raise NotImplementedError("style %r: the reference defines %s" % (style, DETERMINISTIC_STYLES + RANDOM_STYLES))
    idx = np.asarray(offs, dtype=np.int64) + 1                  # the reference's 1-based offsets ...
    return np.minimum(idx, num_frames - 1)                      # ... used as 0-based indices, clamped (:328)


# ----------------------------------------------------------------------------------------------
# per-view random multi-scale crop (the reference's default spatial augmentation of the TTA views)
# ----------------------------------------------------------------------------------------------
MULTISCALE_SCALES = (1, .875, .75, .66)      # transforms.py:291


def fill_fix_offset(image_w, image_h, crop_w, crop_h, more_fix_crop=True):
    """Candidate (offset_w, offset_h) positions in the reference's order (transforms.py:361-385): 5 or 13 of them."""
    ws, hs = (image_w - crop_w) // 4, (image_h - crop_h) // 4
    ret = [(0, 0), (4 * ws, 0), (0, 4 * hs), (4 * ws, 4 * hs), (2 * ws, 2 * hs)]
    if more_fix_crop:
        ret += [(0, 2 * hs), (4 * ws, 2 * hs), (2 * ws, 4 * hs), (2 * ws, 0),
                (ws, hs), (3 * ws, hs), (ws, 3 * hs), (3 * ws, 3 * hs)]
    return ret


def sample_multiscale_crop(image_w, image_h, input_size, rng=_random, scales=MULTISCALE_SCALES, max_distort=1,
                           more_fix_crop=True):
    """One draw of ``_sample_crop_size`` with fix_crop=True (transforms.py:325-354): (crop_w, crop_h, offset_w, offset_h).
    ``rng`` is the ``random`` module or a ``random.Random``; like the reference it is asked for two ``choice`` draws --
    first the (w, h) pair, then the position -- so the same seed gives the same boxes as the reference pipeline."""
    iw, ih = (input_size, input_size) if isinstance(input_size, int) else input_size
    base = min(image_w, image_h)
    sizes = [int(base * x) for x in scales]
    crop_h = [ih if abs(x - ih) < 3 else x for x in sizes]
    crop_w = [iw if abs(x - iw) < 3 else x for x in sizes]
    pairs = [(w, h) for i, h in enumerate(crop_h) for j, w in enumerate(crop_w) if abs(i - j) <= max_distort]
    cw, ch = rng.choice(pairs)
    ow, oh = rng.choice(fill_fix_offset(image_w, image_h, cw, ch, more_fix_crop))
    return cw, ch, ow, oh


def sample_view_crops(image_w, image_h, input_size, n_views, rng=_random):
    """One box per temporal view, drawn view after view (SubgroupWise_MultiScaleCrop_TANet.__call__, transforms.py:301-310)."""
    return [sample_multiscale_crop(image_w, image_h, input_size, rng) for _ in range(n_views)]


def resample_tables(in_size, out_size, in_offset=0, slots=None):
    """Pillow's 8-bit bilinear coefficient tables for one axis (host arithmetic in the library, no GPU involved):
    (bounds (out, 2) int32 = (first source index + in_offset, count), kk (out, slots) int32, 22-bit fixed point)."""
    lib = _lib.load()
    ks = lib.vitta_resample_ksize(int(in_size), int(out_size))
    if ks <= 0:
        raise _lib.VittaError("resample_tables: bad sizes %r -> %r" % (in_size, out_size))
    slots = ks if slots is None else int(slots)
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, slots), np.int32)
    _lib.check(lib.vitta_resample_coeffs_u8(int(in_size), int(out_size), int(in_offset), slots,
                                            bounds.ctypes.data_as(C.POINTER(C.c_int32)),
                                            kk.ctypes.data_as(C.POINTER(C.c_int32))), "vitta_resample_coeffs_u8")
    return bounds, kk


def crop_resize_tables(boxes, out_h, out_w):
    """Stacked per-view tables for ``vitta_gather_crop_resize_normalize_u8``: (hb, hk, vb, vk, slots) numpy int32."""
    lib = _lib.load()
    slots = max(3, max(max(lib.vitta_resample_ksize(int(b[0]), out_w), lib.vitta_resample_ksize(int(b[1]), out_h))
                       for b in boxes))
    hb, hk, vb, vk = [], [], [], []
    for cw, ch, ow, oh in boxes:
        b, k = resample_tables(cw, out_w, ow, slots)
        hb.append(b), hk.append(k)
        b, k = resample_tables(ch, out_h, oh, slots)
        vb.append(b), vk.append(k)
    return np.stack(hb), np.stack(hk), np.stack(vb), np.stack(vk), slots


def scale_center_crop_geometry(image_w, image_h, scale_size, input_size):
    """GroupScale_TANet(scale_size) + GroupCenterCrop_TANet(input_size) (transforms.py:46-52,170-183; torchvision 0.8.2
    ``Resize(int)`` / ``CenterCrop``): smaller edge -> scale_size keeping the aspect ratio (``int(size * long / short)``,
    untouched if it already matches), then the centred S x S window (``int(round((extent - S) / 2.))``).
    Returns (resized_w, resized_h, left, top)."""
    w, h, size = int(image_w), int(image_h), int(scale_size)
    if (w <= h and w == size) or (h <= w and h == size):
        ow, oh = w, h
    elif w < h:
        ow, oh = size, int(size * h / w)
    else:
        ow, oh = int(size * w / h), size
    s = int(input_size)
    if ow < s or oh < s:
        raise _lib.VittaError("scale_center_crop: %dx%d scaled to %dx%d is smaller than the %d crop" % (w, h, ow, oh, s))
    return ow, oh, int(round((ow - s) / 2.)), int(round((oh - s) / 2.))


def scale_center_crop_tables(image_w, image_h, scale_size, input_size, n_views=1):
    """Tables of ``vitta_gather_crop_resize_normalize_u8`` for the scale + centre-crop path: the rows [left, left + S) /
    [top, top + S) of the whole-frame resize (resize-then-crop == computing only those output positions)."""
    ow, oh, left, top = scale_center_crop_geometry(image_w, image_h, scale_size, input_size)
    s = int(input_size)
    hb, hk = resample_tables(image_w, ow)
    vb, vk = resample_tables(image_h, oh)
    slots = max(3, hk.shape[1], vk.shape[1])
    pad = lambda k: np.pad(k, ((0, 0), (0, slots - k.shape[1])))
    rep = lambda a: np.ascontiguousarray(np.broadcast_to(a, (n_views,) + a.shape))
    return (rep(hb[left:left + s]), rep(pad(hk)[left:left + s]), rep(vb[top:top + s]), rep(pad(vk)[top:top + s]), slots)


def full_res_sample_tables(image_w, image_h, scale_size, input_size, n_clips=1):
    """GroupFullResSample_TANet(input_size, scale_size, flip=False) (transforms.py:227-272; ``--test_crops 3``,
    corpus/basics.py:1264-1265): scale the smaller edge to scale_size, then THREE S x S crops at (0, 2h'), (4w', 2h'),
    (2w', 2h') with w' = (W - S) // 4, h' = (H - S) // 4 -- left / right / centre.  The reference emits all frames of crop
    0, then of crop 1, then of crop 2, so the tables come as 3 * n_clips "views": view j uses crop j // n_clips."""
    ow, oh, _, _ = scale_center_crop_geometry(image_w, image_h, scale_size, input_size)
    s = int(input_size)
    ws, hs = (ow - s) // 4, (oh - s) // 4
    hb, hk = resample_tables(image_w, ow)
    vb, vk = resample_tables(image_h, oh)
    slots = max(3, hk.shape[1], vk.shape[1])
    pad = lambda k: np.pad(k, ((0, 0), (0, slots - k.shape[1])))
    hk, vk = pad(hk), pad(vk)
    out = [[], [], [], []]
    for left, top in ((0, 2 * hs), (4 * ws, 2 * hs), (2 * ws, 2 * hs)):
        for _ in range(n_clips):
            for lst, a in zip(out, (hb[left:left + s], hk[left:left + s], vb[top:top + s], vk[top:top + s])):
                lst.append(a)
    return tuple(np.ascontiguousarray(np.stack(a)) for a in out) + (slots,)


def test_clip_indices(num_frames, num_segments, sample_style="uniform-1"):
    """``Video_TANetDataSet._get_test_indices`` (video_dataset.py:270-303) for ``--sample_style`` 'uniform-N' / 'dense-N':
    the same formulas as the deterministic TTA styles -- N = 1: 'uniform' / 'dense', N > 1: the equidistant variants with N
    clips -- followed by the loader's clamp to the last frame (:328)."""
    kind, _, n = str(sample_style).partition('-')
    n = int(n or 1)
    if kind not in ("uniform", "dense"):
        raise NotImplementedError("{} not exist".format(sample_style))       # the reference's message (:303)
    return sample_tta_view_indices(num_frames, num_segments, n, kind if n == 1 else kind + "_equidist")


test_clip_indices.__test__ = False      # not a pytest test, whatever module imports it


def swin_seq_frames(num_frames, clip_len):
    """Frame indices of the Video-Swin loader's clean evaluation clip (``SampleFrames.get_seq_frames`` in test mode,
    models/videoswintransformer_models/transforms_backup.py:548-569; ``--frame_uniform``, its default): the middle frame of
    each of ``clip_len`` segments of ``(num_frames - 1) / clip_len`` frames, segment borders rounded half to even
    (``np.round``), then clamped to the last frame (:690)."""
    seg = float(num_frames - 1) / clip_len
    seq = [(int(np.round(seg * i)) + int(np.round(seg * (i + 1)))) // 2 for i in range(clip_len)]
    return np.minimum(np.asarray(seq, dtype=np.int64), num_frames - 1)


# ---- host-side geometry of the Video-Swin loader (models/videoswintransformer_models/video_dataset.py:66-101); its
# ---- OpenCV resize runs on the device through swin_views_to_device below (K14)
def swin_rescale_size(image_w, image_h, short_edge):
    """``Resize(scale=(-1, short_edge))`` (transforms_backup.py:772-790,834-835): mmcv.rescale_size with the long edge
    unbounded -- factor = short_edge / min(h, w), new size = int(x * factor + 0.5).  mmcv (pinned 1.3.12,
    requirements.txt:25) is neither part of the reference tree nor installed: its published rule is restated, parity unpinned."""
    factor = float(short_edge) / min(image_h, image_w)
    return int(image_w * factor + 0.5), int(image_h * factor + 0.5)


def swin_center_crop_box(image_w, image_h, crop_size):
    """``CenterCrop`` (transforms_backup.py:897-915): (left, top, right, bottom) with floor-halved margins."""
    left, top = (image_w - crop_size) // 2, (image_h - crop_size) // 2
    return left, top, left + crop_size, top + crop_size


def swin_random_resized_crop_bbox(image_h, image_w, area_range=(0.08, 1.0), aspect_ratio_range=(3 / 4, 4 / 3),
                                  max_attempts=10, np_rng=np.random, py_rng=_random):
    """``RandomResizedCrop.get_crop_bbox`` (transforms_backup.py:223-272): ten log-uniform aspect ratios and ten uniform
    areas from numpy's generator, the first candidate that fits placed with two ``random.randint`` draws, else the centred
    square.  One box per VIDEO (all frames of all views share it, :274-281).  Same calls in the same order as the reference,
    so seeding both generators reproduces its boxes.  Returns (left, top, right, bottom)."""
    area = image_h * image_w
    min_ar, max_ar = aspect_ratio_range
    ratios = np.exp(np_rng.uniform(np.log(min_ar), np.log(max_ar), size=max_attempts))
    areas = np_rng.uniform(*area_range, size=max_attempts) * area
    cand_w = np.round(np.sqrt(areas * ratios)).astype(np.int32)
    cand_h = np.round(np.sqrt(areas / ratios)).astype(np.int32)
    for i in range(max_attempts):
        cw, ch = int(cand_w[i]), int(cand_h[i])
        if ch <= image_h and cw <= image_w:
            x = py_rng.randint(0, image_w - cw)
            y = py_rng.randint(0, image_h - ch)
            return x, y, x + cw, y + ch
    size = min(image_h, image_w)
    x, y = (image_w - size) // 2, (image_h - size) // 2
    return x, y, x + size, y + size


def views_to_device(frames_u8, indices, clip_len, arch="tanet", crop=None, mean=synth.INPUT_MEAN, std=synth.INPUT_STD,
                    boxes=None, out_size=None, scale_size=None, three_crops=False):
    """frames_u8: (F, H, W, 3) uint8 CUDA tensor; indices: (V*T,) ints.  Returns the loader tensor of ONE video:
    TANet ``(V*T*3, h, w)`` or Swin ``(V, 3, T, h, w)``, normalised fp32.  crop = (y, x, h, w) or None (whole frame).
    boxes = one (crop_w, crop_h, offset_w, offset_h) per view (``sample_view_crops``) + out_size = S (or (h, w)): every
    view is cropped with its own box and resized to S x S exactly as PIL's BILINEAR does (the reference's
    SubgroupWise_MultiScaleCrop_TANet); ``crop`` must then be None.
    scale_size = Z + out_size = S (no boxes): the reference's other spatial path, GroupScale(Z) + GroupCenterCrop(S)
    (corpus/basics.py:1259-1263), same kernel, same PIL-exact arithmetic; with three_crops=True the reference's
    ``--test_crops 3`` path (GroupFullResSample_TANet): output planes ordered [crop][clip][frame][rgb]."""
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise _lib.VittaError("views_to_device: frames must be a (F, H, W, 3) uint8 CUDA tensor; there is no CPU path")
    frames_u8 = frames_u8.contiguous()
    f, h, w, _ = frames_u8.shape
    idx = torch.as_tensor(np.asarray(indices, dtype=np.int32)).to(frames_u8.device)
    n = idx.numel()
    v = n // clip_len
    layout = 0 if arch == "tanet" else 1
    if boxes is not None or scale_size is not None:
        if crop is not None or out_size is None or (boxes is not None and scale_size is not None):
            raise _lib.VittaError("views_to_device: boxes / scale_size need out_size and exclude crop and each other")
        if n != v * clip_len or (boxes is not None and len(boxes) != v):
            raise _lib.VittaError("views_to_device: one crop box per view (%d indices, clip length %d)" % (n, clip_len))
        if boxes is not None:
            oh, ow = (out_size, out_size) if isinstance(out_size, int) else out_size
            hb, hk, vb, vk, slots = crop_resize_tables(boxes, oh, ow)
        elif three_crops:
            oh = ow = int(out_size)
            hb, hk, vb, vk, slots = full_res_sample_tables(w, h, scale_size, out_size, v)
            idx = idx.repeat(3)                   # every frame once per crop: [crop][clip][frame]
            n, v = 3 * n, 3 * v
            boxes = [(w, h, 0, 0)] * v
        else:
            oh = ow = int(out_size)
            hb, hk, vb, vk, slots = scale_center_crop_tables(w, h, scale_size, out_size, v)
            boxes = [(w, h, 0, 0)] * v            # the tables address the whole frame
        dev = frames_u8.device
        tabs = [torch.from_numpy(t).to(dev) for t in (hb, hk, vb, vk)]
        out = torch.empty((n * 3, oh, ow) if layout == 0 else (v, 3, clip_len, oh, ow), dtype=torch.float32, device=dev)
        bx = (C.c_int32 * (4 * v))(*[int(x) for b in boxes for x in b])
        m3 = (C.c_float * 3)(*[float(x) for x in mean])
        s3 = (C.c_float * 3)(*[float(x) for x in std])
        call("vitta_gather_crop_resize_normalize_u8", ptr(frames_u8), f, h, w, ptr(idx), n, bx, v, ptr(tabs[0]),
             ptr(tabs[1]), ptr(tabs[2]), ptr(tabs[3]), slots, int(oh), int(ow), m3, s3, layout, int(clip_len), ptr(out),
             stream_ptr())
        return out
    y0, x0, oh, ow = crop if crop is not None else (0, 0, h, w)
    out = torch.empty((n * 3, oh, ow) if layout == 0 else (v, 3, clip_len, oh, ow), dtype=torch.float32,
                      device=frames_u8.device)
    m3 = (C.c_float * 3)(*[float(x) for x in mean])
    s3 = (C.c_float * 3)(*[float(x) for x in std])
    call("vitta_gather_normalize_u8", ptr(frames_u8), f, h, w, ptr(idx), n, int(y0), int(x0), int(oh), int(ow), m3, s3,
         layout, int(clip_len), ptr(out), stream_ptr())
    return out


def cv_linear_tables(src, dst, horizontal):
    """OpenCV INTER_LINEAR taps of one axis (host arithmetic in the library): (ofs (dst,) int32, w (dst, 2) int32)."""
    ofs = np.zeros(dst, np.int32)
    w = np.zeros((dst, 2), np.int32)
    _lib.check(_lib.load().vitta_cv_linear_tables(int(src), int(dst), 1 if horizontal else 0,
                                                  ofs.ctypes.data_as(C.POINTER(C.c_int32)),
                                                  w.ctypes.data_as(C.POINTER(C.c_int32))), "vitta_cv_linear_tables")
    return ofs, w


SWIN_MEAN_255 = (123.675, 116.28, 103.53)       # utils/opts.py:8-9 (img_norm_cfg, 0..255 scale)
SWIN_STD_255 = (58.395, 57.12, 57.375)


def swin_views_to_device(frames_u8, indices, clip_len, scale_size, input_size, bbox=None, mean=SWIN_MEAN_255,
                         std=SWIN_STD_255):
    """The Video-Swin loader's item from decoded frames on the device (video_dataset.py:66-101): frames[indices] ->
    ``Resize(scale=(-1, scale_size))`` -> evaluation: ``CenterCrop(input_size)``; TTA views: crop ``bbox`` (left, top,
    right, bottom in the resized frame; ``swin_random_resized_crop_bbox``) + ``Resize((S, S), keep_ratio=False)`` ->
    ``Normalize`` -> ``FormatShape('NCTHW')``: (V, 3, T, S, S) fp32.  Resizes are OpenCV's INTER_LINEAR, bit exact."""
    if not frames_u8.is_cuda or frames_u8.dtype != torch.uint8 or frames_u8.dim() != 4 or frames_u8.shape[-1] != 3:
        raise _lib.VittaError("swin_views_to_device: frames must be a (F, H, W, 3) uint8 CUDA tensor; there is no CPU path")
    frames_u8 = frames_u8.contiguous()
    dev = frames_u8.device
    f, h, w, _ = frames_u8.shape
    idx = torch.as_tensor(np.asarray(indices, dtype=np.int32)).to(dev)
    n = idx.numel()
    if n % clip_len:
        raise _lib.VittaError("swin_views_to_device: %d indices are not a multiple of clip_len %d" % (n, clip_len))
    v, s = n // clip_len, int(input_size)
    nw, nh = swin_rescale_size(w, h, scale_size)
    up = lambda a: torch.from_numpy(a).to(dev)
    xo, xw = cv_linear_tables(w, nw, True)
    yo, yw = cv_linear_tables(h, nh, False)
    big = torch.empty((n, nh, nw, 3), dtype=torch.uint8, device=dev)
    t1 = [up(a) for a in (xo, xw, yo, yw)]
    call("vitta_cv_resize_u8", ptr(frames_u8), f, h, w, ptr(idx), n, 0, 0, w, h, ptr(t1[0]), ptr(t1[1]), ptr(t1[2]),
         ptr(t1[3]), nh, nw, ptr(big), stream_ptr())
    out = torch.empty((v, 3, clip_len, s, s), dtype=torch.float32, device=dev)
    m3 = (C.c_float * 3)(*[float(x) for x in mean])
    s3 = (C.c_float * 3)(*[float(x) for x in std])
    if bbox is None:
        left, top, right, bottom = swin_center_crop_box(nw, nh, s)
        if left < 0 or top < 0:
            raise _lib.VittaError("swin_views_to_device: %dx%d is smaller than the %d crop" % (nw, nh, s))
    else:
        left, top, right, bottom = (int(x) for x in bbox)
    cw, ch = right - left, bottom - top
    xo, xw = cv_linear_tables(cw, s, True)         # identity taps when the crop already has the target size
    yo, yw = cv_linear_tables(ch, s, False)
    t2 = [up(a) for a in (xo, xw, yo, yw)]
    call("vitta_cv_resize_normalize_u8", ptr(big), n, nh, nw, None, n, left, top, cw, ch, ptr(t2[0]), ptr(t2[1]),
         ptr(t2[2]), ptr(t2[3]), s, s, m3, s3, 1, int(clip_len), ptr(out), stream_ptr())
    return out


# ----------------------------------------------------------------------------------------------
# the reference's TANet dataset item, from decoded frames that are already on the device
# ----------------------------------------------------------------------------------------------
class DecodedVideoDataset(torch.utils.data.Dataset):
    """``Video_TANetDataSet.__getitem__`` + its transform (models/tanet_models/video_dataset.py:306-345,
    corpus/basics.py:1221-1290) for videos whose decoded uint8 frames ``(F, H, W, 3)`` already sit in HBM: frame-index
    sampling, per-view random multi-scale crop (``--if_spatial_rand_cropping``, TTA views only) or scale + centre crop,
    resize, /255 and mean/std, as ONE kernel per item.  Decoding itself (decord) stays outside.

    Use it through ``args.dataset_factory = lambda args, split, kind: DecodedVideoDataset(videos, labels, args, kind)``
    with ``num_workers=0`` (the tensors live on the GPU).  TANet layouts and arithmetic only: the Video-Swin loader of the
    reference resizes with mmcv / OpenCV, whose arithmetic differs -- see ``DecodedSwinVideoDataset``."""

    on_device = True      # corpus.basics._loader: iterate in-process, do not pin

    def __init__(self, videos, labels, args, dataset_type='tta', rng=_random):
        if args.arch != 'tanet':
            raise NotImplementedError("DecodedVideoDataset mirrors the TANet loader; the Swin loader's OpenCV resize is "
                                      "not reproduced")
        if len(videos) != len(labels):
            raise _lib.VittaError("DecodedVideoDataset: %d videos, %d labels" % (len(videos), len(labels)))
        if getattr(args, 'test_crops', 1) not in (1, 3):
            raise NotImplementedError(f'{args.test_crops} spatial crops not implemented!')      # basics.py:1264-1267
        self.three_crops = getattr(args, 'test_crops', 1) == 3
        self.videos, self.labels, self.args, self.rng = videos, labels, args, rng
        self.sample_views = bool(args.if_sample_tta_aug_views) if dataset_type == 'tta' else False   # basics.py:1232-1238
        self.rand_crop = (bool(getattr(args, 'if_spatial_rand_cropping', True)) if self.sample_views else False) \
            and not self.three_crops
        self.input_size = args.scale_size if getattr(args, 'full_res', False) else args.input_size  # basics.py:1230

    def __len__(self):
        return len(self.videos)

    def plan(self, index):
        """Host side of one item: (frame indices (V*T,), crop boxes or None).  Draw order as in the reference: the index
        rule is deterministic, then one crop box per temporal view, view after view."""
        a = self.args
        f, h, w, _ = self.videos[index].shape
        t = a.clip_length
        if self.sample_views:
            idx = np.concatenate([sample_tta_view_indices(f, t, a.n_augmented_views, style)
                                  for style in a.tta_view_sample_style_list])
        else:
            idx = test_clip_indices(f, t, getattr(a, 'sample_style', 'uniform-1'))
        boxes = None
        if self.rand_crop:
            if idx.size != a.n_augmented_views * t:        # the reference's transform asserts this (transforms.py:303)
                raise _lib.VittaError("random cropping needs n_augmented_views * clip_length frames, got %d" % idx.size)
            boxes = sample_view_crops(w, h, self.input_size, a.n_augmented_views, self.rng)
        return idx, boxes

    def __getitem__(self, index):
        idx, boxes = self.plan(index)
        a = self.args
        mean = getattr(a, 'input_mean', synth.INPUT_MEAN)
        std = getattr(a, 'input_std', synth.INPUT_STD)
        if boxes is not None:
            x = views_to_device(self.videos[index], idx, a.clip_length, 'tanet', mean=mean, std=std, boxes=boxes,
                                out_size=self.input_size)
        else:
            x = views_to_device(self.videos[index], idx, a.clip_length, 'tanet', mean=mean, std=std,
                                scale_size=a.scale_size, out_size=self.input_size, three_crops=self.three_crops)
        return x, self.labels[index]


class DecodedSwinVideoDataset(torch.utils.data.Dataset):
    """``Video_SwinDataset.__getitem__`` (models/videoswintransformer_models/video_dataset.py:58-110) for decoded uint8
    frames resident in HBM: SampleFrames -> Resize(-1, scale_size) -> [TTA views: RandomResizedCrop (ONE box per video) +
    Resize(S, S)] or [CenterCrop(S)] -> Flip (``--flip_ratio 0``: never flips, but draws) -> Normalize -> NCTHW.
    The random draws are made in the pipeline's order from ``np_rng`` / ``py_rng`` (numpy's legacy generator and
    ``random``, like the reference).  Use through ``args.dataset_factory`` with ``num_workers=0``."""

    on_device = True      # corpus.basics._loader: iterate in-process, do not pin

    def __init__(self, videos, labels, args, dataset_type='tta', np_rng=np.random, py_rng=_random):
        if args.arch != 'videoswintransformer':
            raise NotImplementedError("DecodedSwinVideoDataset mirrors the Video-Swin loader; use DecodedVideoDataset for TANet")
        if len(videos) != len(labels):
            raise _lib.VittaError("DecodedSwinVideoDataset: %d videos, %d labels" % (len(videos), len(labels)))
        if getattr(args, 'flip_ratio', 0) != 0:
            raise NotImplementedError("flip_ratio != 0: horizontal flips are not on the device")
        self.videos, self.labels, self.args, self.np_rng, self.py_rng = videos, labels, args, np_rng, py_rng
        self.sample_views = bool(args.if_sample_tta_aug_views) if dataset_type == 'tta' else False   # basics.py:1197-1201

    def __len__(self):
        return len(self.videos)

    def plan(self, index):
        """Host side of one item: (frame indices, crop box in the resized frame or None)."""
        a = self.args
        f, h, w, _ = self.videos[index].shape
        t = a.clip_length
        bbox = None
        if self.sample_views:
            idx = np.concatenate([sample_tta_view_indices(f, t, a.n_augmented_views, style, np_rng=self.np_rng)
                                  for style in a.tta_view_sample_style_list])
            nw, nh = swin_rescale_size(w, h, a.scale_size)
            bbox = swin_random_resized_crop_bbox(nh, nw, np_rng=self.np_rng, py_rng=self.py_rng)
        else:
            if not getattr(a, 'frame_uniform', True) or getattr(a, 'num_clips', 1) != 1:
                raise NotImplementedError("dense clip sampling (frame_uniform=False / num_clips > 1) is not mirrored")
            idx = swin_seq_frames(f, t)
        self.np_rng.rand()             # Flip.__call__ draws once per item even with flip_ratio 0 (transforms_backup.py:1073)
        return idx, bbox

    def __getitem__(self, index):
        idx, bbox = self.plan(index)
        a = self.args
        cfg = getattr(a, 'img_norm_cfg', None) or {}
        x = swin_views_to_device(self.videos[index], idx, a.clip_length, a.scale_size, a.input_size, bbox,
                                 mean=cfg.get('mean', SWIN_MEAN_255), std=cfg.get('std', SWIN_STD_255))
        return x, self.labels[index]
