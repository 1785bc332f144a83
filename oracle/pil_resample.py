"""TEST INFRASTRUCTURE -- CPU restatement of the spatial part of the reference's view pipeline.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product never does.

The reference crops every temporal view with ``SubgroupWise_MultiScaleCrop_TANet`` (models/tanet_models/transforms.py:277-384;
selected by corpus/basics.py:1238-1245 when ``--if_spatial_rand_cropping`` is on, its default, utils/opts.py:85) and
resizes the crop to the network input with ``PIL.Image.resize(size, Image.BILINEAR)`` (transforms.py:319-322).  The
resize arithmetic lives in a third-party dependency that is not under /root/reference: **Pillow, pinned 8.4.0**
(requirements.txt:37), ``src/libImaging/Resample.c``.  Its published algorithm for 8-bit images is restated here:

  * ``precompute_coeffs``: per output position, ``scale = in/out``, ``filterscale = max(scale, 1)``, ``support = filterscale``
    (bilinear support 1.0), ``ksize = ceil(support)*2 + 1``, window ``[int(center-support+0.5), int(center+support+0.5))``
    clamped to the image, triangle weights ``1 - |(x - center + 0.5)/filterscale|`` normalised to sum 1 (double precision);
  * ``normalize_coeffs_8bpc``: fixed point with PRECISION_BITS = 32 - 8 - 2 = 22, ``(int)(0.5 + k * 2^22)``;
  * ``ImagingResampleHorizontal_8bpc`` then ``ImagingResampleVertical_8bpc``: ``clip8((2^21 + sum pixel*k) >> 22)`` per band,
    the horizontal pass result rounded to uint8 BEFORE the vertical pass; a pass whose size does not change is skipped
    (its coefficients would be the identity, so running it is equivalent).

Parity pin: tests/test_crops.py checks this restatement bit for bit against the Pillow installed in the build container
(12.2.0; the 8bpc bilinear path is unchanged since 5.x) and against outputs of the reference's own transform recorded in
tests/golden/crops.npz (oracle/make_golden.py::run_crops_case).  Integer work: the bar is bit-exact.
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def resample_ksize(in_size, out_size):
    """Number of coefficient slots per output position (Resample.c precompute_coeffs: ksize)."""
    filterscale = max(float(in_size) / out_size, 1.0)
    return int(math.ceil(filterscale)) * 2 + 1


def resample_coeffs(in_size, out_size):
    """(bounds[out, 2] int32 = (first input index, count), kk[out, ksize] int32 fixed-point weights)."""
    scale = float(in_size) / out_size
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = 0.0 + (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = []
        for x in range(xmax):
            a = abs((x + xmin - center + 0.5) * ss)
            w.append(1.0 - a if a < 1.0 else 0.0)
        ww = sum(w)           # same left-to-right double accumulation as the C loop
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + k * (1 << PRECISION_BITS)) if k < 0 else int(0.5 + k * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return bounds, kk


def _pass(img, bounds, kk, axis):
    """One 8bpc pass along ``axis`` (0 = vertical, 1 = horizontal) of a (H, W, 3) uint8 image."""
    img = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + img.shape[1:], np.int64)
    for o in range(bounds.shape[0]):
        lo, n = int(bounds[o, 0]), int(bounds[o, 1])
        acc = np.full(img.shape[1:], 1 << (PRECISION_BITS - 1), np.int64)
        for x in range(n):
            acc += img[lo + x] * int(kk[o, x])
        out[o] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return np.moveaxis(out.astype(np.uint8), 0, axis)


def resize_bilinear_u8(img, out_w, out_h):
    """``Image.fromarray(img).resize((out_w, out_h), Image.BILINEAR)`` for a (H, W, 3) uint8 array."""
    h, w = img.shape[:2]
    if w != out_w:
        img = _pass(img, *resample_coeffs(w, out_w), axis=1)
    if h != out_h:
        img = _pass(img, *resample_coeffs(h, out_h), axis=0)
    return img


def fill_fix_offset(image_w, image_h, crop_w, crop_h):
    """The 13 fixed crop positions (transforms.py:361-385, more_fix_crop=True)."""
    ws, hs = (image_w - crop_w) // 4, (image_h - crop_h) // 4
    return [(0, 0), (4 * ws, 0), (0, 4 * hs), (4 * ws, 4 * hs), (2 * ws, 2 * hs),
            (0, 2 * hs), (4 * ws, 2 * hs), (2 * ws, 4 * hs), (2 * ws, 0),
            (ws, hs), (3 * ws, hs), (ws, 3 * hs), (3 * ws, 3 * hs)]


def crop_candidates(image_w, image_h, input_w, input_h, scales=(1, .875, .75, .66), max_distort=1):
    """(crop_w, crop_h) pairs in the reference's enumeration order (transforms.py:325-345)."""
    base = min(image_w, image_h)
    sizes = [int(base * s) for s in scales]
    ch = [input_h if abs(x - input_h) < 3 else x for x in sizes]
    cw = [input_w if abs(x - input_w) < 3 else x for x in sizes]
    return [(w, h) for i, h in enumerate(ch) for j, w in enumerate(cw) if abs(i - j) <= max_distort]


def sample_crop(image_w, image_h, input_size, rng):
    """One draw of ``_sample_crop_size`` (fix_crop=True): ``rng`` is a ``random.Random``-like object; the two
    ``choice`` calls are made in the reference's order.  Returns (crop_w, crop_h, offset_w, offset_h)."""
    cw, ch = rng.choice(crop_candidates(image_w, image_h, input_size, input_size))
    ow, oh = rng.choice(fill_fix_offset(image_w, image_h, cw, ch))
    return cw, ch, ow, oh


def crop_resize_views(frames, indices, clip_len, boxes, out_size):
    """frames (F, H, W, 3) uint8, indices (V*T,), boxes [(crop_w, crop_h, off_w, off_h)] per view -> (V*T, S, S, 3) uint8."""
    out = []
    for k, f in enumerate(indices):
        cw, ch, ow, oh = boxes[k // clip_len]
        out.append(resize_bilinear_u8(frames[int(f)][oh:oh + ch, ow:ow + cw], out_size, out_size))
    return np.stack(out)


def scale_center_crop_u8(img, scale_size, input_size):
    """GroupScale_TANet(scale_size) + GroupCenterCrop_TANet(input_size) on one (H, W, 3) uint8 frame (transforms.py:46-52,
    170-183): torchvision 0.8.2 (requirements.txt:63) ``Resize(int)`` -- smaller edge to ``scale_size``,
    ``int(size * long / short)`` for the other, untouched when it already matches -- then ``CenterCrop``:
    ``int(round((extent - S) / 2.))``."""
    h, w = img.shape[:2]
    if not ((w <= h and w == scale_size) or (h <= w and h == scale_size)):
        if w < h:
            img = resize_bilinear_u8(img, scale_size, int(scale_size * h / w))
        else:
            img = resize_bilinear_u8(img, int(scale_size * w / h), scale_size)
    h, w = img.shape[:2]
    top, left = int(round((h - input_size) / 2.)), int(round((w - input_size) / 2.))
    return img[top:top + input_size, left:left + input_size]


def full_res_sample_u8(img, scale_size, input_size):
    """GroupFullResSample_TANet(input_size, scale_size, flip=False) on one frame (transforms.py:227-272): the three crops
    (left, right, centre) of the frame scaled to ``scale_size`` -> (3, S, S, 3) uint8."""
    h, w = img.shape[:2]
    if not ((w <= h and w == scale_size) or (h <= w and h == scale_size)):
        if w < h:
            img = resize_bilinear_u8(img, scale_size, int(scale_size * h / w))
        else:
            img = resize_bilinear_u8(img, int(scale_size * w / h), scale_size)
    h, w = img.shape[:2]
    ws, hs = (w - input_size) // 4, (h - input_size) // 4
    return np.stack([img[oh:oh + input_size, ow:ow + input_size]
                     for ow, oh in ((0, 2 * hs), (4 * ws, 2 * hs), (2 * ws, 2 * hs))])
