"""TEST INFRASTRUCTURE -- CPU restatement of the Video-Swin loader's spatial arithmetic (Swin side).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file; the product never does.

The reference's Video-Swin loader (models/videoswintransformer_models/video_dataset.py:66-101) resizes with
``mmcv.imresize(img, (w, h), interpolation='bilinear')`` (transforms_backup.py:794-798), i.e. ``cv2.resize(...,
INTER_LINEAR)`` on uint8 frames.  Both libraries are third-party and absent from the reference tree: mmcv-full 1.3.12
(requirements.txt:25, not installed here -- its size rule ``rescale_size`` is restated from its published source, parity
unpinned) and OpenCV (opencv-python 4.5.x era in requirements.txt; 4.13.0 is installed here and is what this restatement
is pinned against, bit for bit, in tests/test_swin_loader.py).  OpenCV's 8-bit INTER_LINEAR path (imgproc/resize.cpp:
``resizeGeneric_`` with ``HResizeLinear<uchar, int, short, 2048>`` / ``VResizeLinear<uchar, int, short, FixedPtCast<int,
uchar, 22>>``):

  * ``scale = 1 / (dst / src)`` in double; per output position ``f = float((d + 0.5) * scale - 0.5)``, ``s = floor(f)``,
    ``f -= s`` (single precision);
  * horizontal taps only: ``s < 0 -> (s, f) = (0, 0)``; ``s >= src - 1 -> (src - 1, 0)``; weights
    ``saturate_cast<short>((1 - f) * 2048)``, ``(f * 2048)`` = round-half-even of the float products;
  * the horizontal pass keeps 32-bit sums ``S = p[s] * a0 + p[s + 1] * a1`` (no rounding to uint8 in between);
  * vertical taps: no border rule for the weights, the two row indices ``s, s + 1`` are clipped to ``[0, src - 1]``;
  * output ``= (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2``.
The loader's own geometry (transforms_backup.py): ``Resize(scale=(-1, Z))`` -> short edge Z (mmcv rule), evaluation:
``CenterCrop(S)``; TTA views: ``RandomResizedCrop`` (ONE box for all frames of all views) + ``Resize((S, S),
keep_ratio=False)``; ``Normalize``: ``(x - mean) * (1 / std)`` on the 0..255 scale; ``FormatShape('NCTHW')``.
"""
import math

import numpy as np

COEF_BITS = 11
COEF_ONE = 1 << COEF_BITS


def _round_half_even(x):
    return int(np.rint(x))


def linear_tables(src, dst, horizontal):
    """(ofs[dst] int64 first tap (unclipped for the vertical pass), w[dst, 2] int64 11-bit weights)."""
    scale = 1.0 / (float(dst) / src)
    ofs = np.zeros(dst, np.int64)
    w = np.zeros((dst, 2), np.int64)
    for d in range(dst):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if horizontal:
            if s < 0:
                f, s = np.float32(0), 0
            if s >= src - 1:
                f, s = np.float32(0), src - 1
        a0 = np.float32(np.float32(1.0) - f) * np.float32(COEF_ONE)
        a1 = f * np.float32(COEF_ONE)
        ofs[d] = s
        w[d] = (_round_half_even(a0), _round_half_even(a1))
    return ofs, w


def resize_linear_u8(img, out_w, out_h):
    """``cv2.resize(img, (out_w, out_h), interpolation=cv2.INTER_LINEAR)`` for a (H, W, 3) uint8 array."""
    h, w = img.shape[:2]
    if (w, h) == (out_w, out_h):
        return img.copy()
    xo, xa = linear_tables(w, out_w, True)
    yo, yb = linear_tables(h, out_h, False)
    src = img.astype(np.int64)
    x1 = np.minimum(xo + 1, w - 1)
    hs = src[:, xo] * xa[:, 0][None, :, None] + src[:, x1] * xa[:, 1][None, :, None]
    y0, y1 = np.clip(yo, 0, h - 1), np.clip(yo + 1, 0, h - 1)
    b0, b1 = yb[:, 0][:, None, None], yb[:, 1][:, None, None]
    out = (((b0 * (hs[y0] >> 4)) >> 16) + ((b1 * (hs[y1] >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def rescale_size(w, h, short_edge):
    """mmcv.rescale_size((w, h), (inf, short_edge)) -- restated, unpinned (mmcv absent)."""
    factor = float(short_edge) / min(h, w)
    return int(w * factor + 0.5), int(h * factor + 0.5)


def swin_item_u8(frames, indices, scale_size, input_size, bbox=None):
    """uint8 part of one loader item: frames[indices] -> Resize(-1, Z) -> CenterCrop(S) (bbox None) or crop(bbox) ->
    Resize((S, S)).  bbox = (left, top, right, bottom) in the resized frame.  Returns (len(indices), S, S, 3) uint8."""
    out = []
    for i in indices:
        f = frames[int(i)]
        h, w = f.shape[:2]
        nw, nh = rescale_size(w, h, scale_size)
        f = resize_linear_u8(f, nw, nh)
        if bbox is None:
            left, top = (nw - input_size) // 2, (nh - input_size) // 2
            f = f[top:top + input_size, left:left + input_size]
        else:
            left, top, right, bottom = bbox
            f = resize_linear_u8(f[top:bottom, left:right], input_size, input_size)
        out.append(f)
    return np.stack(out)
