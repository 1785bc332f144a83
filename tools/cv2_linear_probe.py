"""STUDY TOOL (not on any product path): restatement of OpenCV's 8-bit INTER_LINEAR resize, checked against the installed cv2.

The Video-Swin loader of the reference resizes with mmcv.imresize = cv2.resize(INTER_LINEAR)
(models/videoswintransformer_models/transforms_backup.py:749-876, video_dataset.py:66-101) instead of PIL, so a GPU
version of that loader needs this arithmetic, not Pillow's (oracle/pil_resample.py).  Findings, bit-exact against cv2 4.13
on the cases below (up- and down-scaling, exact 2x, degenerate sizes):
  * scale = 1 / (dst / src) in double; per output position f = float32((d + 0.5) * scale - 0.5), s = floor(f), f -= s;
  * HORIZONTAL taps: s < 0 -> (s, f) = (0, 0); s >= src - 1 -> (src - 1, 0); weights cvRound((1 - f) * 2048), cvRound(f * 2048)
    (float32 products, round half to even); the pass keeps 32-bit sums (no rounding to uint8 in between);
  * VERTICAL taps: NO such border rule -- the row indices s, s + 1 are clipped to [0, src - 1] and the weights stay as
    computed, so the first / last rows mix the same row with two separately truncated products;
  * output = ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
This probe came first; the restatement now lives in oracle/cv2_resample.py and the device version in preprocess.cu (K14)."""
import numpy as np, cv2, math
def cv_round(x): return int(np.rint(x))
def tabs(ssize, dsize, border_fix):
    scale = 1.0 / (float(dsize) / ssize)
    ofs = np.zeros(dsize, np.int64); al = np.zeros((dsize, 2), np.int64)
    for d in range(dsize):
        f = np.float32((d + 0.5) * scale - 0.5)
        s = int(math.floor(f))
        f = np.float32(f - np.float32(s))
        if border_fix:
            if s < 0: f = np.float32(0); s = 0
            if s >= ssize - 1: f = np.float32(0); s = ssize - 1
        a0 = np.float32(np.float32(1.0) - f) * np.float32(2048); a1 = f * np.float32(2048)
        ofs[d] = s; al[d] = (cv_round(a0), cv_round(a1))
    return ofs, al
def resize(img, dw, dh):
    h, w = img.shape[:2]
    xo, xa = tabs(w, dw, True); yo, ya = tabs(h, dh, False)
    src = img.astype(np.int64)
    x1 = np.minimum(xo + 1, w - 1)
    H = src[:, xo] * xa[:, 0][None, :, None] + src[:, x1] * xa[:, 1][None, :, None]
    y0 = np.clip(yo, 0, h - 1); y1 = np.clip(yo + 1, 0, h - 1)
    S0, S1 = H[y0], H[y1]
    b0, b1 = ya[:, 0][:, None, None], ya[:, 1][:, None, None]
    out = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)
rng = np.random.Generator(np.random.PCG64(0))
bad = 0
for (h, w, dh, dw) in [(240, 320, 256, 341), (256, 341, 224, 224), (100, 100, 224, 224), (37, 53, 16, 16), (180, 210, 224, 224), (240, 320, 120, 160), (64, 64, 32, 32), (9, 7, 23, 31), (480, 640, 256, 341), (224,224,224,224), (128,171,256,342), (5,5,1,1), (1,1,7,9), (2,3,100,50), (300, 200, 77, 51), (96,128,80,107)]:
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
    got = resize(img, dw, dh)
    d = np.abs(got.astype(int) - want.astype(int))
    if d.max(): bad += 1; print(h, w, dh, dw, "max diff", d.max(), "frac", (d > 0).mean())
print("bad", bad, cv2.__version__, cv2.ipp.useIPP() if hasattr(cv2, "ipp") else None)
