"""Video-Swin loader, spatial side (reference: models/videoswintransformer_models/video_dataset.py:66-101,
transforms_backup.py; arithmetic: OpenCV's 8-bit INTER_LINEAR through mmcv.imresize).  CPU: the oracle's restatement against
the installed OpenCV, bit for bit."""
import numpy as np
import pytest

import cases  # noqa: F401  (puts the repo root on sys.path)


def test_oracle_linear_resize_matches_installed_opencv():
    cv2 = pytest.importorskip("cv2")
    from oracle import cv2_resample as R
    rng = np.random.Generator(np.random.PCG64(0))
    for h, w, dh, dw in [(240, 320, 256, 341), (256, 341, 224, 224), (100, 100, 224, 224), (37, 53, 16, 16),
                         (180, 210, 224, 224), (240, 320, 120, 160), (64, 64, 32, 32), (9, 7, 23, 31), (480, 640, 256, 341),
                         (224, 224, 224, 224), (128, 171, 256, 342), (5, 5, 1, 1), (1, 1, 7, 9), (2, 3, 100, 50),
                         (300, 200, 77, 51), (96, 128, 80, 107)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR)
        assert (R.resize_linear_u8(img, dw, dh) == want).all(), (h, w, dh, dw)


def test_oracle_item_pipeline_matches_opencv_composition():
    """Resize(-1, Z) -> CenterCrop(S), and Resize -> crop(bbox) -> Resize((S, S)), composed from cv2 calls."""
    cv2 = pytest.importorskip("cv2")
    from oracle import cv2_resample as R
    rng = np.random.Generator(np.random.PCG64(1))
    frames = rng.integers(0, 256, (6, 48, 64, 3), dtype=np.uint8)
    idx = [0, 5, 2]
    z, s = 40, 32
    nw, nh = R.rescale_size(64, 48, z)
    assert (nw, nh) == (53, 40)
    big = [cv2.resize(frames[i], (nw, nh), interpolation=cv2.INTER_LINEAR) for i in idx]
    left, top = (nw - s) // 2, (nh - s) // 2
    want = np.stack([b[top:top + s, left:left + s] for b in big])
    assert (R.swin_item_u8(frames, idx, z, s) == want).all()
    bbox = (7, 3, 41, 30)
    want = np.stack([cv2.resize(b[3:30, 7:41], (s, s), interpolation=cv2.INTER_LINEAR) for b in big])
    assert (R.swin_item_u8(frames, idx, z, s, bbox) == want).all()
