"""Video-Swin loader, spatial side (reference: models/videoswintransformer_models/video_dataset.py:66-101,
transforms_backup.py; arithmetic: OpenCV's 8-bit INTER_LINEAR through mmcv.imresize).  CPU: the oracle's restatement against
the installed OpenCV, bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

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


def test_normalisation_formula_matches_opencv_for_every_pixel_value():
    """The kernel's normalisation -- float32 (x - mean), then the product with the float64 reciprocal of the float32 std
    rounded once -- against mmcv.imnormalize_'s cv2.subtract / cv2.multiply, for all 256 values of every channel."""
    cv2 = pytest.importorskip("cv2")
    from vitta_b200.corpus.views import SWIN_MEAN_255, SWIN_STD_255
    mean, std = np.array(SWIN_MEAN_255, np.float32), np.array(SWIN_STD_255, np.float32)        # Normalize.__init__ (:1146-1147)
    img = np.stack([np.arange(256)] * 3, -1).astype(np.float32).reshape(16, 16, 3)
    ref = img.copy()
    cv2.subtract(ref, np.float64(mean.reshape(1, -1)), ref)
    cv2.multiply(ref, 1 / np.float64(std.reshape(1, -1)), ref)
    mine = ((img - mean).astype(np.float64) * (1.0 / std.astype(np.float64))).astype(np.float32)
    assert np.array_equal(mine, ref)


def test_library_tap_tables_match_oracle():
    from oracle import cv2_resample as R
    from vitta_b200.corpus.views import cv_linear_tables
    for src, dst in [(320, 341), (240, 256), (341, 224), (100, 224), (224, 224), (53, 16), (7, 31), (1, 9), (5, 1), (640, 341),
                     (34, 32), (27, 32), (455, 224)]:
        for horizontal in (True, False):
            ofs, w = cv_linear_tables(src, dst, horizontal)
            o2, w2 = R.linear_tables(src, dst, horizontal)
            assert (ofs == o2).all() and (w == w2).all(), (src, dst, horizontal)
            assert (w.sum(1) == 2048).all()


def _emulate(src, idx, region, tabs, out_h, out_w):
    """numpy transcription of cv_resize_kernel (preprocess.cu): uint8 result (n, out_h, out_w, 3)."""
    x0, y0, cw, ch = region
    xofs, xw, yofs, yw = tabs
    n = len(idx)
    out = np.zeros((n, out_h, out_w, 3), np.int64)
    for k in range(n):
        fr = src[min(max(int(idx[k]), 0), len(src) - 1)].astype(np.int64)
        for y in range(out_h):
            r0 = y0 + min(max(int(yofs[y]), 0), ch - 1)
            r1 = y0 + min(max(int(yofs[y]) + 1, 0), ch - 1)
            b0, b1 = int(yw[y, 0]), int(yw[y, 1])
            for x in range(out_w):
                c0 = x0 + min(max(int(xofs[x]), 0), cw - 1)
                c1 = x0 + min(max(int(xofs[x]) + 1, 0), cw - 1)
                a0, a1 = int(xw[x, 0]), int(xw[x, 1])
                s0 = fr[r0, c0] * a0 + fr[r0, c1] * a1
                s1 = fr[r1, c0] * a0 + fr[r1, c1] * a1
                out[k, y, x] = np.clip((((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2, 0, 255)
    return out.astype(np.uint8)


def test_kernel_index_arithmetic_emulated_against_oracle():
    """The kernel's tap / clipping / region arithmetic, transcribed to numpy and fed with the library's tables, reproduces
    the oracle: whole-frame resize (with the frame-index clamp), then crop + resize of a region of the result."""
    from oracle import cv2_resample as R
    from vitta_b200.corpus.views import cv_linear_tables, swin_rescale_size
    rng = np.random.Generator(np.random.PCG64(4))
    frames = rng.integers(0, 256, (4, 18, 24, 3), dtype=np.uint8)
    idx = [0, 3, 7]
    nw, nh = swin_rescale_size(24, 18, 15)
    t1 = cv_linear_tables(24, nw, True) + cv_linear_tables(18, nh, False)
    big = _emulate(frames, idx, (0, 0, 24, 18), t1, nh, nw)
    want_big = np.stack([R.resize_linear_u8(frames[min(i, 3)], nw, nh) for i in idx])
    assert (big == want_big).all()
    bbox = (3, 2, 17, 13)
    cw, ch = 14, 11
    t2 = cv_linear_tables(cw, 8, True) + cv_linear_tables(ch, 8, False)
    got = _emulate(big, [0, 1, 2], (3, 2, cw, ch), t2, 8, 8)
    assert (got == R.swin_item_u8(frames, [0, 3, 3], 15, 8, bbox)).all()
    t3 = cv_linear_tables(8, 8, True) + cv_linear_tables(8, 8, False)         # centre crop = identity taps on the region
    got = _emulate(big, [0, 1, 2], (6, 3, 8, 8), t3, 8, 8)
    assert (got == R.swin_item_u8(frames, [0, 3, 3], 15, 8)).all()


def test_entry_points_validate_before_touching_the_device():
    if torch.cuda.is_available():
        pytest.skip("marshalling-only test: meant for the GPU-less container")
    from vitta_b200 import _lib
    from vitta_b200.corpus.views import cv_linear_tables, swin_views_to_device
    src = torch.zeros(2, 18, 24, 3, dtype=torch.uint8)
    out = torch.zeros(2, 8, 8, 3, dtype=torch.uint8)
    tabs = [torch.from_numpy(a) for a in cv_linear_tables(10, 8, True) + cv_linear_tables(9, 8, False)]

    def go(x0):
        _lib.call("vitta_cv_resize_u8", _lib.ptr(src), 2, 18, 24, None, 2, x0, 2, 10, 9, _lib.ptr(tabs[0]), _lib.ptr(tabs[1]),
                  _lib.ptr(tabs[2]), _lib.ptr(tabs[3]), 8, 8, _lib.ptr(out), C.c_void_p(0))
    with pytest.raises(_lib.VittaError) as e:
        go(3)
    assert "outside the frame" not in str(e.value)
    with pytest.raises(_lib.VittaError, match="outside the frame"):
        go(15)
    with pytest.raises(_lib.VittaError):
        swin_views_to_device(src, [0, 1], 2, 15, 8)          # host tensor: no CPU path


def test_dataset_plan_draws_in_the_pipeline_order():
    """SampleFrames (deterministic style) -> RandomResizedCrop (numpy uniform x 2, random.randint x 2) -> Flip (one numpy
    draw per item): the plan of consecutive items equals the pinned pieces called in that order on the same generators."""
    import random
    from vitta_b200.corpus.views import (DecodedSwinVideoDataset, sample_tta_view_indices, swin_random_resized_crop_bbox,
                                         swin_rescale_size, swin_seq_frames)
    from vitta_b200.utils.opts import default_args
    args = default_args(arch="videoswintransformer", clip_length=4, input_size=32, scale_size=40, n_augmented_views=2,
                        if_sample_tta_aug_views=True)
    vids = [torch.zeros(11, 48, 64, 3, dtype=torch.uint8), torch.zeros(20, 60, 44, 3, dtype=torch.uint8)]
    ds = DecodedSwinVideoDataset(vids, [1, 2], args, "tta", np_rng=np.random.RandomState(3), py_rng=random.Random(3))
    nrs, prs = np.random.RandomState(3), random.Random(3)
    for i, v in enumerate(vids):
        idx, bbox = ds.plan(i)
        f, h, w, _ = v.shape
        assert (idx == sample_tta_view_indices(f, 4, 2, "uniform_equidist")).all()
        nw, nh = swin_rescale_size(w, h, 40)
        assert bbox == swin_random_resized_crop_bbox(nh, nw, np_rng=nrs, py_rng=prs)
        nrs.rand()
        assert 0 <= bbox[0] < bbox[2] <= nw and 0 <= bbox[1] < bbox[3] <= nh
    ev = DecodedSwinVideoDataset(vids, [1, 2], args, "eval", np_rng=np.random.RandomState(3), py_rng=random.Random(3))
    idx, bbox = ev.plan(1)
    assert bbox is None and (idx == swin_seq_frames(20, 4)).all()
    with pytest.raises(NotImplementedError):
        DecodedSwinVideoDataset(vids, [1, 2], default_args(arch="tanet"), "tta")
    with pytest.raises(NotImplementedError):
        DecodedSwinVideoDataset(vids, [1, 2], default_args(arch="videoswintransformer", flip_ratio=1), "tta")


def test_dataset_plan_matches_reference_pipeline_draws():
    """RandomResizedCrop.__call__ + Flip.__call__ of the UNMODIFIED reference on two consecutive items under seeded
    generators (tests/golden/swin_seq.npz: pipeline/*): same boxes, and both generators end at the same position."""
    import random
    from vitta_b200.corpus.views import DecodedSwinVideoDataset
    from vitta_b200.utils.opts import default_args
    g = np.load(os.path.join(cases.GOLDEN_DIR, "swin_seq.npz"))
    args = default_args(arch="videoswintransformer", clip_length=4, input_size=32, scale_size=40, n_augmented_views=2,
                        if_sample_tta_aug_views=True)
    vids = [torch.zeros(11, 48, 64, 3, dtype=torch.uint8), torch.zeros(20, 60, 44, 3, dtype=torch.uint8)]
    nrs, prs = np.random.RandomState(3), random.Random(3)
    ds = DecodedSwinVideoDataset(vids, [1, 2], args, "tta", np_rng=nrs, py_rng=prs)
    got = [ds.plan(i)[1] for i in range(2)]
    assert (np.asarray(got) == g["pipeline/boxes"]).all(), (got, g["pipeline/boxes"])
    assert [nrs.rand(), prs.random()] == g["pipeline/next"].tolist()


@pytest.mark.gpu
@pytest.mark.parametrize("with_bbox", [False, True])
def test_swin_views_to_device_vs_oracle(cuda_device, with_bbox):
    from oracle import cv2_resample as R
    from vitta_b200.corpus.views import SWIN_MEAN_255, SWIN_STD_255, swin_views_to_device
    rng = np.random.Generator(np.random.PCG64(6))
    frames = rng.integers(0, 256, (9, 96, 128, 3), dtype=np.uint8)
    idx, t, z, s = [0, 2, 4, 6, 1, 3, 5, 12], 4, 80, 64
    bbox = (11, 5, 83, 70) if with_bbox else None
    out = swin_views_to_device(torch.from_numpy(frames).to(cuda_device), idx, t, z, s, bbox)
    u8 = R.swin_item_u8(frames, np.minimum(idx, 8), z, s, bbox)                                 # (V*T, s, s, 3)
    # mmcv.imnormalize_: float32 difference, product with the float64 reciprocal of the float32 std rounded once
    mean32, std32 = torch.tensor(SWIN_MEAN_255, dtype=torch.float32), torch.tensor(SWIN_STD_255, dtype=torch.float32)
    x = ((torch.from_numpy(u8).float() - mean32).double() * (1.0 / std32.double())).float()
    want = x.reshape(2, t, s, s, 3).permute(0, 4, 1, 2, 3).contiguous()
    torch.testing.assert_close(out.cpu(), want, rtol=0, atol=0)             # bit exact, normalisation included
