"""View sampling: index rule against vectors recorded from the reference's own sampler (CPU),
gather + normalise kernel against torch (GPU)."""
import os

import numpy as np
import pytest
import torch

import cases


def test_view_indices_match_reference_sampler():
    from vitta_b200.corpus.views import sample_tta_view_indices
    g = np.load(os.path.join(cases.GOLDEN_DIR, "views.npz"))
    assert len(g.files) == 4 * 7 * 3 * 3
    for key in g.files:
        style, nf, t, views = key.split("/")
        got = sample_tta_view_indices(int(nf), int(t), int(views), style)
        want = g[key]
        assert got.shape == want.shape and (got == want).all(), (key, got, want)   # index work: bit exact


def test_swin_eval_clip_indices_match_reference_sampler():
    """SampleFrames.get_seq_frames (test mode) of the reference's Video-Swin loader, 33 recorded vectors."""
    from vitta_b200.corpus.views import swin_seq_frames
    g = np.load(os.path.join(cases.GOLDEN_DIR, "swin_seq.npz"))
    keys = [k for k in g.files if not k.startswith(("bbox/", "pipeline/"))]
    assert len(keys) == 33
    for key in keys:
        nf, t = (int(v) for v in key.split("/"))
        got = swin_seq_frames(nf, t)
        assert got.shape == g[key].shape and (got == g[key]).all(), (key, got, g[key])


def test_random_view_styles_match_reference_sampler_under_the_same_seed():
    """uniform_rand / dense_rand / random (video_dataset.py:197-229): same numpy legacy-generator calls in the same order,
    so a seeded stream gives the reference's indices -- two consecutive draws per case, 72 cases."""
    from vitta_b200.corpus.views import sample_tta_view_indices
    g = np.load(os.path.join(cases.GOLDEN_DIR, "views_rand.npz"))
    assert len(g.files) == 3 * 8 * 3
    for key in g.files:
        style, nf, t = key.split("/")
        rs = np.random.RandomState(int(nf) * 100 + int(t))
        got = np.stack([sample_tta_view_indices(int(nf), int(t), 1, style, np_rng=rs) for _ in range(2)])
        assert got.shape == g[key].shape and (got == g[key]).all(), (key, got, g[key])
    np.random.seed(7)                                             # the module-level generator is the default, as in the reference
    a = sample_tta_view_indices(100, 16, 1, "uniform_rand")
    b = sample_tta_view_indices(100, 16, 1, "uniform_rand", np_rng=np.random.RandomState(7))
    assert (a == b).all()


def test_swin_random_resized_crop_boxes_match_reference():
    """RandomResizedCrop.get_crop_bbox of the Video-Swin loader: seeded numpy + random generators, 8 consecutive boxes for
    each of 6 frame sizes (landscape, portrait, square, and a very elongated one where most candidates are rejected)."""
    import random
    from vitta_b200.corpus.views import swin_center_crop_box, swin_random_resized_crop_bbox, swin_rescale_size
    g = np.load(os.path.join(cases.GOLDEN_DIR, "swin_seq.npz"))
    keys = [k for k in g.files if k.startswith("bbox/")]
    assert len(keys) == 6
    for key in keys:
        _, ih, iw = key.split("/")
        ih, iw = int(ih), int(iw)
        nrs, prs = np.random.RandomState(ih * 7 + iw), random.Random(ih * 7 + iw)
        got = np.asarray([swin_random_resized_crop_bbox(ih, iw, np_rng=nrs, py_rng=prs) for _ in range(8)])
        assert (got == g[key]).all(), (key, got, g[key])
    assert swin_center_crop_box(341, 256, 224) == (58, 16, 282, 240)
    assert swin_rescale_size(320, 240, 256) == (341, 256) and swin_rescale_size(240, 320, 256) == (256, 341)


def test_evaluation_clip_indices_match_reference():
    """Video_TANetDataSet._get_test_indices for --sample_style uniform-N / dense-N: 120 recorded vectors."""
    from vitta_b200.corpus.views import test_clip_indices
    g = np.load(os.path.join(cases.GOLDEN_DIR, "test_indices.npz"))
    assert len(g.files) == 5 * 8 * 3
    for key in g.files:
        style, nf, t = key.split("/")
        got = test_clip_indices(int(nf), int(t), style)
        assert got.shape == g[key].shape and (got == g[key]).all(), (key, got, g[key])
    with pytest.raises(NotImplementedError):
        test_clip_indices(100, 16, "random-1")


def test_unknown_style_is_loud():
    from vitta_b200.corpus.views import sample_tta_view_indices
    with pytest.raises(NotImplementedError):
        sample_tta_view_indices(100, 16, 2, "uniform_jitter")


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["tanet", "videoswintransformer"])
def test_views_to_device_vs_torch(cuda_device, arch):
    from vitta_b200 import synth
    from vitta_b200.corpus.views import sample_tta_view_indices, views_to_device
    f, h, w, t, views = 37, 70, 90, 8, 2
    rng = np.random.Generator(np.random.PCG64(5))
    frames = torch.from_numpy(rng.integers(0, 256, (f, h, w, 3), dtype=np.uint8))
    idx = sample_tta_view_indices(f, t, views)
    crop = (3, 11, 64, 64)
    out = views_to_device(frames.to(cuda_device), idx, t, arch, crop)
    x = frames[torch.as_tensor(idx)][:, 3:67, 11:75, :].float() / 255.0           # (V*T, h, w, 3)
    x = (x - torch.tensor(synth.INPUT_MEAN)) / torch.tensor(synth.INPUT_STD)
    x = x.permute(0, 3, 1, 2)                                                      # (V*T, 3, h, w)
    if arch == "tanet":
        want = x.reshape(views * t * 3, 64, 64)
    else:
        want = x.reshape(views, t, 3, 64, 64).permute(0, 2, 1, 3, 4)
    torch.testing.assert_close(out.cpu(), want.contiguous(), rtol=1e-6, atol=1e-6)
