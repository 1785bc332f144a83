"""Per-view random multi-scale crop + bilinear resize (reference:
models/tanet_models/transforms.py:277-384, Pillow 8.4.0 Resample.c for the resize arithmetic).

CPU: the crop-box sampler and the oracle's Pillow restatement against vectors recorded from the unmodified reference
transform (tests/golden/crops.npz, oracle/make_golden.py::run_crops_case) and against the installed Pillow; the library's
host-side coefficient tables against the oracle's.  GPU: the fused gather + crop + resize + normalise kernel against the
oracle.  Everything up to the float normalisation is integer work: bit exact."""
import ctypes as C
import os
import random

import numpy as np
import pytest
import torch

import cases

GOLDEN = os.path.join(cases.GOLDEN_DIR, "crops.npz")


def test_crop_boxes_match_reference_sampler():
    from oracle import pil_resample as R
    from vitta_b200.corpus.views import sample_multiscale_crop
    g = np.load(GOLDEN)
    keys = [k for k in g.files if k.startswith("boxes/")]
    assert len(keys) == 8
    for key in keys:
        _, iw, ih, inp, seed = key.split("/")
        want = g[key]
        r1, r2 = random.Random(int(seed)), random.Random(int(seed))
        got = np.asarray([sample_multiscale_crop(int(iw), int(ih), int(inp), r1) for _ in range(len(want))])
        ora = np.asarray([R.sample_crop(int(iw), int(ih), int(inp), r2) for _ in range(len(want))])
        assert (got == want).all(), key
        assert (ora == want).all(), key
        assert (got[:, 2] + got[:, 0] <= int(iw)).all() and (got[:, 3] + got[:, 1] <= int(ih)).all()


def test_global_random_stream_is_the_default():
    """Like the reference, the sampler draws from the ``random`` module unless told otherwise."""
    from vitta_b200.corpus.views import sample_view_crops
    random.seed(21)
    a = sample_view_crops(320, 240, 224, 3)
    b = sample_view_crops(320, 240, 224, 3, random.Random(21))
    assert a == b and len(a) == 3


def test_oracle_resize_matches_reference_transform_golden():
    from oracle import pil_resample as R
    g = np.load(GOLDEN)
    for name in ("a", "b", "c"):
        inp, views, t, _ = (int(v) for v in g["xf/%s/meta" % name])
        frames, boxes, want = g["xf/%s/frames" % name], g["xf/%s/boxes" % name], g["xf/%s/out" % name]
        got = R.crop_resize_views(frames, np.arange(views * t), t, [tuple(int(x) for x in b) for b in boxes], inp)
        assert got.dtype == np.uint8 and got.shape == want.shape
        assert (got == want).all(), name


def test_oracle_resize_matches_installed_pillow():
    Image = pytest.importorskip("PIL.Image")
    from oracle import pil_resample as R
    rng = np.random.Generator(np.random.PCG64(3))
    for h, w, oh, ow in [(240, 320, 224, 224), (180, 180, 224, 224), (158, 210, 224, 224), (256, 340, 224, 224),
                         (37, 53, 16, 16), (20, 20, 64, 48), (224, 180, 224, 224), (300, 224, 224, 224), (9, 7, 23, 31),
                         (1, 1, 4, 4), (64, 64, 1, 1)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
        assert (R.resize_bilinear_u8(img, ow, oh) == want).all(), (h, w, oh, ow)
    flat = np.full((30, 40, 3), 255, np.uint8)            # saturation: weights sum to one, no overflow past 255
    assert (R.resize_bilinear_u8(flat, 17, 53) == 255).all()


def test_scale_center_crop_matches_reference_transform_golden():
    """GroupScale_TANet + GroupCenterCrop_TANet recorded from the reference: the oracle's restatement, and the product's
    table slicing (rows [left, left + S) of the whole-frame resize) applied with the oracle's pass -- both bit exact."""
    from oracle import pil_resample as R
    from vitta_b200.corpus.views import scale_center_crop_geometry, scale_center_crop_tables
    g = np.load(GOLDEN)
    for name in ("a", "b", "c", "d", "e"):
        z, inp = (int(v) for v in g["sc/%s/meta" % name])
        frames, want = g["sc/%s/frames" % name], g["sc/%s/out" % name]
        h, w = frames.shape[1:3]
        hb, hk, vb, vk, slots = scale_center_crop_tables(w, h, z, inp, n_views=2)
        assert hb.shape == (2, inp, 2) and hk.shape == (2, inp, slots) and (hb[0] == hb[1]).all()
        ow, oh, left, top = scale_center_crop_geometry(w, h, z, inp)
        assert min(ow, oh) == z and 0 <= left <= ow - inp and 0 <= top <= oh - inp
        for f, wf in zip(frames, want):
            assert (R.scale_center_crop_u8(f, z, inp) == wf).all(), name
            got = R._pass(R._pass(f, hb[0], hk[0], axis=1), vb[0], vk[0], axis=0)
            assert (got == wf).all(), name


def test_scale_center_crop_smaller_than_the_crop_is_loud():
    from vitta_b200 import _lib
    from vitta_b200.corpus.views import scale_center_crop_geometry
    with pytest.raises(_lib.VittaError):
        scale_center_crop_geometry(64, 48, 24, 32)


@pytest.mark.parametrize("offset", [0, 7])
def test_library_coefficient_tables_match_oracle(offset):
    from oracle import pil_resample as R
    from vitta_b200.corpus.views import resample_tables
    for i, o in [(240, 224), (180, 224), (158, 224), (210, 224), (224, 224), (256, 224), (512, 112), (7, 31), (1000, 33),
                 (3, 3), (224, 112), (1, 5), (5, 1), (31, 32), (33, 32)]:
        b, k = resample_tables(i, o, offset)
        bw, kw = R.resample_coeffs(i, o)
        assert k.shape[1] == R.resample_ksize(i, o)
        assert (b[:, 0] == bw[:, 0] + offset).all() and (b[:, 1] == bw[:, 1]).all(), (i, o)
        assert (k == kw).all(), (i, o)
        b2, k2 = resample_tables(i, o, offset, slots=k.shape[1] + 2)     # padded slots stay zero
        assert (b2 == b).all() and (k2[:, :k.shape[1]] == k).all() and (k2[:, k.shape[1]:] == 0).all()
        assert (b[:, 0] - offset >= 0).all() and (b[:, 0] - offset + b[:, 1] <= i).all()


def test_coefficient_tables_bad_arguments_are_loud():
    from vitta_b200 import _lib
    from vitta_b200.corpus.views import resample_tables
    with pytest.raises(_lib.VittaError):
        resample_tables(0, 5)
    with pytest.raises(_lib.VittaError):
        resample_tables(10, 5, slots=2)          # fewer slots than vitta_resample_ksize(10, 5) = 5
    with pytest.raises(_lib.VittaError):
        resample_tables(10, 5, in_offset=-1)


def test_crop_resize_entry_validates_before_touching_the_device():
    """No GPU here: a well-formed call gets as far as the launch (and fails there), a bad box is refused before that."""
    if torch.cuda.is_available():
        pytest.skip("marshalling-only test: meant for the GPU-less container")
    from vitta_b200 import _lib
    from vitta_b200.corpus.views import crop_resize_tables
    f, h, w, t, v, s = 4, 48, 64, 2, 2, 32
    frames = torch.zeros(f, h, w, 3, dtype=torch.uint8)
    idx = torch.zeros(v * t, dtype=torch.int32)
    out = torch.empty(v * t * 3, s, s)
    m3, s3 = (C.c_float * 3)(0.4, 0.4, 0.4), (C.c_float * 3)(0.2, 0.2, 0.2)

    def go(boxes):
        hb, hk, vb, vk, slots = crop_resize_tables(boxes, s, s)
        tabs = [torch.from_numpy(x) for x in (hb, hk, vb, vk)]
        bx = (C.c_int32 * (4 * v))(*[int(x) for b in boxes for x in b])
        _lib.call("vitta_gather_crop_resize_normalize_u8", _lib.ptr(frames), f, h, w, _lib.ptr(idx), v * t, bx, v,
                  _lib.ptr(tabs[0]), _lib.ptr(tabs[1]), _lib.ptr(tabs[2]), _lib.ptr(tabs[3]), slots, s, s, m3, s3, 0, t,
                  _lib.ptr(out), C.c_void_p(0))

    with pytest.raises(_lib.VittaError) as e:
        go([(48, 42, 16, 6), (36, 36, 0, 0)])
    assert "outside the frame" not in str(e.value)
    with pytest.raises(_lib.VittaError, match="outside the frame"):
        go([(48, 42, 17, 6), (36, 36, 0, 0)])        # 17 + 48 > 64


def test_views_to_device_has_no_cpu_path():
    from vitta_b200 import _lib
    from vitta_b200.corpus.views import views_to_device
    with pytest.raises(_lib.VittaError):
        views_to_device(torch.zeros(4, 48, 64, 3, dtype=torch.uint8), [0, 1], 2, boxes=[(32, 32, 0, 0)], out_size=32)


def _emulate_kernel(frames, idx, t, hb, hk, vb, vk, slots, out_h, out_w, layout):
    """numpy transcription of gather_crop_resize_kernel (preprocess.cu), index for index: per output pixel the <= slots
    source rows are resampled horizontally to uint8, then combined vertically; planes ordered as the two loader layouts."""
    f_total, h, w, _ = frames.shape
    n = len(idx)
    v_count = n // t
    u8 = np.zeros((n, out_h, out_w, 3), np.int64)
    for k in range(n):
        v = k // t
        f = min(max(int(idx[k]), 0), f_total - 1)
        for y in range(out_h):
            y0, ny = int(vb[v, y, 0]), min(int(vb[v, y, 1]), slots)
            for x in range(out_w):
                x0, nx = int(hb[v, x, 0]), min(int(hb[v, x, 1]), slots)
                acc = np.full(3, 1 << 21, np.int64)
                for yy in range(ny):
                    sy = min(max(y0 + yy, 0), h - 1)
                    hacc = np.full(3, 1 << 21, np.int64)
                    for xx in range(nx):
                        sx = min(max(x0 + xx, 0), w - 1)
                        hacc += frames[f, sy, sx].astype(np.int64) * int(hk[v, x, xx])
                    acc += np.clip(hacc >> 22, 0, 255) * int(vk[v, y, yy])
                u8[k, y, x] = np.clip(acc >> 22, 0, 255)
    if layout == 0:
        return u8.transpose(0, 3, 1, 2).reshape(n * 3, out_h, out_w)                       # [view][frame][rgb] planes
    return u8.reshape(v_count, t, out_h, out_w, 3).transpose(0, 4, 1, 2, 3)                # (V, 3, T, h, w)


def test_kernel_index_arithmetic_emulated_against_oracle():
    """The CUDA kernel cannot run here; its index arithmetic, transcribed to numpy, must reproduce the oracle through the
    same tables the product uploads (crop offsets folded into the window starts, per-view table sets, both layouts)."""
    from oracle import pil_resample as R
    from vitta_b200.corpus.views import crop_resize_tables, scale_center_crop_tables
    rng = np.random.Generator(np.random.PCG64(8))
    frames = rng.integers(0, 256, (5, 30, 40, 3), dtype=np.uint8)
    t, s = 2, 12
    idx = np.asarray([0, 4, 2, 9])                                        # 9 is clamped to the last frame like the reference
    boxes = [(26, 30, 12, 0), (19, 22, 3, 6)]
    hb, hk, vb, vk, slots = crop_resize_tables(boxes, s, s)
    want = R.crop_resize_views(frames, np.minimum(idx, 4), t, boxes, s)   # (V*T, s, s, 3)
    got0 = _emulate_kernel(frames, idx, t, hb, hk, vb, vk, slots, s, s, 0)
    assert (got0 == want.transpose(0, 3, 1, 2).reshape(-1, s, s)).all()
    got1 = _emulate_kernel(frames, idx, t, hb, hk, vb, vk, slots, s, s, 1)
    assert (got1 == want.reshape(2, t, s, s, 3).transpose(0, 4, 1, 2, 3)).all()
    hb, hk, vb, vk, slots = scale_center_crop_tables(40, 30, 16, s, n_views=2)
    want = np.stack([R.scale_center_crop_u8(frames[min(int(i), 4)], 16, s) for i in idx])
    got = _emulate_kernel(frames, idx, t, hb, hk, vb, vk, slots, s, s, 0)
    assert (got == want.transpose(0, 3, 1, 2).reshape(-1, s, s)).all()
    # --test_crops 3: every frame once per crop, tables as 3 * n_clips "views" (crop-major)
    from vitta_b200.corpus.views import full_res_sample_tables
    hb, hk, vb, vk, slots = full_res_sample_tables(40, 30, 16, s, n_clips=2)
    assert hb.shape == (6, s, 2)
    per = np.stack([R.full_res_sample_u8(frames[min(int(i), 4)], 16, s) for i in idx])     # (L, 3, s, s, 3)
    want = per.transpose(1, 0, 2, 3, 4).reshape(-1, s, s, 3)
    got = _emulate_kernel(frames, np.tile(idx, 3), t, hb, hk, vb, vk, slots, s, s, 0)
    assert (got == want.transpose(0, 3, 1, 2).reshape(-1, s, s)).all()


# ----------------------------------------------------------------------------------------------
# whole loader items recorded from the unmodified reference (get_dataset_tanet -> Video_TANetDataSet.__getitem__ with an
# in-memory decoder): tests/golden/loader.npz
# ----------------------------------------------------------------------------------------------
LOADER = os.path.join(cases.GOLDEN_DIR, "loader.npz")
LOADER_CASES = ["tta_randcrop", "tta_center", "eval", "tta_3views", "tta_3crops"]


def _loader_args(g, case):
    from vitta_b200.utils.opts import default_args
    is_tta, views, rand_crop, seed, test_crops = (int(v) for v in g[case + "/meta"])
    args = default_args(arch="tanet", clip_length=4, input_size=32, scale_size=40, n_augmented_views=views,
                        if_sample_tta_aug_views=True, if_spatial_rand_cropping=bool(rand_crop), num_classes=11,
                        batch_size=1, test_crops=test_crops)
    return args, ("tta" if is_tta else "eval"), seed


def _oracle_item(frames, idx, boxes, t, scale_size, input_size, three_crops=False):
    """uint8 frames -> the reference's loader tensor (V*T*3, S, S) through the oracle's spatial pipeline."""
    from oracle import pil_resample as R
    from vitta_b200 import synth
    if boxes is not None:
        u8 = R.crop_resize_views(frames, idx, t, boxes, input_size)
    elif three_crops:         # [crop][frame]: all frames of the left crop, then right, then centre
        per = np.stack([R.full_res_sample_u8(frames[int(i)], scale_size, input_size) for i in idx])    # (L, 3, S, S, 3)
        u8 = per.transpose(1, 0, 2, 3, 4).reshape(-1, input_size, input_size, 3)
    else:
        u8 = np.stack([R.scale_center_crop_u8(frames[int(i)], scale_size, input_size) for i in idx])
    x = torch.from_numpy(u8).permute(0, 3, 1, 2).float().div(255)                  # ToTorchFormatTensor(div=True)
    x = x.reshape(-1, input_size, input_size)                                      # Stack: planes [view][frame][rgb]
    mean = torch.tensor(synth.INPUT_MEAN).repeat(x.shape[0] // 3)[:, None, None]   # GroupNormalize (transforms.py:627-650)
    std = torch.tensor(synth.INPUT_STD).repeat(x.shape[0] // 3)[:, None, None]
    return (x - mean) / std


@pytest.mark.parametrize("case", LOADER_CASES)
def test_dataset_plan_and_oracle_pipeline_match_reference_loader(case):
    """The product's host-side plan (frame indices + crop boxes, drawn in the reference's order from the same seed) fed
    through the oracle's PIL-exact pipeline reproduces the items of the reference's own loader."""
    from vitta_b200.corpus.views import DecodedVideoDataset
    g = np.load(LOADER)
    args, kind, seed = _loader_args(g, case)
    videos = [torch.from_numpy(g["video/v0"]), torch.from_numpy(g["video/v1"])]
    ds = DecodedVideoDataset(videos, [3, 7], args, kind, rng=random.Random(seed))
    assert len(ds) == 2
    for i, name in enumerate(("v0", "v1")):
        idx, boxes = ds.plan(i)
        want = g["%s/%s/x" % (case, name)]
        assert (boxes is not None) == (kind == "tta" and args.if_spatial_rand_cropping and args.test_crops == 1)
        got = _oracle_item(videos[i].numpy(), idx, boxes, args.clip_length, args.scale_size, args.input_size,
                           three_crops=args.test_crops == 3)
        assert got.shape == want.shape, (got.shape, want.shape)
        # same uint8 pixels, then the same three float32 operations: bit identical
        assert np.array_equal(got.numpy(), want)
        assert int(g["%s/%s/y" % (case, name)]) == ds.labels[i]


def test_dataset_refuses_what_it_does_not_mirror():
    from vitta_b200 import _lib
    from vitta_b200.corpus.views import DecodedVideoDataset
    from vitta_b200.utils.opts import default_args
    vids = [torch.zeros(8, 48, 64, 3, dtype=torch.uint8)]
    with pytest.raises(NotImplementedError):
        DecodedVideoDataset(vids, [0], default_args(arch="videoswintransformer"), "tta")
    with pytest.raises(NotImplementedError):
        DecodedVideoDataset(vids, [0], default_args(arch="tanet", test_crops=5), "tta")
    with pytest.raises(_lib.VittaError):
        DecodedVideoDataset(vids, [0, 1], default_args(arch="tanet"), "tta")
    ds = DecodedVideoDataset(vids, [0], default_args(arch="tanet", sample_style="random-1", clip_length=4), "eval")
    with pytest.raises(NotImplementedError):
        ds.plan(0)


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["tanet", "videoswintransformer"])
def test_views_to_device_crop_resize_vs_oracle(cuda_device, arch):
    from oracle import pil_resample as R
    from vitta_b200 import synth
    from vitta_b200.corpus.views import sample_tta_view_indices, sample_view_crops, views_to_device
    f, h, w, t, views, s = 37, 120, 160, 8, 2, 112
    rng = np.random.Generator(np.random.PCG64(5))
    frames = rng.integers(0, 256, (f, h, w, 3), dtype=np.uint8)
    idx = sample_tta_view_indices(f, t, views)
    for seed in (0, 1, 2, 3):
        boxes = sample_view_crops(w, h, s, views, random.Random(seed))
        out = views_to_device(torch.from_numpy(frames).to(cuda_device), idx, t, arch, boxes=boxes, out_size=s)
        u8 = R.crop_resize_views(frames, idx, t, boxes, s)                            # (V*T, s, s, 3) uint8, PIL-exact
        x = torch.from_numpy(u8).float() / 255.0
        x = ((x - torch.tensor(synth.INPUT_MEAN)) / torch.tensor(synth.INPUT_STD)).permute(0, 3, 1, 2)
        if arch == "tanet":
            want = x.reshape(views * t * 3, s, s)
        else:
            want = x.reshape(views, t, 3, s, s).permute(0, 2, 1, 3, 4)
        # one uint8 step is 1/(255*0.225) = 1.7e-2 after normalisation: 1e-5 means every pixel has the exact PIL value
        torch.testing.assert_close(out.cpu(), want.contiguous(), rtol=1e-6, atol=1e-5)


@pytest.mark.gpu
def test_views_to_device_scale_center_crop_vs_oracle(cuda_device):
    from oracle import pil_resample as R
    from vitta_b200 import synth
    from vitta_b200.corpus.views import sample_tta_view_indices, views_to_device
    f, h, w, t, views, z, s = 20, 96, 128, 4, 2, 80, 64
    rng = np.random.Generator(np.random.PCG64(9))
    frames = rng.integers(0, 256, (f, h, w, 3), dtype=np.uint8)
    idx = sample_tta_view_indices(f, t, views)
    out = views_to_device(torch.from_numpy(frames).to(cuda_device), idx, t, "tanet", scale_size=z, out_size=s)
    u8 = np.stack([R.scale_center_crop_u8(frames[int(i)], z, s) for i in idx])
    x = torch.from_numpy(u8).float() / 255.0
    x = ((x - torch.tensor(synth.INPUT_MEAN)) / torch.tensor(synth.INPUT_STD)).permute(0, 3, 1, 2)
    torch.testing.assert_close(out.cpu(), x.reshape(views * t * 3, s, s).contiguous(), rtol=1e-6, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("case", LOADER_CASES)
def test_decoded_video_dataset_vs_reference_loader_golden(cuda_device, case):
    """End to end on the GPU: DecodedVideoDataset items against the items of the unmodified reference loader."""
    from vitta_b200.corpus.views import DecodedVideoDataset
    g = np.load(LOADER)
    args, kind, seed = _loader_args(g, case)
    videos = [torch.from_numpy(g["video/v0"]).to(cuda_device), torch.from_numpy(g["video/v1"]).to(cuda_device)]
    ds = DecodedVideoDataset(videos, [3, 7], args, kind, rng=random.Random(seed))
    for i, name in enumerate(("v0", "v1")):
        x, y = ds[i]
        # K13 normalises with the reference's own three float32 operations: bit identical to the reference loader
        assert np.array_equal(x.cpu().numpy(), g["%s/%s/x" % (case, name)])
        assert y == int(g["%s/%s/y" % (case, name)])


def test_driver_loader_does_not_pin_or_fork_for_device_resident_datasets():
    """corpus.basics._loader keeps the reference's DataLoader settings, except that items already on the device
    (Decoded*VideoDataset.on_device) are neither pinned (torch refuses to pin CUDA tensors) nor produced in workers."""
    from vitta_b200.corpus.basics import _loader
    from vitta_b200.corpus.views import DecodedSwinVideoDataset, DecodedVideoDataset
    from vitta_b200.utils.opts import default_args
    args = default_args(arch="tanet", batch_size=3, workers=4)
    vids = [torch.zeros(8, 48, 64, 3, dtype=torch.uint8)] * 2
    dl = _loader(DecodedVideoDataset(vids, [0, 1], args, "tta"), args)
    assert dl.pin_memory is False and dl.num_workers == 0 and dl.batch_size == 3
    sargs = default_args(arch="videoswintransformer", batch_size=2, workers=4)
    assert _loader(DecodedSwinVideoDataset(vids, [0, 1], sargs, "eval"), sargs).num_workers == 0
    plain = _loader(torch.utils.data.TensorDataset(torch.zeros(4, 2)), args)
    assert plain.pin_memory is True and plain.num_workers == 4 and plain.batch_size == 3
