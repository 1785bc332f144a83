"""Deterministic synthetic weights / videos / source statistics.

There is no network in the build or bench environment, so neither the reference's checkpoints nor
its corrupted UCF-101 videos exist.  Everything is generated from numpy's PCG64 (stable across
machines and numpy versions), keyed by the state-dict entry name so that a reference model, the
oracle and the B200 model all receive bit-identical tensors regardless of construction order.
"""
import zlib

import numpy as np
import torch

INPUT_MEAN = (0.485, 0.456, 0.406)   # reference utils/opts.py:4
INPUT_STD = (0.229, 0.224, 0.225)    # reference utils/opts.py:5


def _rng(seed, key):
    return np.random.Generator(np.random.PCG64([seed & 0xFFFFFFFF, zlib.crc32(key.encode())]))


def synth_tensor_for(key, shape, seed=0):
    """One state-dict entry.  Conv/linear weights are He-scaled so 50 layers stay in fp32 range; norm
    layers get non-trivial affine parameters and running statistics so the folded affine map is
    exercised; the last norm of each residual branch is damped to keep the residual stream bounded."""
    r = _rng(seed, key)
    shape = tuple(shape)
    leaf = key.rsplit(".", 1)[-1]
    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.long)
    if leaf == "relative_position_index":
        raise KeyError(key)  # derived buffer, never synthesised
    if leaf == "running_mean":
        a = r.normal(0.0, 0.1, shape)
    elif leaf == "running_var":
        a = r.uniform(0.5, 2.0, shape)
    elif leaf == "relative_position_bias_table":
        a = r.normal(0.0, 0.2, shape)
    elif len(shape) == 1 and leaf == "weight":          # norm scale
        damp = key.endswith("bn3.weight") or ".downsample.1.weight" in key
        a = r.uniform(0.2, 0.5, shape) if damp else r.uniform(0.6, 1.4, shape)
    elif len(shape) == 1 and leaf == "bias":
        a = r.normal(0.0, 0.1, shape)
    elif leaf == "weight":
        fan_in = int(np.prod(shape[1:]))
        gain = 1.0 if ("new_fc" in key or "fc_cls" in key or ".tam." in key or "attn" in key
                       or "mlp" in key or "reduction" in key) else np.sqrt(2.0)
        a = r.normal(0.0, gain / np.sqrt(fan_in), shape)
    else:
        a = r.normal(0.0, 0.1, shape)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def synth_state_dict(template, seed=0):
    """template: mapping name -> tensor (only shape/dtype are read).  Returns a new state dict."""
    out = {}
    for k, v in template.items():
        if k.endswith("relative_position_index"):
            out[k] = v.clone()
        else:
            out[k] = synth_tensor_for(k, v.shape, seed)
    return out


def synth_video(n_videos, n_views, clip_len, size=224, seed=0, gauss_sigma=0.38, tag="tta"):
    """(N, M, T, 3, H, W) fp32, normalised.  ``gauss_sigma`` > 0 gives the 'gauss-corrupted' input:
    clip(u + N(0, sigma^2), 0, 1) with u ~ U[0,1).  Views of one video share the clean frames of a
    longer source video sampled at different temporal offsets (the reference's uniform_equidist idea,
    models/tanet_models/video_dataset.py:178-186) -- here simply independent draws per view."""
    r = _rng(seed, "video/%s" % tag)
    u = r.random((n_videos, n_views, clip_len, 3, size, size), dtype=np.float32)
    if gauss_sigma > 0:
        u = np.clip(u + r.normal(0.0, gauss_sigma, u.shape).astype(np.float32), 0.0, 1.0)
    mean = np.asarray(INPUT_MEAN, np.float32).reshape(1, 1, 1, 3, 1, 1)
    std = np.asarray(INPUT_STD, np.float32).reshape(1, 1, 1, 3, 1, 1)
    return torch.from_numpy(((u - mean) / std).astype(np.float32))


def tanet_loader_tensor(video):
    """(N, M, T, 3, H, W) -> the TANet loader layout (N, M*T*3, H, W) (basics.py:619-621)."""
    n, m, t, c, h, w = video.shape
    return video.reshape(n, m * t * c, h, w)


def swin_loader_tensor(video):
    """(N, M, T, 3, H, W) -> the Swin loader layout (N, M, 3, T, H, W) (basics.py:624-625)."""
    return video.permute(0, 1, 3, 2, 4, 5).contiguous()


def synth_labels(n_videos, num_classes, seed=0):
    r = _rng(seed, "labels")
    return torch.from_numpy(r.integers(0, num_classes, (n_videos,)).astype(np.int64))


def perturb_stats(mean_list, var_list, seed=0, rel=0.15):
    """Fabricated 'source' statistics when no clean pass is affordable (bench): the measured test
    statistics perturbed by +-rel so that every sign()/difference in the loss is non-degenerate."""
    outm, outv = [], []
    for i, (m, v) in enumerate(zip(mean_list, var_list)):
        r = _rng(seed, "srcstat/%d" % i)
        m = np.asarray(m, np.float32)
        v = np.asarray(v, np.float32)
        outm.append((m + rel * np.sqrt(v) * r.normal(size=m.shape)).astype(np.float32))
        outv.append((v * np.exp(rel * r.normal(size=v.shape))).astype(np.float32))
    return outm, outv
