"""Time and attribute ONE ViTTA adaptation step of Video-Swin on cuda:0 (BASELINE.json configs[2] by default: Swin-T,
8 videos x 2 temporal views x 32 x 224 x 224, statistics alignment + prediction consistency).  Not the driver's bench
line (bench.py times configs[1]); this is the measurement tool for the Swin kernels (K7/K8/K9).
Usage: python tools/swin_step.py [--model tiny|base] [--videos 8] [--views 2] [--frames 32] [--steps 5] [--ncu-step]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench

CFG = {"tiny": dict(embed_dim=96, depths=[2, 2, 6, 2], num_heads=[3, 6, 12, 24]),
       "base": dict(embed_dim=128, depths=[2, 2, 18, 2], num_heads=[4, 8, 16, 32])}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="tiny")
    ap.add_argument("--videos", type=int, default=8)
    ap.add_argument("--views", type=int, default=2)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--drop-path", type=float, default=0.2)
    ap.add_argument("--ncu-step", action="store_true")
    a = ap.parse_args()
    import vitta_b200
    from vitta_b200 import _lib, synth
    from vitta_b200.corpus.basics import OnlineAdapter
    from vitta_b200.models.videoswintransformer_models.recognizer3d import Recognizer3D
    from vitta_b200.utils.opts import default_args
    dev = torch.device("cuda:0")
    vitta_b200.set_fp32_exact()
    cfg = CFG[a.model]
    model = Recognizer3D(num_classes=101, patch_size=(2, 4, 4), window_size=(8, 7, 7), drop_path_rate=a.drop_path, **cfg)
    model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=1))
    model = torch.nn.DataParallel(model.to(dev), device_ids=[0])
    lns = [m for n, m in model.named_modules() if isinstance(m, torch.nn.LayerNorm)][1:]
    src_m = [np.zeros(m.normalized_shape[0], np.float32) for m in lns]
    src_v = [np.ones(m.normalized_shape[0], np.float32) for m in lns]
    args = default_args(arch='videoswintransformer', clip_length=a.frames, batch_size=a.videos,
                        n_augmented_views=a.views, if_pred_consistency=a.views > 1,
                        if_sample_tta_aug_views=a.views > 1, lr=1e-5, momentum_mvg=0.05, lambda_pred_consis=0.05,
                        chosen_blocks=['module.backbone.layers.2', 'module.backbone.layers.3', 'module.backbone.norm'],
                        num_classes=101, input_size=224, num_clips=1)
    ad = OnlineAdapter(model, args, (src_m, src_v))
    x = synth.swin_loader_tensor(synth.synth_video(a.videos, a.views, a.frames, 224, seed=200, tag="tta")).to(dev)
    for _ in range(3):
        ad.adapt(x)
    torch.cuda.synchronize()
    if a.ncu_step:
        torch.cuda.cudart().cudaProfilerStart()
        ad.adapt(x)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count
    e0.record()
    for _ in range(a.steps):
        ad.adapt(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    launches = (_lib.launch_count - l0) // a.steps
    fam = bench.attribute_step(ad, x)
    table = {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                 **({"tflops": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 1)} if v["flops"] else {}),
                 **({"gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)} if v["bytes"] else {})}
             for k, v in sorted(fam.items(), key=lambda kv: -kv[1]["ms"])}
    views = a.videos * a.views
    print(json.dumps({"workload": "Video-Swin-%s ViTTA, %d videos x %d views x %dx224x224, %d hooks, drop_path %.2f"
                                  % (a.model, a.videos, a.views, a.frames, len(ad.stat_reg_hooks), a.drop_path),
                      "ms_per_step": ms, "clip_views_per_s": views * 1000.0 / ms, "videos_per_s": a.videos * 1000.0 / ms,
                      "gpu_launches_per_step": launches, "mem_gb": torch.cuda.max_memory_allocated() / 1e9,
                      "kernels": table}))


if __name__ == "__main__":
    main()
