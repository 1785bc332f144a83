"""STUDY TOOL (not on any product path): the reference's adaptation step executed by PyTorch's own CUDA kernels.

SURVEY.md section 8(d) asks for the "reference-on-GPU" number next to the CPU baseline -- the bar a user of the reference
on the same B200 would see: same Python-level algorithm, torch eager, cuDNN / cuBLAS kernels, fp32 with TF32 off (the
reference's numerics) and on.  The unmodified reference cannot run on the GPU box (/root/reference is not there, mmcv / timm
/ decord are absent), so this drives the oracle port (oracle/vitta_oracle.py, the functional restatement that is pinned
against the unmodified reference on the CPU) with its tensors on cuda:0.  BASELINE.json configs[1]: TANet-R50, 8 videos x
1 view x 16 x 224 x 224, 47 hooks (29 aligned), L1, SGD over all parameters.  Timed with CUDA events after warm-up.
Usage: python tools/reference_gpu_step.py [--videos 8] [--steps 10] [--warmup 3]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch


def run(n_videos, steps, warmup, tf32):
    import cases
    from oracle import vitta_oracle as O
    from vitta_b200 import synth
    K, T, RES = 101, 16, 224
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = True          # reference corpus/main_eval.py:77
    sd = {k: v.to(dev) for k, v in synth.synth_state_dict(cases.tanet_state_template(K, T), seed=1).items()}
    names = [n for n, k in O.tanet_norm_layers() if k != "bn1d"]
    src_m = [torch.zeros(sd[n + ".weight"].shape[0], device=dev) for n in names]
    src_v = [torch.ones(sd[n + ".weight"].shape[0], device=dev) for n in names]
    st = O.TTAState(sd, "tanet", T, src_m, src_v, ["layer3", "layer4"], "l1_loss", True, 0.1, lr=5e-5)
    x = synth.synth_video(n_videos, 1, T, RES, seed=200, tag="tta").view(n_videos, T, 3, RES, RES).to(dev)
    for _ in range(warmup):
        st.adapt_step(x, n_videos, 1, False, dropout_p=0.8)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        st.adapt_step(x, n_videos, 1, False, dropout_p=0.8)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"tf32": tf32, "ms_per_step": ms, "clips_per_s": n_videos * 1000.0 / ms,
            "mem_gb": torch.cuda.max_memory_allocated() / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--videos", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    a = ap.parse_args()
    out = {"what": "reference step (oracle port) on torch-CUDA eager kernels, cuda:0, TANet-R50 ViTTA, %d videos x 1 view x "
                   "16x224x224, L1 alignment on 47 hooks, SGD all params" % a.videos,
           "torch": torch.__version__, "runs": [run(a.videos, a.steps, a.warmup, tf32) for tf32 in (False, True)]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
