"""CPU study: which operand split keeps the adaptation step inside the 1e-4 parity budget?

The conv / linear contractions of the step run on the tensor cores with fp32 operands SPLIT into narrow pieces
(DESIGN.md section 3).  This tool emulates candidate splits on the CPU -- forward, data gradient and weight gradient of
every conv2d / linear in the TANet adaptation step -- by patching the oracle's ``F.conv2d`` / ``F.linear`` with an
autograd function that rounds the operands to the pieces, multiplies the pieces in float64 (so only the split error is
visible) and drops the products the scheme drops.  It then runs a few adaptation steps and reports the deviation of
loss, logits, hooked statistics and weight deltas from the plain fp32 oracle.

Test infrastructure only (imports oracle/): run here, results recorded in DESIGN.md section 3.

  python tools/split_numerics.py [--steps 3] [--res 64] [--frames 8]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import vitta_oracle as O  # noqa: E402
from vitta_b200 import synth  # noqa: E402


# ------------------------------------------------------------------------------------------------
# piece rounding
# ------------------------------------------------------------------------------------------------
def rna_tf32(x):
    """cvt.rna.tf32.f32: round to nearest (ties away) keeping 10 explicit mantissa bits."""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x1000) & ~0x1FFF
    return i.view(torch.float32)


def pieces(x, scheme):
    """fp32 tensor -> list of float64 piece tensors (already de-scaled) for the scheme."""
    x = x.detach().float()
    if scheme == "fp32":
        return [x.double()]
    if scheme.startswith("tf32"):
        hi = rna_tf32(x)
        lo = rna_tf32(x - hi)
        return [hi.double(), lo.double()]
    if scheme.startswith("bf16"):
        out, r = [], x.clone()
        for _ in range(3):
            p = r.to(torch.bfloat16).float()
            out.append(p.double())
            r = r - p
        return out
    if scheme.startswith("fp16"):
        # hi = fp16(x * s), lo = fp16((x*s - hi) * 2^11); s = per-tensor power of two putting amax just below 2^14
        amax = float(x.abs().max())
        # "fp16x3@E": amax lands near 2^E (default 14; smaller E = a looser amax bound, more underflow into subnormals)
        tgt = int(scheme.split("@")[1]) if "@" in scheme else 14
        s = 1.0 if amax == 0 else 2.0 ** (tgt - np.ceil(np.log2(amax)))
        xs = x * s
        hi = xs.to(torch.float16).float()
        if scheme.startswith("fp16u"):   # residual NOT rescaled: one accumulator serves all three products
            lo = (xs - hi).to(torch.float16).float()
            return [hi.double() / s, lo.double() / s]
        lo = ((xs - hi) * 2048.0).to(torch.float16).float()
        return [hi.double() / s, lo.double() / (2048.0 * s)]
    raise ValueError(scheme)


# products kept: (index of A piece, index of B piece)
TERMS = {
    "fp32": [(0, 0)],
    "tf32x1": [(0, 0)],
    "tf32x3": [(0, 0), (0, 1), (1, 0)],
    "bf16x3": [(0, 0), (0, 1), (1, 0)],
    "bf16x4": [(0, 0), (0, 1), (1, 0), (1, 1)],
    "bf16x6": [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)],
    "fp16x3": [(0, 0), (0, 1), (1, 0)],
    "fp16u3": [(0, 0), (0, 1), (1, 0)],
}


def contract(op, a, b, scheme):
    pa, pb = pieces(a, scheme), pieces(b, scheme)
    acc = None
    for i, j in TERMS[scheme.split("@")[0]]:
        t = op(pa[i], pb[j])
        acc = t if acc is None else acc + t
    return acc.float()


class SplitConv2d(torch.autograd.Function):
    scheme = "fp32"

    @staticmethod
    def forward(ctx, x, w, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding)
        return contract(lambda a, b: F.conv2d(a, b, None, stride, padding), x, w, SplitConv2d.scheme)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, padding = ctx.cfg
        sch = SplitConv2d.scheme
        dx = contract(lambda a, b: torch.nn.grad.conv2d_input(x.shape, b, a, stride, padding), dy, w, sch)
        dw = contract(lambda a, b: torch.nn.grad.conv2d_weight(b, w.shape, a, stride, padding), dy, x, sch)
        return dx, dw, None, None


class SplitLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return contract(lambda a, b: a @ b.t(), x, w, SplitConv2d.scheme)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        sch = SplitConv2d.scheme
        dx = contract(lambda a, b: a @ b, dy, w, sch)
        dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
        dw = contract(lambda a, b: a.t() @ b, dy2, x2, sch)
        return dx, dw


class SplitMatmul(torch.autograd.Function):
    """a @ b for the attention products (q k^T, p v) with every operand -- gradients included -- split."""

    @staticmethod
    def forward(ctx, a, b):
        ctx.save_for_backward(a, b)
        return contract(lambda x, y: torch.matmul(x, y), a, b, SplitConv2d.scheme)

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        sch = SplitConv2d.scheme
        ga = contract(lambda x, y: torch.matmul(x, y.transpose(-2, -1)), g, b, sch)
        gb = contract(lambda x, y: torch.matmul(x.transpose(-2, -1), y), a, g, sch)
        return ga, gb


_ATTENTION = False       # --attention: also run the 4-D batched matmuls of window attention through the split
_orig_matmul = torch.Tensor.__matmul__


def _patched_matmul(self, other):
    if _ATTENTION and SplitConv2d.scheme != "fp32" and self.dim() == 4 and other.dim() == 4:
        return SplitMatmul.apply(self, other)
    return _orig_matmul(self, other)


class PatchedF:
    """Stand-in for torch.nn.functional inside the oracle: dense conv2d / linear go through the split emulation;
    the 3-channel stem, grouped convs and everything else stay plain fp32 (they are not on the tensor-core kernels)."""

    def __getattr__(self, name):
        return getattr(F, name)

    @staticmethod
    def conv2d(x, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
        if groups != 1 or x.shape[1] < 8 or SplitConv2d.scheme == "fp32":
            return F.conv2d(x, w, bias, stride, padding, dilation, groups)
        assert bias is None
        st = (stride, stride) if isinstance(stride, int) else tuple(stride)
        pd = (padding, padding) if isinstance(padding, int) else tuple(padding)
        return SplitConv2d.apply(x, w, st, pd)

    @staticmethod
    def linear(x, w, bias=None):
        if SplitConv2d.scheme == "fp32" or x.shape[-1] < 32:
            return F.linear(x, w, bias)
        y = SplitLinear.apply(x, w)
        return y if bias is None else y + bias


def run(scheme, sd, src, clip, steps, T, N, res, lr, swin=None):
    SplitConv2d.scheme = scheme
    O.F = PatchedF()
    torch.Tensor.__matmul__ = _patched_matmul
    try:
        if swin is None:
            st = O.TTAState(sd, "tanet", T, src[0], src[1], ["layer3", "layer4"], "l1_loss", True, 0.1, lr=lr)
        else:   # Video-Swin: the Linear layers (qkv / proj / fc1 / fc2 / reduction) go through the split; attention stays fp32
            st = O.TTAState(sd, "swin", T, src[0], src[1], swin["chosen"], "l1_loss", True, 0.05, lr=lr,
                            swin_cfg=dict(depths=tuple(swin["depths"]), heads=tuple(swin["heads"]),
                                          window=tuple(swin["window"])), name_prefix="module.")
        out = []
        for s in range(steps):
            r = st.adapt_step(clip[s], N, 1 if swin is None else swin["M"], False)
            stats = {k: (t.mean_meter.avg.detach().clone(), t.var_meter.avg.detach().clone())
                     for k, t in st.taps.items() if t.kind != "bn1d"}
            out.append((float(r["loss_reg"]), r["logits"].clone(), stats))
        w = {k: v.detach().clone() for k, v in st.sd.items() if v.requires_grad}
        return out, w
    finally:
        O.F = F
        torch.Tensor.__matmul__ = _orig_matmul


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--res", type=int, default=64)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--schemes", default="tf32x3,bf16x3,bf16x4,bf16x6,fp16x3,tf32x1")
    ap.add_argument("--arch", default="tanet", choices=["tanet", "swin"])
    ap.add_argument("--attention", action="store_true", help="swin: split the q k^T and p v products (fwd + bwd) too")
    a = ap.parse_args()
    global _ATTENTION
    _ATTENTION = a.attention
    torch.set_num_threads(os.cpu_count())
    swin = None
    if a.arch == "tanet":
        from vitta_b200.models.tanet_models.tanet import TSN
        K, T, N, res = 11, a.frames, 1, a.res
        model = TSN(K, T, 'RGB', base_model='resnet50', consensus_type='avg', img_feature_dim=256, tam=True,
                    non_local=False, partial_bn=False)
        sd = synth.synth_state_dict(model.state_dict(), seed=1)
        clean = synth.tanet_loader_tensor(synth.synth_video(N, 1, T, res, seed=100, gauss_sigma=0.0, tag="clean"))
        src = O.collect_source_stats(sd, "tanet", T, [clean.view(N, T, 3, res, res)])
        clips = [synth.tanet_loader_tensor(synth.synth_video(N, 1, T, res, seed=200 + s, tag="tta")).view(N, T, 3, res, res)
                 for s in range(a.steps)]
        title = f"TANet-R50 {N}x{T}x{res}x{res}"
    else:
        import cases
        swin = dict(cases.SWIN_CASES["swin_tiny_t32_r56_stats_l1"])
        K, T, N, res = swin["K"], swin["T"], swin["N"], swin["res"]
        swin["M"] = 1
        sd = synth.synth_state_dict(cases.swin_state_template(K, swin["embed_dim"], swin["depths"], swin["heads"],
                                                              swin["window"]), seed=1)
        cfg = dict(depths=tuple(swin["depths"]), heads=tuple(swin["heads"]), window=tuple(swin["window"]))
        clean = synth.swin_loader_tensor(synth.synth_video(N, 1, T, res, seed=100, gauss_sigma=0.0, tag="clean"))
        src = O.collect_source_stats(sd, "swin", T, [clean], cfg)
        clips = [synth.swin_loader_tensor(synth.synth_video(N, 1, T, res, seed=200 + s, tag="tta")) for s in range(a.steps)]
        title = f"Video-Swin (embed {swin['embed_dim']}, depths {swin['depths']}) {N}x{T}x{res}x{res}"
    ref, wref = run("fp32", sd, src, clips, a.steps, T, N, res, a.lr, swin)
    w0 = {k: v for k, v in sd.items() if k in wref}
    print(f"{title}, {a.steps} adaptation steps, lr {a.lr}; deviation from the fp32 oracle")
    print("| scheme | loss_reg rel | logits max rel (vs max|logit|) | EMA mean (max abs / layer scale) | EMA var max rel "
          "| weight-delta rel (l2, all tensors) |")
    print("|---|---|---|---|---|---|")
    for sch in a.schemes.split(","):
        got, wg = run(sch, sd, src, clips, a.steps, T, N, res, a.lr, swin)
        e_loss = max(abs(g[0] - r[0]) / abs(r[0]) for g, r in zip(got, ref))
        e_log = max(float((g[1] - r[1]).abs().max() / r[1].abs().max()) for g, r in zip(got, ref))
        e_mu = e_var = 0.0
        for g, r in zip(got, ref):
            for k in r[2]:
                mr, vr = r[2][k]
                mg, vg = g[2][k]
                scale = float(vr.mean().sqrt())
                e_mu = max(e_mu, float((mg - mr).abs().max()) / scale)
                e_var = max(e_var, float(((vg - vr).abs() / vr.abs().clamp_min(1e-12)).max()))
        num = sum(float(((wg[k] - wref[k]).double() ** 2).sum()) for k in wref)
        den = sum(float(((wref[k] - w0[k]).double() ** 2).sum()) for k in wref)
        print(f"| {sch} | {e_loss:.2e} | {e_log:.2e} | {e_mu:.2e} | {e_var:.2e} | {np.sqrt(num / max(den, 1e-300)):.2e} |")


if __name__ == "__main__":
    main()
