"""Per-shape timing of the TANet-R50 convolutions (BASELINE configs[1]: 8 clips x 16 frames = 128 frames) on the
tcgen05 3xTF32 kernels: forward, data gradient and weight gradient of every distinct convolution, with its
multiplicity in the network, for both A-operand forms of the GEMM kernel (tensor memory / shared memory).

  python tools/conv_shapes.py [--frames 128] [--reps 5] [--check]

Prints a markdown table (us per launch, algorithmic TFLOP/s = 2*M*N*K / t) and the per-step totals.  --check also
compares the two operand forms (to fp32 rounding) and the forward against a float64 convolution on a frame subset."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def r50_convs():
    """(name, cin, cout, k, stride, pad, H_in, multiplicity).  TemporalBottleneck = torchvision Bottleneck v1.5 with the
    TAM after conv1 (models/tanet_models/temporal_module.py:85-106): the stride sits on conv2 of block 0."""
    out = []
    inpl, res = 64, 56
    for li, (planes, blocks, stride) in enumerate([(64, 3, 1), (128, 4, 2), (256, 6, 2), (512, 3, 2)], start=1):
        for b in range(blocks):
            s = stride if b == 0 else 1
            r_in = res
            r_out = res // s
            tag = "layer%d.%s" % (li, "0" if b == 0 else "1+")
            out.append((tag + ".conv1", inpl, planes, 1, 1, 0, r_in))
            out.append((tag + ".conv2", planes, planes, 3, s, 1, r_in))
            out.append((tag + ".conv3", planes, planes * 4, 1, 1, 0, r_out))
            if b == 0:
                out.append((tag + ".downsample", inpl, planes * 4, 1, s, 0, r_in))
            inpl = planes * 4
            res = r_out
    merged = {}
    for name, cin, cout, k, s, p, r in out:
        key = (cin, cout, k, s, p, r)
        if key in merged:
            merged[key][1] += 1
        else:
            merged[key] = [name, 1]
    return [(v[0],) + k + (v[1],) for k, v in merged.items()]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--check", action="store_true")
    a = ap.parse_args()
    import vitta_b200
    from vitta_b200 import _lib, ops
    dev = torch.device("cuda:0")
    vitta_b200.set_fp32_exact()
    CL = torch.channels_last
    g = torch.Generator(device="cpu").manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, cin, cout, k, s, p, r, mult in r50_convs():
        f = a.frames
        x = torch.randn(f, cin, r, r, generator=g).to(dev).contiguous(memory_format=CL).requires_grad_(True)
        w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev).requires_grad_(True)
        ro = (r + 2 * p - k) // s + 1
        gy = torch.randn(f, cout, ro, ro, generator=g).to(dev).contiguous(memory_format=CL)
        flops = 2.0 * f * ro * ro * cout * cin * k * k
        res = {}
        outs = {}
        for form in (2, 1):
            _lib.call("vitta_gemm_set_operand_form", form)
            ops.bump_weight_epoch()
            for _ in range(2):   # warm-up (weight splits cached afterwards)
                y = ops.conv2d(x, w, s, p)
                y.backward(gy)
                x.grad = w.grad = None
            acc = {}
            for _ in range(a.reps):
                flush.zero_()
                _lib.profile = []
                y = ops.conv2d(x, w, s, p)
                fwd_n = len(_lib.profile)
                y.backward(gy)
                torch.cuda.synchronize()
                prof, _lib.profile = _lib.profile, None
                for i, (nm, e0, e1, _args) in enumerate(prof):
                    kind = "fwd" if i < fwd_n else ("wgrad" if "wgrad" in nm else "dgrad")
                    if "amax" in nm:      # VITTA_GEMM_PRECISION=f16x3 bring-up: standalone amax passes, reported apart
                        kind = "amax"
                    acc[kind] = acc.get(kind, 0.0) + e0.elapsed_time(e1) * 1e3 / a.reps
                outs[form] = (y.detach().clone(), x.grad.clone(), w.grad.clone())
                x.grad = w.grad = None
            res[form] = acc
        _lib.call("vitta_gemm_set_operand_form", 0)
        chk = ""
        if a.check:
            same = all(float((u - v).abs().max()) <= 5e-5 * float(v.abs().max()) for u, v in zip(outs[2], outs[1]))
            nf = min(f, 4)
            ref = torch.nn.functional.conv2d(x.detach()[:nf].double(), w.detach().double(), None, s, p)
            err = ((outs[1][0][:nf].double() - ref).abs().max() / ref.abs().max()).item()
            chk = " | %s | %.1e" % ("agree" if same else "DIFF", err)
        rows.append((name, cin, cout, k, s, r, mult, flops, res, chk))
        del x, w, gy, outs
    hdr = "| conv | Cin→Cout k/s @H | × | " + " | ".join("%s %s us (TF/s)" % (kd, fm) for kd in ("fwd", "dgrad", "wgrad")
                                                           for fm in ("tmem", "smem"))
    print(hdr + (" | forms | fwd err |" if a.check else " |"))
    print("|---|---|---:|" + "---:|" * 6 + ("---|---:|" if a.check else ""))
    tot = {(kd, fm): 0.0 for kd in ("fwd", "dgrad", "wgrad") for fm in (2, 1)}
    totf = 0.0
    for name, cin, cout, k, s, r, mult, flops, res, chk in rows:
        cells = []
        for kd in ("fwd", "dgrad", "wgrad"):
            for fm in (2, 1):
                us = res[fm].get(kd, 0.0)
                tot[(kd, fm)] += us * mult
                cells.append("%.0f (%.0f)" % (us, flops / us / 1e6 if us > 0 else 0))
        totf += flops * mult
        print("| %s | %d→%d %d/%d @%d | %d | %s%s |" % (name, cin, cout, k, s, r, mult, " | ".join(cells), chk))
    print()
    amax_us = sum(res[2].get("amax", 0.0) * mult for _n, _ci, _co, _k, _s, _r, mult, _f, res, _c in rows)
    print("operand split: %s%s" % (ops.gemm_precision(), "   (standalone amax passes: %.2f ms per step, not in the "
                                   "columns above)" % (amax_us / 1e3) if amax_us else ""))
    for kd in ("fwd", "dgrad", "wgrad"):
        print("total %-5s: tmem-A %.2f ms (%.0f TF/s)   smem-A %.2f ms (%.0f TF/s)" % (
            kd, tot[(kd, 2)] / 1e3, totf / tot[(kd, 2)] / 1e6, tot[(kd, 1)] / 1e3, totf / tot[(kd, 1)] / 1e6))


if __name__ == "__main__":
    main()
