"""Device time (CUPTI) of the stem pool pair and the TAM stencil pair at the shapes of the TANet bench step
(8 clips x 16 frames): bytes moved / time against the copy peak.  An L2 flush (256 MB fill) precedes every launch.
Usage: python tools/hbm_probe.py            (or under ncu: -k regex:'bn_relu_pool|tam_' ... python tools/hbm_probe.py --once)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def cupti_us(fn, flush, match, reps=3):
    from torch.profiler import profile, ProfilerActivity
    best = {}
    for it in range(reps):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        for e in prof.key_averages():
            for m in match:
                if m in e.key:
                    t = getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
                    best[m] = min(best.get(m, 1e30), t) if it else t
    return best


def main():
    import vitta_b200
    from vitta_b200 import ops
    once = "--once" in sys.argv
    dev = torch.device("cuda:0")
    CL = torch.channels_last
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    # stem: BN + ReLU + MaxPool of the 128 x 64 x 112 x 112 conv1 output
    f, c, h, w = 128, 64, 112, 112
    x = torch.randn(f, c, h, w, device=dev, generator=g).contiguous(memory_format=CL).requires_grad_(True)
    bw, bb = torch.rand(c, device=dev) + 0.5, torch.randn(c, device=dev) * 0.3
    rm, rv = torch.randn(c, device=dev) * 0.2, torch.rand(c, device=dev) + 0.5
    bw.requires_grad_(True); bb.requires_grad_(True)
    go = torch.randn(f, c, h // 2, w // 2, device=dev, generator=g).contiguous(memory_format=CL)

    def stem():
        out = ops.BnReluPoolFn.apply(x, bw, bb, rm, rv, 1e-5)
        out.backward(go)
    nb_in, nb_out = f * c * h * w * 4, f * c * (h // 2) * (w // 2) * 4
    if once:
        stem()
    else:
        t = cupti_us(stem, flush, ["bn_relu_pool_fwd", "bn_relu_pool_bwd"])
        print("bn_relu_pool_fwd  %7.1f us  %5.0f GB/s   (x read + pooled write + code)" % (
            t["bn_relu_pool_fwd"], (nb_in + nb_out * 1.25) / t["bn_relu_pool_fwd"] / 1e3))
        print("bn_relu_pool_bwd  %7.1f us  %5.0f GB/s   (x + pooled gradient + code read, gx write)" % (
            t["bn_relu_pool_bwd"], (2 * nb_in + nb_out * 1.25) / t["bn_relu_pool_bwd"] / 1e3))
    del x, go
    # TAM stencil: one per bottleneck, on the conv1 output
    for c, hw in ((64, 56), (128, 28), (256, 14), (512, 7)):
        n, T = 8, 16
        x = torch.randn(n * T, c, hw, hw, device=dev, generator=g).contiguous(memory_format=CL).requires_grad_(True)
        kern = torch.randn(n, 3, c, device=dev, generator=g).requires_grad_(True)
        act = torch.rand(n, T, c, device=dev, generator=g).requires_grad_(True)
        go = torch.randn(n * T, c, hw, hw, device=dev, generator=g).contiguous(memory_format=CL)

        def tam():
            out = ops.TamStencilFn.apply(x, kern, act, T)
            out.backward(go)
        nb = n * T * c * hw * hw * 4
        if once:
            tam()
            continue
        t = cupti_us(tam, flush, ["tam_fwd", "tam_bwd_kernel", "tam_bwd_finish"])
        print("C %3d @%2d  tam_fwd %6.1f us %5.0f GB/s | tam_bwd %6.1f us %5.0f GB/s | finish %5.1f us" % (
            c, hw, t["tam_fwd"], 2 * nb / t["tam_fwd"] / 1e3, t["tam_bwd_kernel"], 3 * nb / t["tam_bwd_kernel"] / 1e3,
            t["tam_bwd_finish"]))
        del x, go
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
