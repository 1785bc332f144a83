"""Times ln_fwd / ln_bwd (K9) at the LayerNorm shapes of the Video-Swin-T step (16 clips x 16x56x56 tokens).
Device time of the ln_* launches from CUPTI (torch.profiler), after an L2 flush (a 256 MB fill).  Usage: python tools/ln_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    import vitta_b200
    from vitta_b200 import ops, ops_swin
    dev = torch.device("cuda:0")
    ops.set_gemm_precision("f16x3")
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    shapes = [(802816, 96), (200704, 192), (50176, 384), (12544, 768), (12544, 1536), (50176, 768), (200704, 384)]
    for rows, c in shapes:
        g = torch.Generator(device=dev).manual_seed(1)
        x = torch.randn(rows, c, device=dev, generator=g)
        gy = torch.randn(rows, c, device=dev, generator=g)
        gadd = torch.randn(rows, c, device=dev, generator=g)
        w = torch.rand(c, device=dev, generator=g) + 0.5
        b = torch.randn(c, device=dev, generator=g)
        coef = tuple(torch.randn(c, device=dev, generator=g) * 1e-3 for _ in range(3))
        gs = torch.ones(1, device=dev)
        y, mean, rstd = ops_swin.ln_fwd(x, w, b, 1e-5, rows, c)
        res = {}
        for name, fn in (("fwd", lambda: ops_swin.ln_fwd(x, w, b, 1e-5, rows, c)),
                         ("bwd", lambda: ops_swin.ln_bwd(gy, x, w, b, mean, rstd, rows, c)),
                         ("bwd+gadd+hook", lambda: ops_swin.ln_bwd(gy, x, w, b, mean, rstd, rows, c, gadd=gadd,
                                                                   coef=tuple(ops_swin.ptr(t) for t in coef), gscale=gs))):
            # device time of the launches themselves (CUPTI): CUDA events around the Python call would include the host-side
            # launch latency, which exceeds the kernel time of the small shapes
            from torch.profiler import profile, ProfilerActivity
            ts = []
            for it in range(3):
                flush.fill_(1.0)
                torch.cuda.synchronize()
                with profile(activities=[ProfilerActivity.CUDA]) as prof:
                    fn()
                    torch.cuda.synchronize()
                t = 0.0
                for e in prof.key_averages():
                    if "vitta::ln_" in e.key:
                        t += getattr(e, "device_time_total", None) or getattr(e, "cuda_time_total", 0)
                ts.append(t)
            res[name] = min(ts[1:])
        nb = rows * c * 4
        print("rows %7d C %4d  fwd %7.1f us %5.0f GB/s | bwd %7.1f us %5.0f GB/s | bwd+gadd+hook %7.1f us %5.0f GB/s" % (
            rows, c, res["fwd"], 2 * nb / res["fwd"] / 1e3, res["bwd"], 3 * nb / res["bwd"] / 1e3,
            res["bwd+gadd+hook"], 4 * nb / res["bwd+gadd+hook"] / 1e3))
        del x, gy, gadd, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
