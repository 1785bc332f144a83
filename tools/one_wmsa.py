"""One 3-D window attention forward + backward on the tcgen05 kernels (ncu target for K7):
  python tools/one_wmsa.py [VIEWS] [D] [H] [HEADS] [SHIFT]
Defaults = Video-Swin-T stage 1 of BASELINE.json configs[2] for 4 views (D = 16, 56 x 56 tokens, 3 heads, window (8,7,7)),
the geometry that carries most of the attention time (profiles/r01_swin_tiny_step_launches.md).  Prints CUDA-event times
of the forward and the two backward launches unless run under ncu (a number printed under ncu is never a bench value)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    a = [int(v) for v in sys.argv[1:]]
    views, d, h, heads, shifted = (a + [4, 16, 56, 3, 1][len(a):])[:5]
    from vitta_b200 import ops_swin
    dev = torch.device("cuda:0")
    window, shift = (8, 7, 7), ((4, 3, 3) if shifted else (0, 0, 0))
    c = heads * 32
    rows = views * d * h * h
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(rows, 3 * c, generator=g) * 1.2).to(dev)
    table = (torch.randn((2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1), heads, generator=g) * 0.5).to(dev)
    go = torch.randn(rows, c, generator=g).to(dev)
    scale = 32 ** -0.5
    dims = (views, d, h, h)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    for rep in range(3):
        ev[0].record()
        out, lse = ops_swin.wmsa3d_fwd(qkv, table, dims, heads, window, shift, scale)
        ev[1].record()
        ops_swin.wmsa3d_bwd(qkv, table, out, go, lse, dims, heads, window, shift, scale, 0)
        ev[2].record()
    torch.cuda.synchronize()
    n = 392
    win = rows // n
    fl_f = 4.0 * win * heads * n * n * 32              # QK^T + PV
    fl_b = 10.0 * win * heads * n * n * 32             # S, dP recomputed in both launches + dQ, dK, dV
    tf, tb = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    print("wmsa3d %d views x %d x %d x %d, %d heads, shift %s: fwd %.3f ms (%.1f TFLOP/s), bwd %.3f ms (%.1f TFLOP/s issued-equivalent)"
          % (views, d, h, h, heads, shift, tf, fl_f / tf / 1e9, tb, fl_b / tb / 1e9))


if __name__ == "__main__":
    main()
