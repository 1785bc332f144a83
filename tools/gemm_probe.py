"""Bring-up probe for the tcgen05 GEMM: small -> large, prints after every case (run under `timeout`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vitta_b200 import ops

dev = torch.device("cuda:0")
torch.manual_seed(0)
cases = [(128, 64, 32, 64), (128, 64, 32, 0), (128, 128, 32, 128), (128, 128, 128, 128), (256, 128, 64, 128), (128, 256, 64, 256),
         (1024, 256, 256, 0), (200, 96, 100, 0), (6272, 2048, 512, 0), (25088, 256, 1024, 0), (401408, 64, 64, 0), (401408, 256, 64, 0)]
for m, n, k, fb in cases:
    a = torch.randn(m, k, device=dev)
    b = torch.randn(n, k, device=dev) / k ** 0.5
    bh, bl = ops.split_tf32(b)
    torch.cuda.synchronize()
    out = ops.gemm_tf32x3(a, bh, bl, n, force_bn=fb)
    torch.cuda.synchronize()
    sub = slice(0, min(m, 2048))
    ref = a[sub].double() @ b.double().t()
    ap = a[sub].double().abs() @ b.double().abs().t()
    err = (out[sub].double() - ref).abs()
    rel = float((err / ap).max())
    ref32 = float(((a[sub] @ b.t()).double() - ref).abs().div(ap).max())
    # timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        ops.gemm_tf32x3(a, bh, bl, n, force_bn=fb, out=out)
    e0.record()
    it = 10
    for _ in range(it):
        ops.gemm_tf32x3(a, bh, bl, n, force_bn=fb, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / it
    print("M=%d N=%d K=%d bn=%d  max err/absprod %.2e (torch fp32: %.2e)  %.3f ms  %.1f TFLOP/s" %
          (m, n, k, fb, rel, ref32, ms, 2.0 * m * n * k / ms / 1e9), flush=True)
