"""One convolution forward + backward on the tcgen05 kernels (ncu target):
  python tools/one_conv.py CIN COUT K STRIDE H [FRAMES] [FORM]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    cin, cout, k, s, h = [int(v) for v in sys.argv[1:6]]
    f = int(sys.argv[6]) if len(sys.argv) > 6 else 128
    form = int(sys.argv[7]) if len(sys.argv) > 7 else 0
    from vitta_b200 import _lib, ops
    dev = torch.device("cuda:0")
    _lib.call("vitta_gemm_set_operand_form", form)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(f, cin, h, h, generator=g).to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev).requires_grad_(True)
    for _ in range(2):
        y = ops.conv2d(x, w, s, k // 2)
        y.backward(torch.ones_like(y))
        x.grad = w.grad = None
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
