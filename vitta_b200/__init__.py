"""vitta_b200 -- B200-native (sm_100a) kernels behind ViTTA's test-time-adaptation inner loop.

Layout
  csrc/ + libvitta_b200.so   hand-written CUDA kernels and their C ABI (include/vitta_b200.h)
  _lib.py                    ctypes binding (no fallback: missing library -> exception)
  ops.py, nn.py              autograd operators / fused norm call path
  utils/, models/, corpus/   host-side mirror of the reference's hook / model / driver interface
"""
__version__ = "0.1.0"


def set_fp32_exact():
    """ViTTA parity is specified in fp32 (BASELINE.json north_star: 1e-4 relative).  PyTorch lets cuDNN
    convolutions use TF32 by default; switch that (and TF32 matmuls) off for the library GEMM/conv calls."""
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
