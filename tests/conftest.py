import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: CPU test that takes more than ~20 s")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


# Collection order of the GPU suite: the driver runs `pytest -m gpu -x`, so one late unit failure must not blank the
# model-level parity evidence.  Model goldens first, then the kernel unit files, then the rest.
_GPU_ORDER = ["test_gpu_tanet", "test_gpu_swin", "test_gpu_kernels", "test_gpu_gemm", "test_crops", "test_swin_loader",
              "test_views", "test_gpu_multi", "test_gpu_gemm_f16"]


def pytest_collection_modifyitems(config, items):
    def rank(item):
        mod = os.path.splitext(os.path.basename(str(item.fspath)))[0]
        if item.get_closest_marker("gpu") is None:
            return -1                       # CPU tests keep their place in front
        return _GPU_ORDER.index(mod) if mod in _GPU_ORDER else len(_GPU_ORDER)
    items.sort(key=rank)                    # stable: order inside a file is unchanged
