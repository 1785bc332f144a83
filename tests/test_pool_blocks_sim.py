"""CPU restatement of the index logic of the stem's BN + ReLU + MaxPool(3, 2, 1) kernels (csrc/stem.cu): the forward's
"first maximum of the raw BN output, ReLU applied to the maximum" rule and the backward's 2 x 2 pixel blocks (a block lies in
exactly the four windows (i..i+1) x (j..j+1); window positions 3 * (dy - 2a + 1) + (dx - 2b + 1)) against torch's CPU
autograd of BatchNorm2d(eval) -> ReLU -> MaxPool2d.  (The kernels themselves run against torch on the GPU:
tests/test_gpu_kernels.py::test_bn_relu_pool_one_pass_vs_torch.)"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F


def forward_codes(y):
    """y (H, W) raw BN output of one channel -> pooled (Ho, Wo) and the window position of the first maximum of y."""
    H, W = y.shape
    Ho, Wo = (H + 1) // 2, (W + 1) // 2
    out = np.zeros((Ho, Wo), np.float32)
    code = np.zeros((Ho, Wo), np.int64)
    for ho in range(Ho):
        for wo in range(Wo):
            m, c = -np.inf, 0
            for t in range(9):
                h, w = 2 * ho - 1 + t // 3, 2 * wo - 1 + t % 3
                if 0 <= h < H and 0 <= w < W and y[h, w] > m:
                    m, c = y[h, w], t
            out[ho, wo] = max(m, 0.0)          # ReLU on the maximum, not on the nine candidates
            code[ho, wo] = c
    return out, code


def backward_blocks(y, code, gpool):
    """dL/dy_relu_input by 2 x 2 blocks, the ReLU mask taken from y > 0."""
    H, W = y.shape
    Ho, Wo = code.shape
    g = np.zeros((H, W), np.float64)
    for i in range(Ho):
        for j in range(Wo):
            for px in range(4):
                dy, dx = px >> 1, px & 1
                h, w = 2 * i + dy, 2 * j + dx
                if h >= H or w >= W:
                    continue
                acc = 0.0
                for q in range(4):
                    a, b = q >> 1, q & 1
                    ry, rx = dy - 2 * a + 1, dx - 2 * b + 1
                    if ry < 0 or rx < 0 or i + a >= Ho or j + b >= Wo:
                        continue
                    if code[i + a, j + b] == 3 * ry + rx:
                        acc += gpool[i + a, j + b]
                g[h, w] = acc if y[h, w] > 0 else 0.0
    return g


@pytest.mark.parametrize("h,w", [(8, 8), (9, 7), (15, 22), (2, 3)])
def test_block_rule_routes_gradients_like_torch(h, w):
    g = torch.Generator().manual_seed(h * 100 + w)
    # quantised values: plenty of exact ties, positive and non-positive window maxima
    y = (torch.randint(-3, 4, (1, 1, h, w), generator=g).float() * 0.5).requires_grad_(True)
    ref = F.max_pool2d(F.relu(y), 3, 2, 1)
    go = torch.randn(ref.shape, generator=g)
    ref.backward(go)
    out, code = forward_codes(y.detach().numpy()[0, 0])
    np.testing.assert_array_equal(out, ref.detach().numpy()[0, 0])
    got = backward_blocks(y.detach().numpy()[0, 0], code, go.numpy()[0, 0].astype(np.float64))
    np.testing.assert_allclose(got, y.grad.numpy()[0, 0], rtol=0, atol=1e-6)


def test_every_pixel_of_a_block_lies_in_its_four_windows_only():
    """Window (ho, wo) covers rows 2ho-1 .. 2ho+1: pixel row 2i is in window row i only, pixel row 2i+1 in rows i and i+1."""
    for h in range(0, 40):
        rows = [ho for ho in range(0, 30) if 2 * ho - 1 <= h <= 2 * ho + 1]
        i, dy = divmod(h, 2)
        assert rows == ([i] if dy == 0 else [i, i + 1])
        for a, ho in enumerate(rows):
            assert h - (2 * ho - 1) == dy - 2 * a + 1
