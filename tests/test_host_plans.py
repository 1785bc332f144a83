"""Host-side planning logic that needs no GPU: the split-K plan of the weight-gradient kernel and the padded token volume
of the window attention (reference swin_transformer.py:71-84, 222-227)."""
import ctypes as C

import pytest

from vitta_b200 import _lib

SMS = 148
# (F, H, W, Cin, Cout, k, stride): every distinct ResNet-50 convolution of the TANet step at 128 frames x 224 x 224, and the
# linear layers of Video-Swin-T (rows as a 1 x 1 "image")
RESNET = [(128, 56, 56, 64, 64, 1, 1), (128, 56, 56, 64, 64, 3, 1), (128, 56, 56, 64, 256, 1, 1), (128, 56, 56, 256, 64, 1, 1),
          (128, 56, 56, 256, 128, 1, 1), (128, 56, 56, 128, 128, 3, 2), (128, 28, 28, 128, 512, 1, 1),
          (128, 56, 56, 256, 512, 1, 2), (128, 28, 28, 512, 128, 1, 1), (128, 28, 28, 128, 128, 3, 1),
          (128, 28, 28, 512, 256, 1, 1), (128, 28, 28, 256, 256, 3, 2), (128, 14, 14, 256, 1024, 1, 1),
          (128, 28, 28, 512, 1024, 1, 2), (128, 14, 14, 1024, 256, 1, 1), (128, 14, 14, 256, 256, 3, 1),
          (128, 14, 14, 1024, 512, 1, 1), (128, 14, 14, 512, 512, 3, 2), (128, 7, 7, 512, 2048, 1, 1),
          (128, 14, 14, 1024, 2048, 1, 2), (128, 7, 7, 2048, 512, 1, 1), (128, 7, 7, 512, 512, 3, 1)]
SWIN = [(1, 1, 802816, 96, 288, 1, 1), (1, 1, 802816, 96, 384, 1, 1), (1, 1, 802816, 384, 96, 1, 1),
        (1, 1, 200704, 192, 576, 1, 1), (1, 1, 50176, 384, 1536, 1, 1), (1, 1, 12544, 768, 3072, 1, 1),
        (1, 1, 12544, 3072, 768, 1, 1)]


def _plan(shape, f16=1):
    f, h, w, cin, cout, k, s = shape
    out = (C.c_int * 4)()
    rc = _lib.load().vitta_conv2d_wgrad_plan(f, h, w, cin, cout, k, k, s, k // 2, f16, out)
    assert rc == 0
    return tuple(out)


def _cost(base, boxes, s):
    waves = -(-base * s // SMS)
    return waves * (-(-boxes // s) + 6)


@pytest.mark.parametrize("shape", RESNET + SWIN)
@pytest.mark.parametrize("f16", [0, 1])
def test_wgrad_split_plan_minimises_the_makespan(shape, f16):
    """The persistent grid walks base x splits items round-robin over 148 SMs: the chosen split count has the smallest
    ceil(items / SMs) * (stages per item + fixed cost) of all admissible counts -- in particular never the "just over two
    items per SM" of the old rule (297 ... 360 items: a third pass with 1-64 busy CTAs)."""
    bn, base, splits, boxes = _plan(shape, f16)
    assert bn in (64, 128, 192, 256) and base >= 1 and splits >= 1 and boxes >= 1
    max_by_k = max(1, boxes // 8)
    assert splits <= max_by_k
    best = min(_cost(base, boxes, s) for s in range(1, min(max_by_k, 1024) + 1) if -(-base * s // SMS) <= 5)
    assert _cost(base, boxes, splits) == best
    old = min(max(-(-2 * SMS // base), 1), max_by_k, 1024)
    assert _cost(base, boxes, splits) <= _cost(base, boxes, old)
    items = base * splits
    waves = -(-items // SMS)
    # the last pass is never nearly empty unless a single pass cannot be filled at all
    assert waves == 1 or items - (waves - 1) * SMS >= SMS // 2 or base > SMS, (items, waves)


@pytest.mark.parametrize("dims,window,want", [
    ((2, 16, 56, 56), (8, 7, 7), None),                 # every 224 x 224 stage: multiples of the window
    ((2, 16, 7, 7), (8, 7, 7), None),
    ((2, 4, 7, 7), (8, 7, 7), None),                    # clamped window (4, 7, 7)
    ((2, 8, 10, 9), (8, 7, 7), (2, 8, 14, 14)),
    ((1, 12, 7, 16), (8, 7, 7), (1, 16, 7, 21)),
    ((3, 10, 14, 14), (8, 7, 7), (3, 16, 14, 14)),
    ((1, 3, 5, 20), (8, 7, 7), (1, 3, 5, 21)),          # D and H below the window: clamped, not padded
])
def test_padded_token_dims_follow_the_reference(dims, window, want):
    from vitta_b200 import ops_swin
    assert ops_swin.padded_token_dims(dims, window) == want
    # the oracle's restatement of get_window_size + F.pad agrees
    from oracle import vitta_oracle as O
    ws, _ = O.swin_window_and_shift(dims[1:], window, (0, 0, 0))
    ext = tuple(x + (ws_i - x % ws_i) % ws_i for x, ws_i in zip(dims[1:], ws))
    assert (want is None and ext == tuple(dims[1:])) or want == (dims[0],) + ext
