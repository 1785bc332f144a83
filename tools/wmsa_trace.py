"""Timeline of the window-attention forward kernel (K7) from its own time stamps (vitta_wmsa3d_fwd_trace): for one item of
CTA 0, per tile, when each role passed its hand-over points, in SM clock cycles relative to the tile's S_FULL.
  python tools/wmsa_trace.py [--bwd] [VIEWS] [D] [H] [HEADS] [SHIFT] [ITEM]
--bwd: the two backward launches instead (vitta_wmsa3d_bwd_set_trace), query-outer first."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

NAMES = {1: "sm S_FULL passed", 2: "sm pass1 done", 3: "sm max exchanged", 10: "sm chunk computed", 11: "sm P_FREE passed",
         12: "sm P_READY arrived", 5: "sm sum exchanged", 6: "sm O_FULL passed", 7: "sm epilogue done",
         20: "pv P_READY passed", 21: "pv V_READY/O_FREE passed", 22: "pv chunk issued", 30: "s  issue start",
         31: "s  issued", 40: "ld item start", 41: "ld K ready", 42: "ld Q ready"}


BNAMES = {1: "row SC_FULL passed", 2: "row computed", 3: "row E_READY arrived", 6: "row ACC_FULL passed",
          20: "sc COL_READY passed", 21: "sc SC_FREE passed", 22: "sc issued", 30: "acc E_READY passed", 31: "acc issued",
          40: "ld item start", 42: "ld rows ready"}


def show(rec, names, keep, item, tiles_per_item, ev_start=1):
    t0 = rec[0][0]
    print("%d records, span %d cycles" % (len(rec), rec[-1][0] - t0))
    return t0


def bwd_main(views, d, h, heads, shifted, item):
    from vitta_b200 import _lib, ops_swin
    from vitta_b200._lib import call, ptr
    dev = torch.device("cuda:0")
    window, shift = (8, 7, 7), ((4, 3, 3) if shifted else (0, 0, 0))
    c = heads * 32
    rows = views * d * h * h
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(rows, 3 * c, generator=g) * 1.2).to(dev)
    table = (torch.randn((2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1), heads, generator=g) * 0.5).to(dev)
    go = torch.randn(rows, c, generator=g).to(dev)
    dims = (views, d, h, h)
    out, lse = ops_swin.wmsa3d_fwd(qkv, table, dims, heads, window, shift, 32 ** -0.5)
    cap = 1 << 13
    for rep in range(2):
        trace = torch.zeros(2 * 16 * cap, dtype=torch.int64, device=dev)
        call("vitta_wmsa3d_bwd_set_trace", ptr(trace), cap)
        ops_swin.wmsa3d_bwd(qkv, table, out, go, lse, dims, heads, window, shift, 32 ** -0.5, 0)
        torch.cuda.synchronize()
        call("vitta_wmsa3d_bwd_set_trace", None, 0)
    t = trace.cpu().numpy().astype("uint64")
    for launch in range(2):
        part = t[launch * 16 * cap:(launch + 1) * 16 * cap]
        rec = sorted(((int(v) >> 16, (int(v) >> 8) & 0xff, int(v) & 0xff) for v in part if v))
        print("==== launch %d (%s): %d records, span %d cycles" % (launch, "query-outer" if launch == 0 else "key-outer",
                                                               len(rec), rec[-1][0] - rec[0][0]))
        # chunk starts of row warp 0: event 1
        st = [r[0] for r in rec if r[1] == 0 and r[2] == 1]
        per_tile = 13
        tiles = [st[i] for i in range(0, len(st), per_tile)]
        print("tile starts (row warp 0), cycles between:", [tiles[i + 1] - tiles[i] for i in range(min(len(tiles) - 1, 12))])
        lo = tiles[item * 4 + 1]
        hi = tiles[item * 4 + 2] + 4000
        for clk, w, e in rec:
            if lo - 2000 <= clk <= hi and w in (0, 4, 8, 12, 13, 14, 15):
                print("%8d  w%-2d %s" % (clk - lo, w, BNAMES.get(e, str(e))))


def main():
    bwd = "--bwd" in sys.argv
    a = [int(v) for v in sys.argv[1:] if v != "--bwd"]
    if bwd:
        views, d, h, heads, shifted, item = (a + [16, 16, 14, 12, 1, 2][len(a):])[:6]
        return bwd_main(views, d, h, heads, shifted, item)
    views, d, h, heads, shifted, item = (a + [16, 16, 14, 12, 1, 2][len(a):])[:6]
    from vitta_b200 import _lib
    from vitta_b200._lib import call, ptr, stream_ptr
    dev = torch.device("cuda:0")
    window, shift = (8, 7, 7), ((4, 3, 3) if shifted else (0, 0, 0))
    c = heads * 32
    rows = views * d * h * h
    g = torch.Generator().manual_seed(0)
    qkv = (torch.randn(rows, 3 * c, generator=g) * 1.2).to(dev)
    table = (torch.randn((2 * window[0] - 1) * (2 * window[1] - 1) * (2 * window[2] - 1), heads, generator=g) * 0.5).to(dev)
    qam = qkv.abs().max().reshape(1).contiguous()
    out = torch.empty(rows, c, device=dev)
    nwin = rows // 392
    lse = torch.empty(nwin * heads * 392, device=dev)
    cap = 1 << 13
    i3 = lambda v: (C.c_int * 3)(*v)
    for rep in range(2):
        trace = torch.zeros(15 * cap, dtype=torch.int64, device=dev)
        call("vitta_wmsa3d_fwd_trace", ptr(qkv), ptr(qam), ptr(table), ptr(out), ptr(lse), views, d, h, h, heads, 32, i3(window),
             i3(shift), float(32 ** -0.5), ptr(trace), cap, stream_ptr())
        torch.cuda.synchronize()
    t = trace.cpu().numpy().astype("uint64")
    rec = sorted(((int(v) >> 16, (int(v) >> 8) & 0xff, int(v) & 0xff) for v in t if v))
    t0 = rec[0][0]
    print("%d records, span %d cycles" % (len(rec), rec[-1][0] - t0))
    # tiles of softmax warp 0: event 1 marks a tile start
    starts = [r[0] for r in rec if r[1] == 0 and r[2] == 1]
    print("tile starts (warp 0), cycles between:", [starts[i + 1] - starts[i] for i in range(min(len(starts) - 1, 16))])
    lo, hi = starts[item * 4], starts[item * 4 + 5] if len(starts) > item * 4 + 5 else rec[-1][0]
    keep = {0, 4, 8, 12, 13, 14}
    for clk, w, e in rec:
        if lo - 3000 <= clk <= hi and w in keep:
            print("%8d  w%-2d %s" % (clk - lo, w, NAMES.get(e, str(e))))


if __name__ == "__main__":
    main()
