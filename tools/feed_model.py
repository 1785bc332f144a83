"""Analytic model (NOT a measurement) of the conv GEMM kernels: per ResNet-50 layer, the cycles a 128-row M tile needs
from (i) the tensor pipe and (ii) the L2 -> SM operand feed, for the shipped tf32x3 kernels and the opt-in f16x3 / CTA-pair
variants, next to the round-1 measured time of the same layer where profiles/r01_conv_shapes.md has one.

  python tools/feed_model.py [--frames 128] [--feed-bpc 43] [--pipe-eff 0.73]

Inputs of the model:
  * MMA work: 3 products per element pair; kind::tf32 2048 MAC/clk/SM nominal, kind::f16 4096; the pipe is power-limited,
    so `--pipe-eff` scales it (0.73 = measured bf16 burst / nominal, MEASURED_PEAKS.json; 0.63 sustained).
  * Operand feed: bytes TMA pulls from L2 per stage / `--feed-bpc` bytes per clock per SM (B300 guide: ~6300 B/clk chip-wide).
    A tile: fp32 activations 128 x K x 4 B; B tile: BN x K x 8 B (tf32 hi+lo) or x 4 B (fp16 hi+lo); CTA pairs load half of B.
  * Wave quantisation of the persistent grid (148 CTAs, or 74 pairs).
  * HBM floor of the whole launch: input pixels x Cin x 4 B read once + M x N x 4 B written, at the measured copy bandwidth
    (6458 GB/s) -- this is what bounds layer 1 (K = 64, output 4x the input)."""
import argparse
import os
import re
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.conv_shapes import r50_convs  # noqa: E402

SMS = 148
CLK_GHZ = 1.965
HBM_GBS = 6458.4


def pick_bn(n, tiles_m):
    """Mirror of pick_bn in csrc/gemm_tf32.cu."""
    if n <= 64:
        return 64
    if n <= 128 or n % 256 != 0:
        return 128
    eff = lambda t: t / (((t + SMS - 1) // SMS) * SMS)
    return 128 if eff(tiles_m * (n // 128)) * 0.88 > eff(tiles_m * (n // 256)) else 256


def model(m, n, k, bn, f16, pair, feed_bpc, pipe_eff):
    tiles_m = (m + 127) // 128
    tiles_n = (n + bn - 1) // bn
    mac_per_clk = (4096 if f16 else 2048) * pipe_eff
    mma = 3.0 * 128 * bn * k / mac_per_clk                       # cycles per (128 x bn) tile, whole K
    b_bytes = bn * k * (4 if f16 else 8) * (0.5 if pair else 1.0)
    feed = (128 * k * 4 + b_bytes) / feed_bpc
    per_tile = max(mma, feed)
    if pair:
        units, slots = ((tiles_m + 1) // 2) * tiles_n, SMS // 2
    else:
        units, slots = tiles_m * tiles_n, SMS
    waves = (units + slots - 1) // slots
    return per_tile * waves / (CLK_GHZ * 1e3), mma, feed, waves   # microseconds


def measured_fwd():
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r01_conv_shapes.md")
    out = {}
    if os.path.exists(path):
        for line in open(path):
            c = [x.strip() for x in line.split("|")]
            if len(c) > 5 and re.match(r"layer\d", c[1]):
                out[c[1]] = float(c[5].split()[0])               # "fwd smem us (TF/s)" column = shipped default for BN >= 128
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=128)
    ap.add_argument("--feed-bpc", type=float, default=43.0)
    ap.add_argument("--pipe-eff", type=float, default=0.73)
    a = ap.parse_args()
    meas = measured_fwd()
    print("MODEL, not a measurement (tools/feed_model.py; feed %.0f B/clk/SM, pipe efficiency %.2f): forward conv GEMM per "
          "layer, microseconds; 'bound' = what limits a tile in the shipped kernel\n" % (a.feed_bpc, a.pipe_eff))
    print("| conv | M x N x K | BN | measured fwd | tf32x3 model | bound | f16x3 | f16x3 + pairs | tf32x3 + pairs |")
    print("|---|---|---:|---:|---:|---|---:|---:|---:|")
    tot = {"meas": 0.0, "tf32": 0.0, "f16": 0.0, "f16p": 0.0, "tf32p": 0.0}
    for name, cin, cout, kk, s, p, r, mult in r50_convs():
        ro = (r + 2 * p - kk) // s + 1
        m, n, k = a.frames * ro * ro, cout, cin * kk * kk
        bn = pick_bn(n, (m + 127) // 128)
        hbm = (a.frames * r * r * cin * 4.0 + m * n * 4.0) / (HBM_GBS * 1e3)      # microseconds
        t0, mma, feed, _ = model(m, n, k, bn, False, False, a.feed_bpc, a.pipe_eff)
        t1 = model(m, n, k, bn, True, False, a.feed_bpc, a.pipe_eff)[0]
        can_pair = n % 256 == 0
        t2 = model(m, n, k, 256, True, True, a.feed_bpc, a.pipe_eff)[0] if can_pair else t1
        t3 = model(m, n, k, 256, False, True, a.feed_bpc, a.pipe_eff)[0] if can_pair else t0
        t2, t3 = min(t2, t1), min(t3, t0)                        # pairs are a per-layer choice
        hbm_bound = hbm >= t0
        t0, t1, t2, t3 = (max(t, hbm) for t in (t0, t1, t2, t3))
        ms = meas.get(name)
        print("| %s | %d x %d x %d | %d | %s | %.0f | %s | %.0f | %.0f | %.0f |" % (
            name, m, n, k, bn, "%.0f" % ms if ms else "-", t0, "hbm" if hbm_bound else ("pipe" if mma >= feed else "feed"), t1, t2,
            t3))
        for key, v in (("meas", ms or 0.0), ("tf32", t0), ("f16", t1), ("f16p", t2), ("tf32p", t3)):
            tot[key] += v * mult
    print("\nper-step totals over all layers (forward only, with multiplicity), ms: measured %.2f | tf32x3 model %.2f | f16x3 "
          "%.2f | f16x3 + pairs %.2f | tf32x3 + pairs %.2f" % tuple(tot[k] / 1e3 for k in ("meas", "tf32", "f16", "f16p",
                                                                                            "tf32p")))
    print("(the data-gradient pass has the same shapes with Cin and Cout exchanged; rows marked hbm cannot gain from a faster "
          "contraction)")


if __name__ == "__main__":
    main()
