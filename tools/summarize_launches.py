"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table (markdown).
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
import csv
import re
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"\(.*", "", name)           # drop the argument list
    name = re.sub(r"^void\s+", "", name)
    return name[:90]


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    head = next(rd)
    ki, ui, vi, gi, bi = (head.index(k) for k in ("Kernel Name", "Metric Unit", "Metric Value", "Grid Size", "Block Size"))
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    n = 0
    for r in rd:
        v = float(r[vi].replace(",", ""))
        unit = r[ui]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += us
        total += us
        n += 1
    ours = sum(t for k, (c, t) in agg.items() if k.startswith("vitta::"))
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, c, t, 100 * t / total))
    print()
    print("%d launches, %.2f ms serialised device time; vitta:: kernels %.2f ms (%.1f%%)" % (n, total / 1e3, ours / 1e3,
                                                                                          100 * ours / total))


if __name__ == "__main__":
    main(sys.argv[1])
