"""DRAM traffic per launch of the kernels in .ncu-rep files (dram__bytes_read.sum + dram__bytes_write.sum, averaged over
the captured launches of each kernel family) -> JSON that bench.py reports as roofline.traffic.
Usage: python tools/ncu_traffic.py out.json key=path.ncu-rep [key=path.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def traffic(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    ir, iw, it = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum"), head.index("gpu__time_duration.sum")
    tot = dur = 0.0
    for r in rows[2:]:
        tot += float(r[ir]) * UNIT[units[ir]] + float(r[iw]) * UNIT[units[iw]]
        dur += float(r[it]) * {"us": 1.0, "ms": 1e3, "ns": 1e-3, "s": 1e6}.get(units[it], 1.0)
    n = len(rows) - 2
    return {"dram_bytes_per_launch": tot / n, "launches_captured": n, "avg_us_under_ncu": dur / n,
            "source": "ncu --set full --clock-control none, bench.py --ncu-step (one warm eager step)"}


if __name__ == "__main__":
    res = {}
    for kv in sys.argv[2:]:
        k, p = kv.split("=", 1)
        try:
            res[k] = traffic(p)
        except Exception as e:      # a missing capture must not hide the others
            res[k] = None
            sys.stderr.write("%s: %s\n" % (k, e))
    json.dump(res, open(sys.argv[1], "w"), indent=1)
    print(json.dumps(res))
