"""Per-kernel top stall sites of an .ncu-rep (SASS level, with the CUDA source line when -lineinfo / --import-source on were
used): for every distinct kernel name in the report, the instructions with the most warp-stall samples and their two
dominant stall reasons.  Usage: python tools/ncu_stalls.py prof.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys


def kernels(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    ki, ii = rows[0].index("Kernel Name"), rows[0].index("ID")
    seen = {}
    for r in rows[2:]:
        seen.setdefault(r[ki], r[ii])
    return seen


def main(path, top=28):
    for name, kid in kernels(path).items():
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", kid,
                              "--launch-count", "1"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = next((r for r in rows if "# Samples" in r), None)
        if hdr is None:
            print("##", name[:100], ": no source page")
            continue
        si, src = hdr.index("# Samples"), hdr.index("Source")
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data, tot = [], 0
        for r in rows[rows.index(hdr) + 1:]:
            if len(r) <= si or not r[si].isdigit():
                continue
            n = int(r[si])
            tot += n
            data.append((n, r))
        print("## %s  (launch id %s, %d samples)" % (name[:110], kid, tot))
        reasons = {}
        for n, r in data:
            for i in stall:
                reasons[hdr[i]] = reasons.get(hdr[i], 0) + int(r[i] or 0)
        print("   stall totals:", ", ".join("%s %.0f%%" % (k[6:], 100.0 * v / max(tot, 1))
                                            for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:6]))
        data.sort(key=lambda x: -x[0])
        for n, r in data[:top]:
            st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
            print("   %6d  %-72s %s" % (n, r[src].strip()[:72], " ".join("%s=%d" % (b, a) for a, b in st)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 28)
