"""Per-kernel top CUDA source lines by warp-stall samples of an .ncu-rep captured with --import-source on from a
-lineinfo build.  Usage: python tools/ncu_lines.py prof.ncu-rep [top_n] [kernel substring]"""
import csv
import io
import subprocess
import sys

from ncu_stalls import kernels


def main(path, top=40, flt=""):
    for name, kid in kernels(path).items():
        if flt not in name:
            continue
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda", "--launch-skip", kid,
                              "--launch-count", "1"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr = next((r for r in rows if "# Samples" in r), None)
        if hdr is None:
            print("##", name[:100], ": no source page")
            continue
        si, src = hdr.index("# Samples"), hdr.index("Source")
        li = hdr.index("#") if "#" in hdr else 0
        stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        data, tot = [], 0
        for r in rows[rows.index(hdr) + 1:]:
            if len(r) <= si or not r[si].isdigit():
                continue
            n = int(r[si])
            tot += n
            data.append((n, r))
        print("## %s  (launch id %s, %d samples)" % (name[:110], kid, tot))
        data.sort(key=lambda x: -x[0])
        for n, r in data[:top]:
            st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall), reverse=True)[:2]
            print("   %6d %5.1f%%  L%-5s %-90s %s" % (n, 100.0 * n / max(tot, 1), r[li], r[src].strip()[:90],
                                                    " ".join("%s=%d" % (b, a) for a, b in st)))


if __name__ == "__main__":
    sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else "")
