"""Print the roofline-relevant metrics of every launch in an .ncu-rep (needs ncu on PATH; no GPU).
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_xxx_ncu.md"""
import csv
import io
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pct"),
        ("sm__inst_executed_pipe_uniform.sum", None),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("lts__t_bytes.sum", "l2_bytes"), ("l1tex__data_bank_conflicts_pipe_lsu.sum", "bank_conf")]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units = rows[0], rows[1]
    cols = [(head.index(m), lab, units[head.index(m)]) for m, lab in WANT if lab and m in head]
    ki = head.index("Kernel Name")
    print("| kernel | " + " | ".join("%s [%s]" % (lab, u) for _, lab, u in cols) + " |")
    print("|---|" + "---:|" * len(cols))
    for r in rows[2:]:
        name = r[ki].split("(")[0][:60]
        print("| `%s` | " % name + " | ".join(r[i] for i, _, _ in cols) + " |")


if __name__ == "__main__":
    main(sys.argv[1])
