"""Print the interesting parts of a bench.py JSON line (helper for reading gpurun logs)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads([ln for ln in open(f) if ln.startswith("{")][0])
    except Exception as e:
        print(f, "ERR", e)
        continue
    print(f, "value %.1f  ms %.2f  e2e %.1f  launches %d  with_eval %.1f" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"],
        d["config"].get("with_eval_fwd_clips_per_s", 0)))
    r, s = d["roofline"], d["roofline_stats"]
    print("  gemm %.1f TF/s frac %.3f (%.2f ms)   bn_act_fwd %.0f GB/s frac %.3f   k1 frac %.3f" % (
        r["achieved"], r["frac"], r["ms_per_step"], s["achieved"], s["frac"], s["k1_standalone"]["frac"]))
    if d.get("gpu_reference"):
        print("  gpu_reference", [(x["tf32"], round(x["ms_per_step"], 1)) for x in d["gpu_reference"]["runs"]])
    if d.get("parity_check"):
        print("  parity", d["parity_check"])
    for sec in d.get("secondary") or []:
        print("  sec %s: %.1f ms, %.1f videos/s" % (sec["workload"][:48], sec["ms_per_step"], sec["videos_per_s"]))
        for k, v in list(sec["kernels"].items())[:9]:
            print("      ", k, v)
    tot = 0.0
    for k, v in d["kernels"].items():
        tot += v["ms"]
        print("   ", k, v)
    print("   sum of our kernels %.2f ms of %.2f" % (tot, d["ms_per_step"]))
    if d.get("kernels_eval_forward"):
        print("   -- evaluation forward:")
        tot = 0.0
        for k, v in d["kernels_eval_forward"].items():
            tot += v["ms"]
            print("    ", k, v)
        print("     sum %.2f ms" % tot)
