#!/bin/bash
# Swin iteration call: Swin GPU tests, then the two Swin records of bench.py (through tools/swin_step.py)
mkdir -p gpurun_out
TAG=${1:-swin}
timeout 900 python -m pytest tests/test_gpu_swin.py -q -x -p no:cacheprovider > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python tools/swin_step.py --model tiny > gpurun_out/${TAG}_tiny.json 2> gpurun_out/${TAG}_tiny.err; echo "swin tiny rc=$?"
timeout 600 python tools/swin_step.py --model base --videos 4 > gpurun_out/${TAG}_base.json 2> gpurun_out/${TAG}_base.err; echo "swin base rc=$?"
python - <<PY
import json
for f in ("gpurun_out/${TAG}_tiny.json", "gpurun_out/${TAG}_base.json"):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][0])
    except Exception as e:
        print(f, "ERR", e); continue
    print(d["workload"][:60], "ms %.1f" % d["ms_per_step"], "launches", d["gpu_launches_per_step"])
    for k, v in list(d["kernels"].items())[:12]:
        print("    ", k, v)
PY
