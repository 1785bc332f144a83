#!/bin/bash
# ncu --set full captures of the step's top kernels (one warm eager step of bench.py --ncu-step), with source correlation.
# The reports are summarised ON the box (gpurun copies back at most 64 MiB): raw metric table + per-line stall samples.
mkdir -p gpurun_out
TAG=${1:-r2prof}
prof() {   # name regex skip count
  timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"$2" --launch-skip $3 -c $4 -o /tmp/${TAG}_$1 python bench.py --ncu-step > gpurun_out/${TAG}_$1.log 2>&1; echo "ncu $1 rc=$?"
  python tools/ncu_summary.py /tmp/${TAG}_$1.ncu-rep > gpurun_out/${TAG}_$1_summary.md 2>&1
  python tools/ncu_stalls.py /tmp/${TAG}_$1.ncu-rep > gpurun_out/${TAG}_$1_stalls.txt 2>&1
  ls -la /tmp/${TAG}_$1.ncu-rep
}
prof bnfwd "bn_act_fwd_kernel" 3 6
prof gemm "gemm_tf32x3_kernel" 4 24
prof wgrad "wgrad_tf32x3_kernel" 0 10
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --ncu-step > gpurun_out/${TAG}_ncu_step.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_traffic.py gpurun_out/${TAG}_traffic.json gemm_tf32x3_kernel=/tmp/${TAG}_gemm.ncu-rep bn_act_fwd_kernel=/tmp/${TAG}_bnfwd.ncu-rep wgrad_tf32x3_kernel=/tmp/${TAG}_wgrad.ncu-rep > /dev/null
du -sh gpurun_out
