#!/bin/bash
# ncu --set full captures of the step's top kernels (one warm eager step of bench.py --ncu-step), with source correlation
mkdir -p gpurun_out
TAG=${1:-r2prof}
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"bn_act_fwd_kernel" --launch-skip 3 -c 8 -o gpurun_out/${TAG}_bnfwd python bench.py --ncu-step > gpurun_out/${TAG}_bnfwd.log 2>&1; echo "ncu bnfwd rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tf32x3_kernel" --launch-skip 1 -c 30 -o gpurun_out/${TAG}_gemm python bench.py --ncu-step > gpurun_out/${TAG}_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"wgrad_tf32x3_kernel" -c 12 -o gpurun_out/${TAG}_wgrad python bench.py --ncu-step > gpurun_out/${TAG}_wgrad.log 2>&1; echo "ncu wgrad rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --ncu-step > gpurun_out/${TAG}_ncu_step.log 2>&1; echo "ncu list rc=$?"
ls -la gpurun_out/${TAG}_*
