#!/bin/bash
# One gpurun call: GPU parity tests, bench line, Swin step tables, ncu launch lists, ncu full captures of the top kernels.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
head -c 400 gpurun_out/bench.log; echo
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.log 2> gpurun_out/bench_reference.err; echo "bench ref rc=$?"; head -c 300 gpurun_out/bench_reference.log; echo
timeout 600 python tools/swin_step.py --model tiny > gpurun_out/swin_tiny.json 2> gpurun_out/swin_tiny.err; echo "swin tiny rc=$?"; head -c 600 gpurun_out/swin_tiny.json; echo
timeout 600 python tools/swin_step.py --model base --videos 4 > gpurun_out/swin_base.json 2> gpurun_out/swin_base.err; echo "swin base rc=$?"; head -c 600 gpurun_out/swin_base.json; echo
if [ "$1" == "ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step > gpurun_out/ncu_step.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_swin.csv python tools/swin_step.py --model tiny --ncu-step > gpurun_out/ncu_swin_step.log 2>&1; echo "ncu swin list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tf32x3_kernel|bn_act_fwd|wgrad_tf32x3" -c 12 -o gpurun_out/prof_tanet python bench.py --ncu-step > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"bn_act_bwd|wgrad_tf32x3" --launch-skip 88 -c 12 -o gpurun_out/prof_tanet_bwd python bench.py --ncu-step > gpurun_out/ncu_full_bwd.log 2>&1; echo "ncu full bwd rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"wmsa3d|ln_fwd|ln_bwd" -c 10 -o gpurun_out/prof_swin python tools/swin_step.py --model tiny --videos 2 --ncu-step > gpurun_out/ncu_full_swin.log 2>&1; echo "ncu full swin rc=$?"
fi
