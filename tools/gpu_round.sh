#!/bin/bash
# One gpurun call: GPU parity tests, bench line, per-kernel step table, ncu launch list, ncu full capture of K1.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log
timeout 300 python tools/step_profile.py > gpurun_out/step_profile.txt 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python bench.py --ncu-step > gpurun_out/ncu_step.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:stats -c 6 -o gpurun_out/prof_stats python bench.py --ncu-step > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -5 gpurun_out/pytest_gpu.log
