#!/bin/bash
# Round 2, call 2: suite under the new default (f16x3), the rewritten bench (both splits), launch list of the f16 step.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/r2c2_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c2_pytest_gpu.log
tail -15 gpurun_out/r2c2_pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c2_bench.log 2> gpurun_out/r2c2_bench.err; echo "bench rc=$?"; head -c 700 gpurun_out/r2c2_bench.log; echo; tail -3 gpurun_out/r2c2_bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-precision tf32x3 > gpurun_out/r2c2_bench_tf32.log 2> gpurun_out/r2c2_bench_tf32.err; echo "bench tf32 rc=$?"; head -c 400 gpurun_out/r2c2_bench_tf32.log; echo
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c2_bench_reference.log 2>&1; echo "bench ref rc=$?"; head -c 300 gpurun_out/r2c2_bench_reference.log; echo
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2c2_launches.csv python bench.py --ncu-step > gpurun_out/r2c2_ncu_step.log 2>&1; echo "ncu list rc=$?"
python __graft_entry__.py --smoke > gpurun_out/r2c2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2c2_smoke.log
