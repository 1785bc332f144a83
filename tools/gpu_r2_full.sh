#!/bin/bash
# The round's reference run: what the driver runs (full GPU suite with -x, smoke, both bench arms), then the ncu evidence.
mkdir -p gpurun_out
TAG=${1:-r2full}
timeout 1800 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest_gpu.log
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.log 2> gpurun_out/${TAG}_bench_reference.err; echo "bench ref rc=$?"; head -c 300 gpurun_out/${TAG}_bench_reference.log; echo
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench.log | head -60
if [ "$2" == "prof" ]; then bash tools/gpu_r2_prof.sh ${TAG}; fi
