#!/bin/bash
# N-GPU bench line only (what the driver's scaling run launches)
mkdir -p gpurun_out
TAG=${1:-r2scale}
N=${2:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_${N}gpu.log 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench $N gpu rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench_${N}gpu.log 2>/dev/null | head -9
grep -v "Warning\|run_backward\|warn\|OMP_NUM\|\*\*\*\*" gpurun_out/${TAG}_bench_${N}gpu.err | tail -5
