#!/bin/bash
# 2-GPU validation: bench.py --gpus 2 (bucketed + graph-captured collectives, sharded Swin-B record, parity check with a
# ragged step / idle rank), then the drop-in entry script on 1 and on 2 ranks over the same synthetic stream.
mkdir -p gpurun_out
TAG=${1:-r2multi}
N=${2:-2}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.log 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench $N gpu rc=$?"
python tools/show_bench.py gpurun_out/${TAG}_bench_${N}gpu.log | head -12
grep -v "Warning\|run_backward\|warn" gpurun_out/${TAG}_bench_${N}gpu.err | tail -5
export VITTA_SYNTHETIC=1 VITTA_N_CORRUPTIONS=1
VITTA_RESULT_DIR=gpurun_out/${TAG}_entry_1gpu timeout 600 python tta_tanet_ucf101.py --batch_size 4 > gpurun_out/${TAG}_entry_1gpu.log 2>&1; echo "entry 1 gpu rc=$?"
VITTA_RESULT_DIR=gpurun_out/${TAG}_entry_${N}gpu timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tta_tanet_ucf101.py --batch_size 4 > gpurun_out/${TAG}_entry_${N}gpu.log 2>&1; echo "entry $N gpu rc=$?"
tail -2 gpurun_out/${TAG}_entry_1gpu/*_all_result gpurun_out/${TAG}_entry_${N}gpu/*_all_result
tail -3 gpurun_out/${TAG}_entry_${N}gpu.log
