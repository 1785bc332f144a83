#!/bin/bash
# entry-script comparison only (1 rank vs N ranks over the same synthetic stream)
mkdir -p gpurun_out
TAG=${1:-r2multi}
N=${2:-2}
export VITTA_SYNTHETIC=1 VITTA_N_CORRUPTIONS=1
VITTA_RESULT_DIR=gpurun_out/${TAG}_entry_1gpu timeout 600 python tta_tanet_ucf101.py --batch_size 4 > gpurun_out/${TAG}_entry_1gpu.log 2>&1; echo "entry 1 gpu rc=$?"
VITTA_RESULT_DIR=gpurun_out/${TAG}_entry_${N}gpu timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tta_tanet_ucf101.py --batch_size 4 > gpurun_out/${TAG}_entry_${N}gpu.log 2>&1; echo "entry $N gpu rc=$?"
grep 'TTA Epoch1' gpurun_out/${TAG}_entry_1gpu.log | sed 's/.*TTA/1gpu TTA/' | cut -c1-110
grep 'TTA Epoch1' gpurun_out/${TAG}_entry_${N}gpu.log | sed 's/.*TTA/Ngpu TTA/' | cut -c1-110
