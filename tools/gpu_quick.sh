#!/bin/bash
# quick iteration call: kernel unit tests + model goldens, then the headline bench without the secondary / baseline legs
mkdir -p gpurun_out
TAG=${1:-quick}
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "${2:-not f16 and not crops and not swin_loader}" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/${TAG}_bench.log 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; python tools/show_bench.py gpurun_out/${TAG}_bench.log
