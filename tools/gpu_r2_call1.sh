#!/bin/bash
# Round 2, first gpurun call: the WHOLE GPU suite (no -x, no env gates left), then the measurements round 1 never got.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2c1_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c1_pytest_gpu.log
tail -40 gpurun_out/r2c1_pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2c1_bench.log 2> gpurun_out/r2c1_bench.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2c1_bench.log; echo
timeout 600 python tools/reference_gpu_step.py > gpurun_out/r2c1_reference_gpu_step.json 2> gpurun_out/r2c1_reference_gpu_step.err; echo "reference-on-GPU rc=$?"; head -c 800 gpurun_out/r2c1_reference_gpu_step.json; echo
timeout 400 python tools/conv_shapes.py --reps 4 > gpurun_out/r2c1_conv_shapes_tf32.md 2>&1; tail -4 gpurun_out/r2c1_conv_shapes_tf32.md
VITTA_GEMM_PRECISION=f16x3 timeout 400 python tools/conv_shapes.py --reps 4 > gpurun_out/r2c1_conv_shapes_f16.md 2>&1; tail -4 gpurun_out/r2c1_conv_shapes_f16.md
VITTA_GEMM_CTA_PAIR=1 timeout 400 python tools/conv_shapes.py --reps 4 > gpurun_out/r2c1_conv_shapes_tf32_pair.md 2>&1; tail -4 gpurun_out/r2c1_conv_shapes_tf32_pair.md
VITTA_GEMM_CTA_PAIR=1 VITTA_GEMM_PRECISION=f16x3 timeout 400 python tools/conv_shapes.py --reps 4 > gpurun_out/r2c1_conv_shapes_f16_pair.md 2>&1; tail -4 gpurun_out/r2c1_conv_shapes_f16_pair.md
VITTA_GEMM_PRECISION=f16x3 timeout 900 python -m pytest tests/test_gpu_tanet.py tests/test_gpu_swin.py -q -p no:cacheprovider > gpurun_out/r2c1_f16_models.log 2>&1; echo "f16 models rc=$?"; tail -5 gpurun_out/r2c1_f16_models.log
timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-precision f16x3 > gpurun_out/r2c1_bench_f16.log 2> gpurun_out/r2c1_bench_f16.err; echo "bench f16 rc=$?"; head -c 300 gpurun_out/r2c1_bench_f16.log; echo
timeout 120 python tools/one_wmsa.py > gpurun_out/r2c1_one_wmsa.log 2>&1; tail -5 gpurun_out/r2c1_one_wmsa.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wmsa3d_fwd|wmsa3d_bwd2" --launch-skip 6 -c 3 -o gpurun_out/r2c1_prof_wmsa python tools/one_wmsa.py > gpurun_out/r2c1_ncu_wmsa.log 2>&1; echo "ncu wmsa rc=$?"
timeout 600 python tools/swin_step.py --model tiny > gpurun_out/r2c1_swin_tiny.json 2> gpurun_out/r2c1_swin_tiny.err; echo "swin tiny rc=$?"; head -c 400 gpurun_out/r2c1_swin_tiny.json; echo
