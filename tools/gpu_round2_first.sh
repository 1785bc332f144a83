#!/bin/bash
# First gpurun call of round 2: validate what round 1 prepared without a GPU, then the usual round.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_round2_first.sh'
mkdir -p gpurun_out
# 1. fp16-split GEMM / conv / dgrad kernels against float64 (and against the tf32 kernels)
VITTA_TEST_F16X3=1 timeout 900 python -m pytest tests/test_gpu_gemm_f16.py -q -x > gpurun_out/f16_tests.log 2>&1; echo "f16 tests rc=$?"; tail -5 gpurun_out/f16_tests.log
# 2. option-mode goldens and the live-target BNS test that round 1 could only pin on the CPU oracle
VITTA_TEST_UNVERIFIED=1 timeout 900 python -m pytest tests/test_gpu_tanet.py tests/test_gpu_kernels.py tests/test_gpu_swin.py tests/test_crops.py tests/test_swin_loader.py -q -k "swin_views_to_device or option_modes or live_running or crop_resize_vs_oracle or scale_center_crop_vs_oracle or reference_loader_golden" > gpurun_out/unverified_tests.log 2>&1; echo "unverified rc=$?"; tail -5 gpurun_out/unverified_tests.log
# 3. whole-model parity with the fp16 split routed in (forward + data gradient; weight gradient stays tf32)
VITTA_GEMM_PRECISION=f16x3 timeout 1200 python -m pytest tests/test_gpu_tanet.py tests/test_gpu_swin.py -q > gpurun_out/f16_models.log 2>&1; echo "f16 models rc=$?"; tail -5 gpurun_out/f16_models.log
# 3b. CTA pairs on for the whole model (tf32 split), only if their unit tests passed above
VITTA_GEMM_CTA_PAIR=1 timeout 900 python -m pytest tests/test_gpu_tanet.py -q -k "golden and not option" > gpurun_out/pair_models.log 2>&1; echo "pair models rc=$?"; tail -3 gpurun_out/pair_models.log
# 4. per-shape timing, both splits
timeout 600 python tools/conv_shapes.py --reps 4 > gpurun_out/conv_shapes_tf32.md 2>&1; tail -5 gpurun_out/conv_shapes_tf32.md
VITTA_GEMM_PRECISION=f16x3 timeout 600 python tools/conv_shapes.py --reps 4 > gpurun_out/conv_shapes_f16.md 2>&1; tail -5 gpurun_out/conv_shapes_f16.md
VITTA_GEMM_CTA_PAIR=1 timeout 600 python tools/conv_shapes.py --reps 4 > gpurun_out/conv_shapes_tf32_pair.md 2>&1; tail -5 gpurun_out/conv_shapes_tf32_pair.md
VITTA_GEMM_CTA_PAIR=1 VITTA_GEMM_PRECISION=f16x3 timeout 600 python tools/conv_shapes.py --reps 4 > gpurun_out/conv_shapes_f16_pair.md 2>&1; tail -5 gpurun_out/conv_shapes_f16_pair.md
# 4b. the step itself with the opt-in paths (only meaningful if the tests above passed)
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-precision f16x3 > gpurun_out/bench_f16.log 2> gpurun_out/bench_f16.err; echo "bench f16 rc=$?"; head -c 300 gpurun_out/bench_f16.log; echo
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --gemm-precision f16x3 --cta-pair > gpurun_out/bench_f16_pair.log 2> gpurun_out/bench_f16_pair.err; echo "bench f16+pair rc=$?"; head -c 300 gpurun_out/bench_f16_pair.log; echo
# 4c. W-MSA backward: what the row warps wait on (round 1's last measurement rejected the issue-bound hypothesis,
#     profiles/r01_late_checks.md) -- full capture with source-level warp-state samples of the attention kernels alone
timeout 120 python tools/one_wmsa.py > gpurun_out/one_wmsa.log 2>&1; cat gpurun_out/one_wmsa.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"wmsa3d_fwd|wmsa3d_bwd2" --launch-skip 6 -c 3 -o gpurun_out/prof_wmsa python tools/one_wmsa.py > gpurun_out/ncu_wmsa.log 2>&1; echo "ncu wmsa rc=$?"
# 4d. the reference step on torch's own CUDA kernels (the 'reference-on-GPU' bar), TF32 off and on
timeout 600 python tools/reference_gpu_step.py > gpurun_out/reference_gpu_step.json 2> gpurun_out/reference_gpu_step.err; echo "reference-on-GPU rc=$?"; head -c 600 gpurun_out/reference_gpu_step.json; echo
# 5. the regular round (tests, bench, Swin tables)
bash tools/gpu_round.sh
