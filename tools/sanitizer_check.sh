# compute-sanitizer memcheck over the kernels touched in round 2 (small shapes): triangular application, panel kernels with the
# converter stage (through the resident sweeper and the standalone kernel tests), tile spill of the 3-byte panel.
set -u
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -m gpu -x -q \
  -k "triangular_apply_matches_fp64 and (31-3 or 1000-30 or 2000-1) or test_panel16_kernel_matches_fp64_product or resident_partial_matches or test_hi_only_panel_kernels" \
  > gpurun_out/${TAG:-r3o}_sanitizer.log 2>&1
echo "sanitizer rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|error" gpurun_out/${TAG:-r3o}_sanitizer.log | head -12
