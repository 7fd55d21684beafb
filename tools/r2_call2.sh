set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 300 python tools/gemm_split_debug.py > gpurun_out/r2b_gemm_debug.log 2>&1
el "gemm debug rc=$?"; cat gpurun_out/r2b_gemm_debug.log
timeout 400 compute-sanitizer --tool memcheck --print-limit 5 python tools/gemm_split_debug.py crash > gpurun_out/r2b_sanitizer.log 2>&1
el "sanitizer rc=$?"; grep -v "^=========     at\|^=========         by\|^=========     Host" gpurun_out/r2b_sanitizer.log | head -60
timeout 300 python -m pytest tests -m gpu -x -q -k "wrapper or abi or host_resident or kmm" > gpurun_out/r2b_pytest_subset.log 2>&1
el "subset rc=$?"; tail -5 gpurun_out/r2b_pytest_subset.log
