set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 200 python tools/gemm_split_debug.py > gpurun_out/r2c_gemm_debug.log 2>&1
el "gemm debug rc=$?"; cat gpurun_out/r2c_gemm_debug.log | grep -v "bad 0 of" 
timeout 100 python tools/gemm_split_debug.py crash > gpurun_out/r2c_gemm_crash.log 2>&1
el "crash shape rc=$?"; tail -3 gpurun_out/r2c_gemm_crash.log
timeout 500 python tools/precond_tc_check.py 2500 10000 > gpurun_out/r2c_precond_tc.log 2>&1
el "precond tc rc=$?"; cat gpurun_out/r2c_precond_tc.log | tail -40
ODF_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "split_gemm or overlapped" > gpurun_out/r2c_experimental.log 2>&1
el "experimental rc=$?"; tail -5 gpurun_out/r2c_experimental.log
