set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
ODF_EXPERIMENTAL=1 timeout 400 python -m pytest tests/test_gpu_parity.py -q -k "split_gemm or blocked or hi_only or overlapped" > gpurun_out/r2a_experimental.log 2>&1
el "experimental rc=$?"; tail -40 gpurun_out/r2a_experimental.log
timeout 300 python tools/precond_probe.py > gpurun_out/r2a_precond_probe.log 2>&1
el "probe rc=$?"; cat gpurun_out/r2a_precond_probe.log
ODF_PRECOND_PROFILE=1 timeout 200 python tools/precond_time.py 10000 > gpurun_out/r2a_precond_time.log 2>&1
el "precond_time rc=$?"; tail -12 gpurun_out/r2a_precond_time.log
