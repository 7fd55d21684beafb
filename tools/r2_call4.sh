set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
ODF_PRECOND_TRACE=2 timeout 300 python - > gpurun_out/r2d_precond_trace.log 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, "online-detection_b200")
from odf import ops
for M in (2500, 5000, 10000, 10000, 30000):
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(M, 256, device="cuda", generator=g)
    X *= 20.0 / X.norm(dim=1).mean()
    K = ops.kmm(ops.Prepared(X), 15.0)
    torch.cuda.synchronize()
    ops.precond_build_tc(K, 1e-5, 1e-5)
    torch.cuda.synchronize()
    del K
PY
el "trace rc=$?"; cat gpurun_out/r2d_precond_trace.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -15 gpurun_out/r2d_pytest_gpu.log
