set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
cat > /tmp/trace.py <<'PY'
import sys, torch
sys.path.insert(0, "online-detection_b200")
from odf import ops
for M in [int(a) for a in sys.argv[1:]]:
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(M, 256, device="cuda", generator=g)
    X *= 20.0 / X.norm(dim=1).mean()
    K = ops.kmm(ops.Prepared(X), 15.0)
    for rep in range(2):
        Kc = K.clone()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.precond_build_tc(Kc, 1e-5, 1e-5)
        e1.record()
        torch.cuda.synchronize()
        print("M=%d total %.2f ms" % (M, e0.elapsed_time(e1)), flush=True)
    del K
PY
for NBV in 1024 2048 4096; do
  echo "== ODF_PRECOND_NB=$NBV"
  ODF_PRECOND_NB=$NBV ODF_PRECOND_TRACE=2 timeout 200 python /tmp/trace.py 10000 2>&1 | tail -4
done
echo "== ODF_PRECOND_NB=1024 no copy"
ODF_PRECOND_POTRF_COPY=0 ODF_PRECOND_NB=1024 ODF_PRECOND_TRACE=2 timeout 200 python /tmp/trace.py 10000 2>&1 | tail -3
echo "== default, M=2500 5000 30000"
ODF_PRECOND_TRACE=2 timeout 200 python /tmp/trace.py 2500 5000 30000 2>&1 | grep -v "^M=.*total" | cat
ODF_PRECOND_TRACE=0 timeout 200 python /tmp/trace.py 1000 2000 2500 5000 30000 2>&1 | tail -12
el "traces done"
timeout 500 python tools/precond_tc_check.py 2500 10000 > gpurun_out/r2e_precond_tc.log 2>&1
el "precond tc rc=$?"; cat gpurun_out/r2e_precond_tc.log | tail -30
