set -u
mkdir -p gpurun_out
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
cat > /tmp/apply_time.py <<'PY'
import sys, os, torch
sys.path.insert(0, "online-detection_b200")
from odf import ops
for M, T in ((10000, 30), (10000, 32), (4000, 30), (2000, 1), (30000, 21)):
    g = torch.Generator(device="cuda").manual_seed(0)
    U = torch.randn(M, M, device="cuda", generator=g).triu()
    B = torch.randn(M, T, device="cuda", generator=g)
    out = torch.empty_like(B)
    ref = U.double() @ B.double()
    reft = U.double().T @ B.double()
    for tr, r in ((False, ref), (True, reft)):
        for _ in range(3):
            ops.precond_apply(U, B, out, tr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.precond_apply(U, B, out, tr)
        e1.record()
        torch.cuda.synchronize()
        err = float((out.double() - r).abs().max() / r.abs().max())
        print("%s M=%d T=%d transposed=%d: %.3f ms  err %.2e" % (os.environ.get("ODF_PRECOND_APPLY", "own"), M, T, tr, e0.elapsed_time(e1) / 20, err), flush=True)
    del U, ref, reft
PY
timeout 200 python /tmp/apply_time.py 2>&1 | tail -12
ODF_PRECOND_APPLY=cublas timeout 200 python /tmp/apply_time.py 2>&1 | tail -12
el "apply timing done"
timeout 300 python tools/small_fit_probe.py > gpurun_out/r2j_small_fit_probe.log 2>&1
el "small fit probe rc=$?"; cat gpurun_out/r2j_small_fit_probe.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest_gpu.log 2>&1
el "gpu suite rc=$?"; tail -6 gpurun_out/r2j_pytest_gpu.log
timeout 400 python bench.py --steps 3 --warmup 3 --no-parity --no-c1-pair --no-cpu-baseline --no-streaming-compare > gpurun_out/r2j_bench_c2.json 2> gpurun_out/r2j_bench_c2.err
el "bench rc=$?"; python -c "
import json; j=json.load(open('gpurun_out/r2j_bench_c2.json'))
print({k: j[k] for k in ('ms_per_step','phases_ms','gpu_launches')}, j['e2e']['ms_per_step'])"
