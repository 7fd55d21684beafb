set -u
mkdir -p gpurun_out
cat > /tmp/apply_one.py <<'PY'
import sys, torch
sys.path.insert(0, "online-detection_b200")
from odf import ops
M, T = 10000, 30
g = torch.Generator(device="cuda").manual_seed(0)
U = torch.randn(M, M, device="cuda", generator=g).triu()
B = torch.randn(M, T, device="cuda", generator=g)
out = torch.empty_like(B)
for tr in (False, True, False, True):
    ops.precond_apply(U, B, out, tr)
torch.cuda.synchronize()
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tri_apply -s 2 -c 2 -f -o gpurun_out/r2p_tri_apply python /tmp/apply_one.py > gpurun_out/r2p_ncu.log 2>&1
tail -3 gpurun_out/r2p_ncu.log; ls -la gpurun_out/r2p_tri_apply.ncu-rep
