set -u
mkdir -p gpurun_out
cat > /tmp/onefit.py <<'PY'
import sys, torch
sys.path.insert(0, "."); sys.path.insert(0, "online-detection_b200")
import odf
from oracle import falkon_oracle as orc
N, M, d = 6000, 2000, 2048
X, c, Y = orc.make_synthetic(N, d, 1, seed=0, pos_fraction=0.5)
y = Y[:, 0].contiguous().cuda(); Xg = X.cuda()
C = Xg[torch.randperm(N)[:M].cuda()].contiguous()
for _ in range(3):
    m = odf.InCoreFalkon(kernel=odf.GaussianKernel(5.0), penalty=1e-4, M=M)
    m.fit(Xg, y, centres=C)
torch.cuda.synchronize()
print("iters", m.fit_times_)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r3g_small_launches.csv python /tmp/onefit.py > gpurun_out/r3g_small.log 2>&1
tail -2 gpurun_out/r3g_small.log
python tools/ncu_launch_summary.py gpurun_out/r3g_small_launches.csv > gpurun_out/r3g_small_summary.md; head -40 gpurun_out/r3g_small_summary.md
