import os, sys, torch
sys.path.insert(0, "online-detection_b200")
from odf import ops
n, M, T = 131072, 10000, 30
ldp = (M + 127)//128*128
P = torch.rand(n, ldp, device="cuda")
W = torch.zeros(n, 32, device="cuda"); W[:, :T] = torch.randn(n, T, device="cuda")
L = ops._lib.load()
S = int(L.odf_panel_splits(n, M))
out = torch.empty(S, M, 32, device="cuda")
ref = (P[:, :M].double().T @ W.double())
for env in (None, "1"):
    if env: os.environ["ODF_PANEL_SCALAR"] = env
    ops.panel_tmm(P, W, n, M, out); torch.cuda.synchronize()
    err = float((out.double().sum(0) - ref).abs().max() / ref.abs().max())
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.panel_tmm(P, W, n, M, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/10
    print(f"panel_tmm scalar={env} S={S}: {ms:.3f} ms  {n*ldp*4/ms/1e6:.0f} GB/s  err={err:.2e}")
