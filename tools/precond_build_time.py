import os, sys, torch
sys.path.insert(0, "online-detection_b200")
from odf import ops
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(M, 256, device="cuda", generator=g)
X *= 20.0 / X.norm(dim=1).mean()
K = ops.kmm(ops.Prepared(X), 15.0)
for rep in range(3):
    Kc = K.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.precond_build_tc(Kc, 1e-5, 1e-5)
    e1.record()
    torch.cuda.synchronize()
    print("M=%d rep %d: %.2f ms" % (M, rep, e0.elapsed_time(e1)), flush=True)
