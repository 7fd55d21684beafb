"""Times the pieces of the preconditioner build (K_MM is random SPD here): potrf, T T^T, explicit inverses."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
from odf import ops  # noqa: E402

def ev():
    return torch.cuda.Event(enable_timing=True)

for M in [int(a) for a in sys.argv[1:]] or [10000]:
    g = torch.Generator(device="cuda").manual_seed(0)
    X = torch.randn(M, 256, device="cuda", generator=g)
    X *= 20.0 / X.norm(dim=1).mean()
    pc = ops.Prepared(X)
    K = ops.kmm(pc, 15.0)
    for rep in range(2):
        Kc = K.clone()
        torch.cuda.synchronize()
        e = [ev() for _ in range(4)]
        e[0].record()
        Tm, Am = ops.precond_init(Kc, 1e-5, 1e-5)
        e[1].record()
        Ti = ops.precond_invert(Tm)
        e[2].record()
        Ai = ops.precond_invert(Am)
        e[3].record()
        torch.cuda.synchronize()
        print(f"M={M} rep={rep}: precond_init {e[0].elapsed_time(e[1]):.1f} ms  invert(T) {e[1].elapsed_time(e[2]):.1f} ms  invert(A) {e[2].elapsed_time(e[3]):.1f} ms", flush=True)
    I = torch.eye(M, device="cuda")
    r1 = float((Ti.double()[:2000, :] @ Tm.double()[:, :2000] - I[:2000, :2000].double()).abs().max())
    r2 = float((Ai.double()[-2000:, :] @ Am.double()[:, -2000:] - I[-2000:, -2000:].double()).abs().max())
    A_ref = (Tm.double()[:1500] @ Tm.double()[:1500].T) / M
    A_got = (Am.double().T @ Am.double())[:1500, :1500]
    print(f"   |Tinv T - I| {r1:.2e}  |Ainv A - I| {r2:.2e}  |A^T A - (T T^T/M + lam I)| {float((A_got - A_ref - 1e-5 * I[:1500, :1500].double()).abs().max()):.2e}")
