"""Short profiling target for ncu (not a test): a few launches of the fused Gaussian tile at a
BASELINE-config-2-shaped slice (rows 131072 x centres 10000 x d 1024, T=30); ODF_MODE=resident runs the
resident-panel sweeps (transposed + forward tile with spill once, then two panel16 passes per sweep)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
from odf import ops  # noqa: E402

n, M, d, T = int(os.environ.get("ODF_N", 131072)), 10000, int(os.environ.get("ODF_D", 1024)), 30
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(n, d, device="cuda", generator=g)
X *= 20.0 / X[:1024].norm(dim=1).mean()
C = X[torch.randperm(n, device="cuda", generator=g)[:M]].contiguous()
px, pc = ops.Prepared(X), ops.Prepared(C)
v = torch.randn(M, T, device="cuda", generator=g)
mode = os.environ.get("ODF_MODE", "panel16")
sw = ops.Sweeper(px, pc, 20.0, T, mode=mode)
out = torch.empty(M, T, device="cuda")
if mode == "resident":
    # launches: transposed tile (spill) | forward tile (spill), panel16 (K^T w) | then per rep panel16 (K v), panel16 (K^T w)
    sw.dmmv(None, torch.randn(n, T, device="cuda", generator=g), out, 1.0, 1.0 / n)
for _ in range(int(os.environ.get("ODF_REPS", 3))):
    sw.dmmv(v, None, out)
torch.cuda.synchronize()
print("done", float(out.abs().sum()))
