"""2+ GPU timing of the distributed preconditioner build (ODF_PRECOND_PROFILE=1 prints the segments on rank 0)."""
import os, sys
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "online-detection_b200"))
import odf  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
M, d, n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000, 256, 20000
g = torch.Generator(device="cuda").manual_seed(0)
X = torch.randn(n, d, device=dev, generator=g)
X *= 20.0 / X.norm(dim=1).mean()
Y = torch.sign(torch.randn(n, 3, device=dev, generator=g))
C = X[:M].contiguous() if M <= n else torch.randn(M, d, device=dev, generator=g) * (20.0 / d ** 0.5)
dist.broadcast(C, src=0)
for rep in range(3):
    m = odf.InCoreFalkon(kernel=odf.GaussianKernel(15.0), penalty=1e-3, M=M, process_group=None, options=odf.FalkonOptions(distributed_precond=True))
    m.fit(X, Y, centres=C)
    if rank == 0:
        print("rep", rep, {k: round(v, 1) for k, v in m.fit_times_.items() if k.endswith("_ms")}, flush=True)
dist.destroy_process_group()
