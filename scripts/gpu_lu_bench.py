"""LU with partial pivoting (El.LU(A, P)) timed on the device, with the solve check of the reference's own driver
(tests/lapack_like/LU.cpp: ||A X - B|| after lu::SolveAfter).  usage: python scripts/gpu_lu_bench.py [n] [nb] [reps]
(under torchrun for more than one GPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 128
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from elemental_b200 import api as El

El.Initialize()
g = El.Grid({1: 1, 2: 1, 4: 2, 8: 2}.get(world, 0)) if world > 1 else El.Grid()
rank = g.Rank()
A = El.DistMatrix(np.float64, El.MC, El.MR, g)
A.Resize(n, n)
F = El.DistMatrix(np.float64, El.MC, El.MR, g)
El.PushBlocksizeStack(nb)
best = None
for it in range(reps + 1):
    A.HashFill(0, 77)
    El.Copy(A, F)
    P = El.DistPermutation(g)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    El.LU(F, P)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if it > 0:
        best = ms.item() if best is None else min(best, ms.item())
# solve check with 16 right-hand sides
B = El.DistMatrix(np.float64, El.MC, El.MR, g)
B.Resize(n, 16)
B.HashFill(0, 78)
X = El.DistMatrix(np.float64, El.MC, El.MR, g)
El.Copy(B, X)
El.LUSolveAfter(El.NORMAL, F, X, P)
R = El.DistMatrix(np.float64, El.MC, El.MR, g)
El.Copy(B, R)
El.Gemm(El.NORMAL, El.NORMAL, -1.0, A, X, 1.0, R)
res = El.FrobeniusNorm(R) / (n * np.finfo(np.float64).eps * El.FrobeniusNorm(A) * El.FrobeniusNorm(X))
El.PopBlocksizeStack()
if rank == 0:
    print(f"LU partial pivoting double n={n} nb={nb} on {world} GPU(s): {best:.1f} ms  {2.0 / 3.0 * n ** 3 / best / 1e6:.0f} GFLOP/s  "
          f"solve residual ||A X - B||_F / (n eps ||A||_F ||X||_F) = {res:.3e}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
