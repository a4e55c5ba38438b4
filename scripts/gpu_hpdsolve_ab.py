"""ZHPDSolve (configs[3]) with the Trsm block factor 1 (the reference's loop) and 4 (default), same process.
usage: python scripts/gpu_hpdsolve_ab.py [n] [rhs]   (under torchrun for more than one GPU)"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
rhs = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from elemental_b200 import api as El

El.Initialize()
g = El.Grid({1: 1, 2: 1, 4: 2, 8: 2}.get(world, 0)) if world > 1 else El.Grid()
El.SetBlocksize(128)
dt = np.complex128
A = El.DistMatrix(dt, El.MC, El.MR, g, n, n).HashFill(1, 7, float(n))
B0 = El.DistMatrix(dt, El.MC, El.MR, g, n, rhs).HashFill(0, 8)
X = El.DistMatrix(dt, El.MC, El.MR, g, n, rhs)
flops = 4.0 * (n ** 3 / 3.0 + 2.0 * n * n * rhs)
for factor in ("1", "4", "8", "1", "4"):
    os.environ["ELB200_TRSM_BLOCK_FACTOR"] = factor
    El.Copy(B0, X)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    El.HPDSolve(El.LOWER, El.NORMAL, A, X)
    t1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([t0.elapsed_time(t1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    R = El.DistMatrix(dt, El.MC, El.MR, g)
    El.Copy(B0, R)
    # A is Hermitian-filled on both triangles: residual with a plain Gemm
    El.Gemm(El.NORMAL, El.NORMAL, -1.0, A, X, 1.0, R)
    res = El.FrobeniusNorm(R) / (n * np.finfo(np.float64).eps * El.FrobeniusNorm(A) * El.FrobeniusNorm(X))
    if g.Rank() == 0:
        print(f"factor {factor}: {ms.item():.1f} ms  {flops / ms.item() / 1e6:.0f} GFLOP/s  residual {res:.3e}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
