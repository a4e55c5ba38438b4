"""GemmHost (host-resident operands) at configs[1] under torchrun: sweep of the number of C bands / A chunks, with the
peer-memory path on and off, in ONE launch.  usage: torchrun ... scripts/gpu_gemmhost_bands.py [n]"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from elemental_b200 import api as El

El.Initialize()
El.SetBlocksize(128)
H = {1: 1, 2: 1, 4: 2, 8: 2}.get(world, 0)


def run(grid, tag, variants):
    r, c = grid.Height(), grid.Width()
    lh, lw = (n + r - 1) // r, (n + c - 1) // c
    pinned = [torch.empty((lw, lh), dtype=torch.float64, pin_memory=True) for _ in range(3)]
    for i, t in enumerate(pinned):
        t.uniform_(-1, 1, generator=torch.Generator().manual_seed(100 + i + 10 * grid.Rank()))
    for bands, chunks in variants:
        os.environ["ELB200_GEMMHOST_BANDS"] = str(bands)
        os.environ["ELB200_GEMMHOST_CHUNKS"] = str(chunks)
        ts = []
        for it in range(3):
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            El.GemmHost(El.NORMAL, El.NORMAL, 1.0, grid, n, n, n, pinned[0], pinned[1], 1.0, pinned[2], El.GEMM_SUMMA_C)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if it > 0:
                ts.append(dt.item())
        if grid.Rank() == 0:
            best = min(ts)
            print(f"{tag} bands={bands} chunks={chunks}: {best * 1e3:.1f} ms  {2.0 * n ** 3 / best / 1e12:.1f} TF/s  (runs {[round(x * 1e3, 1) for x in ts]})", flush=True)
    del pinned


g1 = El.Grid(H) if world > 1 else El.Grid()
variants = [(8, 8), (4, 8), (2, 8), (4, 4)]
if os.environ.get("BANDS_SWEEP"):   # e.g. BANDS_SWEEP="8x8,16x8,16x16"
    variants = [tuple(int(x) for x in v.split("x")) for v in os.environ["BANDS_SWEEP"].split(",")]
run(g1, "p2p=on ", variants)
if world > 1:
    os.environ["ELB200_P2P"] = "0"
    g2 = El.Grid(H)
    run(g2, "p2p=off", [(8, 8), (4, 8)])
    dist.barrier()
    dist.destroy_process_group()
