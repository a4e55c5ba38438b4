"""World-size-2 (and 4) execution of the redistribution / sum-scatter plans over
torch.distributed with the gloo backend on CPU tensors: the same plans the CUDA path feeds
to pack kernel -> ncclSend/ncclRecv (ncclReduceScatter) -> unpack kernel, here executed with
numpy packs and gloo isend/irecv (all_reduce + slice stands in for reduce-scatter, which
gloo lacks).  Expected values: the oracle's definition of each distribution."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import elemental_oracle as O
from planutil import (LEGAL, MC, MR, NAMES, STAR, Layout, contract_plan, flat_local, gather_lattice, local_flat,
                      redist_plan, scatter_lattice)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, r, c, errs):
    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        dist.init_process_group("gloo", rank=rank, world_size=world)
        i, j = rank % r, rank // r  # column-major grid order (src/core/Grid.cpp:165)
        rng = np.random.default_rng(5)
        G = rng.standard_normal((11, 9))
        stride = {0: r, 2: c, 3: r * c, 4: r * c, 5: 1}
        for t, ((u, v), (u2, v2)) in enumerate([(a, b) for a in LEGAL for b in LEGAL]):
            transpose = (t % 3 == 0)
            la = Layout(u, v, t % stride[u], (t // 2) % stride[v])
            lb = Layout(u2, v2, (t // 3) % stride[u2], (t // 5) % stride[v2])
            h, w = (G.shape[1], G.shape[0]) if transpose else G.shape
            mine = O.local_part(G, NAMES[u], NAMES[v], la.colAlign, la.rowAlign, r, c, i, j)
            ldA = max(mine.shape[0], 1)
            fa = local_flat(mine, ldA)
            want = O.local_part(G.T if transpose else G, NAMES[u2], NAMES[v2], lb.colAlign, lb.rowAlign, r, c, i, j)
            ldB = max(want.shape[0], 1)
            fb = np.full(ldB * max(want.shape[1], 1), np.nan)
            plan = redist_plan(r, c, i, j, h, w, la, ldA, lb, ldB, transpose)
            reqs, recvs = [], []
            for m in plan:
                peer = m.peerRow + r * m.peerCol
                if m.kind == 0:
                    buf = torch.from_numpy(np.ascontiguousarray(gather_lattice(fa, m, True)))
                    reqs.append(dist.isend(buf, peer, tag=t))
                elif m.kind == 1:
                    buf = torch.empty((m.nrows, m.ncols), dtype=torch.float64)
                    reqs.append(dist.irecv(buf, peer, tag=t))
                    recvs.append((m, buf))
                else:
                    scatter_lattice(fb, m, gather_lattice(fa, m, True))
            for q in reqs:
                q.wait()
            for m, buf in recvs:
                scatter_lattice(fb, m, buf.numpy())
            got = flat_local(fb, want.shape[0], want.shape[1], ldB)
            if not np.array_equal(got, want):
                raise AssertionError(f"rank {rank}: [{NAMES[u]},{NAMES[v]}]->[{NAMES[u2]},{NAMES[v2]}] t={transpose}")
        # ---- sum-scatter (AxpyContract) ----
        row_group = {a: dist.new_group([a + r * b for b in range(c)]) for a in range(r)}
        col_group = {b: dist.new_group([a + r * b for a in range(r)]) for b in range(c)}
        for (u, v) in [(0, 5), (5, 2), (2, 5), (5, 0), (5, 5)]:
            la = Layout(u, v, 0, 0)
            lb = Layout(MC, MR, 0, 0)
            h, w = 10, 7
            part = np.random.default_rng(100 + rank).standard_normal((h, w))
            mine = O.local_part(part, NAMES[u], NAMES[v], 0, 0, r, c, i, j)
            ldA = max(mine.shape[0], 1)
            kind, T, chunk, packs = contract_plan(r, c, i, j, h, w, la, ldA, lb)
            send = np.zeros(chunk * len(packs))
            fa = local_flat(mine, ldA)
            for q, m in enumerate(packs):
                if m.nrows:
                    scatter_lattice(send[q * chunk:(q + 1) * chunk], m, gather_lattice(fa, m, True))
            group, me = (row_group[i], j) if kind == 0 else ((col_group[j], i) if kind == 1 else (None, rank))
            tsend = torch.from_numpy(send)
            dist.all_reduce(tsend, group=group)
            red = tsend.numpy()[me * chunk:(me + 1) * chunk]
            # expected: sum of the parts of the ranks in my reduce group, my piece of T
            members = ([i + r * b for b in range(c)] if kind == 0 else
                       ([a + r * j for a in range(r)] if kind == 1 else list(range(world))))
            tot = sum(np.random.default_rng(100 + q).standard_normal((h, w)) for q in members)
            want = O.local_part(tot, NAMES[T.colDist], NAMES[T.rowDist], T.colAlign, T.rowAlign, r, c, i, j)
            got = flat_local(red, want.shape[0], want.shape[1], max(want.shape[0], 1))
            np.testing.assert_allclose(got, want, atol=1e-12)
        dist.barrier()
        dist.destroy_process_group()
    except Exception as e:  # pragma: no cover
        import traceback
        errs.put(f"rank {rank}: {e}\n{traceback.format_exc()}")
        raise


@pytest.mark.parametrize("r,c", [(1, 2), (2, 1), (2, 2)])
def test_plans_over_gloo(r, c):
    world = r * c
    ctx = mp.get_context("spawn")
    errs = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(k, world, port, r, c, errs)) for k in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    msgs = []
    while not errs.empty():
        msgs.append(errs.get())
    assert not msgs, "\n".join(msgs)
    assert all(p.exitcode == 0 for p in procs)
