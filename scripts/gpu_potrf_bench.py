"""Device time of the single-CTA diagonal-block potrf (elb200_?potrf) and of the panel trsm.
usage: python scripts/gpu_potrf_bench.py"""
import ctypes as C, sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib()
dev = torch.device("cuda:0")
for dt, suf in ((torch.float64, "d"), (torch.complex128, "z"), (torch.float32, "s")):
    for n in (128, 256):
        g = torch.randn(n, n, dtype=dt, device=dev)
        H = g @ g.conj().T + n * torch.eye(n, dtype=dt, device=dev)
        reps = 20
        work = [H.clone().T.contiguous() for _ in range(reps + 2)]   # column-major copies
        info = torch.zeros(1, dtype=torch.int32, device=dev)
        fn = getattr(L, f"elb200_{suf}potrf")
        def run(w):
            check(fn(G.ch("L"), G.i64(n), C.c_void_p(w.data_ptr()), G.i64(n), C.c_void_p(info.data_ptr()), G.stream()), "potrf")
        run(work[0]); run(work[1]); torch.cuda.synchronize()
        clk = (C.c_ulonglong * 4)()
        L.elb200_potrf_phase_clocks(clk, 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            run(work[2 + i])
        e1.record(); torch.cuda.synchronize()
        Lf = torch.tril(work[2].T)           # stored column-major: work is the transpose view
        res = float(torch.linalg.norm(Lf @ Lf.conj().T - H) / torch.linalg.norm(H))
        L.elb200_potrf_phase_clocks(clk, 3)
        print("   clocks per launch: diag %.0f solve %.0f update %.0f (launches %d)" % (clk[0] / max(clk[3], 1), clk[1] / max(clk[3], 1), clk[2] / max(clk[3], 1), clk[3]))
        print(f"potrf {suf} n={n}: {1e3 * e0.elapsed_time(e1) / reps:.1f} us   info={int(info.item())}  ||LL^H-A||/||A||={res:.2e}", flush=True)
# panel trsm of the Cholesky step: X L^H = A21 with A21 (rows x nb)
for rows in (100, 2048, 8192, 32768):
    nb = 256
    g = torch.randn(nb, nb, dtype=torch.float64, device=dev)
    Lm = torch.tril(g) + nb * torch.eye(nb, dtype=torch.float64, device=dev)
    Lc = Lm.T.contiguous()
    B = torch.randn(nb, rows, dtype=torch.float64, device=dev)   # column-major rows x nb
    def run():
        check(L.elb200_dtrsm(G.ch("R"), G.ch("L"), G.ch("T"), G.ch("N"), G.i64(rows), G.i64(nb), C.c_double(1.0),
                             C.c_void_p(Lc.data_ptr()), G.i64(nb), C.c_void_p(B.data_ptr()), G.i64(rows), G.stream()), "trsm")
    B0 = B.clone()
    outs = {}
    for flag in (1, 0):
        L.elb200_trsm_set_debug_flags(flag)
        B.copy_(B0); run(); torch.cuda.synchronize()
        outs[flag] = B.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        # residual of the first solve: X L^T = B0  (B is stored column-major rows x nb -> tensor [nb, rows])
        X = outs[flag].T
        res = float(torch.linalg.norm(X @ Lm.T - B0.T) / (torch.linalg.norm(X) * torch.linalg.norm(Lm)))
        print(f"trsm RLTN {rows}x{nb} {'generic' if flag else 'slab   '}: {1e3 * e0.elapsed_time(e1) / 10:.1f} us  residual {res:.2e}", flush=True)
    print(f"   slab vs generic max diff {float((outs[0] - outs[1]).abs().max()):.2e}")
    L.elb200_trsm_set_debug_flags(0)
