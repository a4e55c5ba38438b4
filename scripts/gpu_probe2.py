"""Second GPU probe: DGEMM tile configs (1 = 128x128/1 CTA per SM, 2 = 128x64/2 CTAs per SM) across the
hot shapes; ZGEMM DMMA kernel speed."""
import ctypes as C, json, sys, os
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib()
dev = torch.device("cuda:0")
out = {}
def time_fn(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
def dgemm(ta, tb, m, n, k, A, lda, B, ldb, Cm, ldc, alpha=1.0, beta=1.0):
    check(L.elb200_dgemm(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), C.c_double(alpha), C.c_void_p(A.data_ptr()), G.i64(lda),
                         C.c_void_p(B.data_ptr()), G.i64(ldb), C.c_double(beta), C.c_void_p(Cm.data_ptr()), G.i64(ldc), G.stream()))
for (m, n, k) in [(8192, 8192, 8192), (16384, 16384, 128), (32768, 32768, 128), (16384, 8192, 128), (16384, 16384, 256), (16384, 8192, 256), (4096, 4096, 256), (2048, 2048, 256), (16384, 256, 8192)]:
    for ta, tb in (("N", "N"), ("T", "N"), ("N", "T")):
        ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
        A = torch.empty(ac, ar, dtype=torch.float64, device=dev).uniform_(-1, 1)
        B = torch.empty(bc, br, dtype=torch.float64, device=dev).uniform_(-1, 1)
        Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
        res = {}
        for cfg in (1, 2):
            L.elb200_dgemm_set_config(cfg)
            t = time_fn(lambda: dgemm(ta, tb, m, n, k, A, ar, B, br, Cm, m), reps=3)
            res[cfg] = 2 * m * n * k / t / 1e9
        print(f"dgemm {ta}{tb} {m}x{n}x{k}: cfg1 {res[1]:.2f}  cfg2 {res[2]:.2f} TFLOP/s")
        out[f"dgemm_{ta}{tb}_{m}_{n}_{k}"] = res
        del A, B, Cm
L.elb200_dgemm_set_config(0)
# zgemm
def zgemm(ta, tb, m, n, k, A, lda, B, ldb, Cm, ldc):
    from elemental_b200._lib import c64
    check(L.elb200_zgemm(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), c64(1.0, 0.0), C.c_void_p(A.data_ptr()), G.i64(lda),
                         C.c_void_p(B.data_ptr()), G.i64(ldb), c64(1.0, 0.0), C.c_void_p(Cm.data_ptr()), G.i64(ldc), G.stream()))
for (m, n, k) in [(4096, 4096, 4096), (8192, 8192, 128), (8192, 4096, 128), (8192, 256, 4096)]:
    for ta, tb in (("N", "N"), ("C", "N"), ("N", "C")):
        ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
        A = torch.randn(ac, ar, dtype=torch.complex128, device=dev)
        B = torch.randn(bc, br, dtype=torch.complex128, device=dev)
        Cm = torch.zeros(n, m, dtype=torch.complex128, device=dev)
        t = time_fn(lambda: zgemm(ta, tb, m, n, k, A, ar, B, br, Cm, m), reps=3)
        print(f"zgemm {ta}{tb} {m}x{n}x{k}: {8*m*n*k/t/1e9:.2f} TFLOP/s (real flops)")
        out[f"zgemm_{ta}{tb}_{m}_{n}_{k}"] = 8 * m * n * k / t / 1e9
a = torch.randn(4096, 4096, dtype=torch.complex128, device=dev); b = torch.randn(4096, 4096, dtype=torch.complex128, device=dev)
t = time_fn(lambda: torch.matmul(a, b)); print(f"cuBLAS zgemm 4096: {8*4096**3/t/1e9:.2f} TFLOP/s")
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe2.json", "w"), indent=1)
