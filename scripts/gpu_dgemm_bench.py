"""DGEMM kernel micro-benchmark: cp.async kernel (cfg 2) vs persistent TMA kernel (cfg 3) on the hot shapes."""
import ctypes as C, json, sys, os
import torch
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
shapes = [(8192, 8192, 8192), (32768, 32768, 128), (16384, 8192, 128), (16384, 16384, 256), (8192, 4096, 256), (4096, 4096, 256), (2048, 2048, 256), (16384, 256, 8192), (2000, 2000, 32768)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    shapes = [(8192, 8192, 8192), (32768, 32768, 128), (16384, 16384, 256)]
for (m, n, k) in shapes:
    for ta, tb in (("N", "N"), ("T", "N"), ("N", "T")):
        ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
        A = torch.empty(ac, ar, dtype=torch.float64, device=dev).uniform_(-1, 1)
        B = torch.empty(bc, br, dtype=torch.float64, device=dev).uniform_(-1, 1)
        Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
        res = {}
        for cfg in (2, 3):
            L.elb200_dgemm_set_config(cfg)
            t = time_fn(lambda: dgemm(ta, tb, m, n, k, A, ar, B, br, Cm, m), reps=3)
            res[cfg] = 2 * m * n * k / t / 1e9
        print(f"dgemm {ta}{tb} {m}x{n}x{k}: cp.async {res[2]:.2f}  tma {res[3]:.2f} TFLOP/s", flush=True)
        out[f"dgemm_{ta}{tb}_{m}_{n}_{k}"] = res
        del A, B, Cm
L.elb200_dgemm_set_config(0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/dgemm_bench.json", "w"), indent=1)
