"""3xTF32 tcgen05 SGEMM probe: python scripts/gpu_tf32_probe.py check|perf [ta tb]
check: error of elb200_sgemm_3xtf32 vs an FP64 product, beside the exact-FFMA kernel, on several shapes.
perf : device-timed TFLOP/s on the shapes of BASELINE.json configs[4]."""
import ctypes as C, sys
import numpy as np
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib()
dev = torch.device("cuda:0")

def run(fn, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
    check(fn(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), C.c_float(alpha), C.c_void_p(A.data_ptr()), G.i64(lda),
             C.c_void_p(B.data_ptr()), G.i64(ldb), C.c_float(beta), C.c_void_p(Cm.data_ptr()), G.i64(ldc), G.stream()), "sgemm")

def mats(ta, tb, m, n, k, pad=0):
    ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
    lda, ldb, ldc = ar + pad, br + pad, m + pad
    A = torch.empty(ac, lda, dtype=torch.float32, device=dev).uniform_(-1, 1)
    B = torch.empty(bc, ldb, dtype=torch.float32, device=dev).uniform_(-1, 1)
    Cm = torch.empty(n, ldc, dtype=torch.float32, device=dev).uniform_(-1, 1)
    return A, lda, B, ldb, Cm, ldc

def op(X, ld, rows, t):
    M = X[:, :rows].double().T  # rows x cols (column-major storage viewed row-major transposed)
    return M if t == "N" else M.T

mode = sys.argv[1]
if mode == "check":
    combos = [(sys.argv[2], sys.argv[3])] if len(sys.argv) > 3 else [("T", "N"), ("N", "N"), ("N", "T"), ("T", "T")]
    worst = 0.0
    for ta, tb in combos:
        for (m, n, k, pad, alpha, beta) in [(128, 128, 32, 0, 1.0, 0.0), (128, 128, 256, 0, 1.0, 0.0), (256, 384, 96, 0, 3.0, 4.0),
                                          (100, 60, 40, 4, 3.0, 4.0), (1000, 900, 1000, 0, 1.0, 1.0), (2000, 2000, 4096, 0, 1.0, 0.0)]:
            A, lda, B, ldb, Cm, ldc = mats(ta, tb, m, n, k, pad)
            C0 = Cm.clone()
            ar = m if ta == "N" else k; br = k if tb == "N" else n
            ref = alpha * (op(A, lda, ar, ta) @ op(B, ldb, br, tb)) + beta * C0[:, :m].double().T
            run(L.elb200_sgemm_3xtf32, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc)
            torch.cuda.synchronize()
            got = Cm[:, :m].double().T
            C1 = C0.clone()
            run(L.elb200_sgemm, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C1, ldc)
            torch.cuda.synchronize()
            ex = C1[:, :m].double().T
            den = abs(alpha) * torch.linalg.norm(op(A, lda, ar, ta)) * torch.linalg.norm(op(B, ldb, br, tb))
            e3 = float(torch.linalg.norm(got - ref) / den); e1 = float(torch.linalg.norm(ex - ref) / den)
            padok = bool(torch.equal(Cm[:, m:], C0[:, m:]))
            worst = max(worst, e3)
            print(f"{ta}{tb} m={m} n={n} k={k} pad={pad}: 3xTF32 err {e3:.3e} (x 2^-24 = {e3 / 2**-24:.2f})  FFMA err {e1:.3e}  padding untouched {padok}", flush=True)
    print("worst", worst, "OK" if worst < 2e-6 else "FAIL")
else:
    shapes = [(8192, 8192, 32768), (2000, 2000, 32768), (2000, 2000, 262144), (8192, 8192, 128), (16384, 8192, 128), (4096, 4096, 4096)]
    combos = [(sys.argv[2], sys.argv[3])] if len(sys.argv) > 3 else [("T", "N"), ("N", "N")]
    for ta, tb in combos:
        for (m, n, k) in shapes:
            A, lda, B, ldb, Cm, ldc = mats(ta, tb, m, n, k)
            for name, fn in (("3xTF32", L.elb200_sgemm_3xtf32), ("FFMA", L.elb200_sgemm)):
                if name == "FFMA" and 2.0 * m * n * k > 3e13:
                    continue
                run(fn, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 1.0, Cm, ldc)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 3
                e0.record()
                for _ in range(reps):
                    run(fn, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 1.0, Cm, ldc)
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                print(f"{name} {ta}{tb} {m}x{n}x{k}: {ms:.3f} ms  {2.0 * m * n * k / ms / 1e9:.1f} TFLOP/s", flush=True)
            del A, B, Cm
