"""Exact-FFMA float GEMM: register-tiled kernel (gemm_f32_ffma.cu) against the generic SIMT kernel, all four
orientations, with a correctness check against an FP64 product.  usage: python scripts/gpu_sgemm_ffma.py"""
import ctypes as C, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib(); dev = torch.device("cuda:0")
L.elb200_sgemm_set_mode(0)
def run(m, n, k, ta, tb, path, chk=False):
    L.elb200_sgemm_set_ffma_path(path)
    ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
    A = torch.empty(ac, ar, dtype=torch.float32, device=dev).uniform_(-1, 1)
    B = torch.empty(bc, br, dtype=torch.float32, device=dev).uniform_(-1, 1)
    Cm = torch.empty(n, m, dtype=torch.float32, device=dev).uniform_(-1, 1)
    C0 = Cm.clone()
    fn = lambda: check(L.elb200_sgemm(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), C.c_float(1.5), C.c_void_p(A.data_ptr()), G.i64(ar),
                                      C.c_void_p(B.data_ptr()), G.i64(br), C.c_float(0.5), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream()))
    fn(); torch.cuda.synchronize()
    err = None
    if chk:
        opA = A.double().T if ta == "N" else A.double()
        opB = B.double().T if tb == "N" else B.double()
        ref = 1.5 * (opA @ opB) + 0.5 * C0.double().T
        err = float((Cm.double().T - ref).norm() / (k * 2.0 ** -23 * opA.norm() * opB.norm()))
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    kern = L.elb200_sgemm_ffma_last_kernel()
    L.elb200_sgemm_set_ffma_path(0)
    return 2.0 * m * n * k / best / 1e9, err, kern
for (m, n, k) in [(4096, 4096, 4096), (8192, 8192, 8192), (8192, 8192, 32768), (1000, 900, 1100), (16384, 8192, 128)]:
    for ta, tb in (("N", "N"), ("N", "T"), ("T", "N"), ("T", "T")):
        chk = m * n * k <= 4096 ** 3
        new = run(m, n, k, ta, tb, 0, chk)
        old = run(m, n, k, ta, tb, 1, False) if (ta, tb) == ("N", "N") else (0, None, 1)
        print(f"{ta}{tb} {m}x{n}x{k}: register-tiled {new[0]:.2f} TF/s (kernel {new[2]}, err {new[1]})   generic {old[0]:.2f} TF/s", flush=True)
