"""Rasterisation band width of the third-generation DGEMM kernel (flags bits 20..25) on the rank-nb update shapes."""
import ctypes as C, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib(); dev = torch.device("cuda:0")
L.elb200_dgemm_set_config(3)
def run(m, n, k, bw):
    A = torch.empty(k, m, dtype=torch.float64, device=dev).uniform_(-1, 1)
    B = torch.empty(n, k, dtype=torch.float64, device=dev).uniform_(-1, 1)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    L.elb200_dgemm_set_debug_flags(bw << 20)
    fn = lambda: check(L.elb200_dgemm(G.ch("N"), G.ch("N"), G.i64(m), G.i64(n), G.i64(k), C.c_double(1.0), C.c_void_p(A.data_ptr()), G.i64(m),
                                      C.c_void_p(B.data_ptr()), G.i64(k), C.c_double(1.0), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream()))
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    L.elb200_dgemm_set_debug_flags(0)
    return 2*m*n*k/best/1e9
for (m, n, k) in [(32768, 32768, 128), (16384, 8192, 128), (16384, 16384, 256), (8192, 8192, 8192), (32768, 32768, 512)]:
    print(f"NN {m}x{n}x{k}: " + "  ".join(f"bw{bw}: {run(m, n, k, bw):.2f}" for bw in (1, 2, 4, 8, 16, 32)), flush=True)
