"""Start offset of consumer group 1 (flags bits 26..29, microseconds) on the rank-nb update shapes; 5 timings each."""
import ctypes as C, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib(); dev = torch.device("cuda:0")
L.elb200_dgemm_set_config(3)
def run(m, n, k, st, alpha=1.0, ta="N", tb="N"):
    ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
    A = torch.empty(ac, ar, dtype=torch.float64, device=dev).uniform_(-1, 1)
    B = torch.empty(bc, br, dtype=torch.float64, device=dev).uniform_(-1, 1)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    L.elb200_dgemm_set_debug_flags(st << 26)
    fn = lambda: check(L.elb200_dgemm(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), C.c_double(alpha), C.c_void_p(A.data_ptr()), G.i64(ar),
                                      C.c_void_p(B.data_ptr()), G.i64(br), C.c_double(1.0), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream()))
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    L.elb200_dgemm_set_debug_flags(0)
    return "/".join(f"{2*m*n*k/t/1e9:.1f}" for t in ts)
for (m, n, k, al, ta, tb) in [(32768, 32768, 128, 1.0, "N", "N"), (32768, 32768, 128, -1.0, "N", "N"), (32768, 32768, 128, 1.0, "N", "T"), (16384, 8192, 128, 1.0, "N", "N"),
                      (16384, 16384, 256, -1.0, "N", "T"), (32768, 32768, 128, 3.0, "N", "N"), (8192, 8192, 8192, 1.0, "N", "N")]:
    print(f"{ta}{tb} {m}x{n}x{k} alpha {al}: " + "  ".join(f"st{st}: {run(m, n, k, st, al, ta, tb)}" for st in (0, 2, 4, 6, 8, 10, 12)), flush=True)
