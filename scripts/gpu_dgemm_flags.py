"""TMA DGEMM kernel under debug flag combinations: bit0 stagger the groups, bit1 masked epilogue everywhere, bit2 one group only,
bit4 no L2-reduction epilogue, bit5/6 start offsets per SM / group, bit7 NO C update (diagnostic), bit8 plain stores (diagnostic), bit9 fragment prefetch one k-step ahead (experiment).
usage: python scripts/gpu_dgemm_flags.py 0 128 256 512 16"""
import ctypes as C, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib(); dev = torch.device("cuda:0")
def time_fn(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
L.elb200_dgemm_set_config(3)
for (m, n, k) in [(8192, 8192, 8192), (32768, 32768, 128), (16384, 16384, 256), (4096, 4096, 256)]:
    A = torch.empty(k, m, dtype=torch.float64, device=dev).uniform_(-1, 1)
    B = torch.empty(n, k, dtype=torch.float64, device=dev).uniform_(-1, 1)
    Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
    row = []
    for flags in [int(x) for x in sys.argv[1:]]:
        L.elb200_dgemm_set_debug_flags(flags)
        t = time_fn(lambda: check(L.elb200_dgemm(G.ch("N"), G.ch("N"), G.i64(m), G.i64(n), G.i64(k), C.c_double(1.0), C.c_void_p(A.data_ptr()), G.i64(m),
                                               C.c_void_p(B.data_ptr()), G.i64(k), C.c_double(1.0), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream())))
        row.append(f"flags{flags}: {2*m*n*k/t/1e9:.2f}")
    print(f"NN {m}x{n}x{k}  " + "  ".join(row), flush=True)
    del A, B, Cm
