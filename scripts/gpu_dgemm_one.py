"""Run one DGEMM / DTRRK / ZGEMM shape a few times (for ncu captures):
   python scripts/gpu_dgemm_one.py m n k ta tb cfg reps [gemm|trrkL|trrkU|zgemm]"""
import ctypes as C, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check, c64
import gpuutil as G
L = lib()
m, n, k = (int(x) for x in sys.argv[1:4]); ta, tb = sys.argv[4], sys.argv[5]; cfg = int(sys.argv[6]); reps = int(sys.argv[7])
what = sys.argv[8] if len(sys.argv) > 8 else "gemm"
dev = torch.device("cuda:0")
ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
dt = torch.complex128 if what == "zgemm" else torch.float64
A = torch.empty(ac, ar, dtype=torch.float64, device=dev).uniform_(-1, 1).to(dt)
B = torch.empty(bc, br, dtype=torch.float64, device=dev).uniform_(-1, 1).to(dt)
Cm = torch.zeros(n, m, dtype=dt, device=dev)
L.elb200_dgemm_set_config(cfg)
for _ in range(reps):
    if what == "gemm":
        check(L.elb200_dgemm(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), C.c_double(-1.0), C.c_void_p(A.data_ptr()), G.i64(ar),
                             C.c_void_p(B.data_ptr()), G.i64(br), C.c_double(1.0), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream()))
    elif what == "zgemm":
        check(L.elb200_zgemm(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), c64(-1.0, 0.0), C.c_void_p(A.data_ptr()), G.i64(ar),
                             C.c_void_p(B.data_ptr()), G.i64(br), c64(1.0, 0.0), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream()))
    else:
        check(L.elb200_dtrrk(G.ch(what[-1]), G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), C.c_double(-1.0), C.c_void_p(A.data_ptr()), G.i64(ar),
                             C.c_void_p(B.data_ptr()), G.i64(br), C.c_double(1.0), C.c_void_p(Cm.data_ptr()), G.i64(m),
                             G.i64(0), G.i64(1), G.i64(0), G.i64(1), G.stream()))
torch.cuda.synchronize()
print("done")
