"""Run one DGEMM shape a few times (for ncu captures): python scripts/gpu_dgemm_one.py m n k ta tb cfg reps"""
import ctypes as C, sys
import torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from elemental_b200._lib import lib, check
import gpuutil as G
L = lib()
m, n, k = (int(x) for x in sys.argv[1:4]); ta, tb = sys.argv[4], sys.argv[5]; cfg = int(sys.argv[6]); reps = int(sys.argv[7])
dev = torch.device("cuda:0")
ar, ac = (m, k) if ta == "N" else (k, m); br, bc = (k, n) if tb == "N" else (n, k)
A = torch.empty(ac, ar, dtype=torch.float64, device=dev).uniform_(-1, 1)
B = torch.empty(bc, br, dtype=torch.float64, device=dev).uniform_(-1, 1)
Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
L.elb200_dgemm_set_config(cfg)
for _ in range(reps):
    check(L.elb200_dgemm(G.ch(ta), G.ch(tb), G.i64(m), G.i64(n), G.i64(k), C.c_double(1.0), C.c_void_p(A.data_ptr()), G.i64(ar),
                         C.c_void_p(B.data_ptr()), G.i64(br), C.c_double(1.0), C.c_void_p(Cm.data_ptr()), G.i64(m), G.stream()))
torch.cuda.synchronize()
print("done")
