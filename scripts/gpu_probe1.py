"""First GPU trip: DMMA ceiling, cuBLAS DGEMM yardstick, elb200_dgemm/dtrrk correctness + speed."""
import ctypes as C, json, sys, time
import torch
sys.path.insert(0, ".")
from elemental_b200._lib import lib, check

L = lib()
check(L.elb200_device_check())
dev = torch.device("cuda:0")
out = {}

# 1. DMMA peak
f = C.c_double(); ms = C.c_float()
for it in (2000, 20000):
    check(L.elb200_dmma_peak(it, C.byref(f), C.byref(ms)))
    print(f"dmma_peak iters={it}: {f.value/1e12:.2f} TFLOP/s in {ms.value:.2f} ms")
out["dmma_peak_tflops"] = f.value / 1e12

# 2. cuBLAS DGEMM yardstick
def time_fn(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    t = time_fn(lambda: torch.matmul(a, b))
    print(f"cuBLAS dgemm n={n}: {2*n**3/t/1e9:.2f} TFLOP/s ({t:.2f} ms)")
    out[f"cublas_dgemm_{n}_tflops"] = 2 * n**3 / t / 1e9

def dgemm(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
    s = torch.cuda.current_stream().cuda_stream
    check(L.elb200_dgemm(C.c_char(ta.encode()), C.c_char(tb.encode()), C.c_int64(m), C.c_int64(n), C.c_int64(k),
                         C.c_double(alpha), C.c_void_p(A.data_ptr()), C.c_int64(lda), C.c_void_p(B.data_ptr()), C.c_int64(ldb),
                         C.c_double(beta), C.c_void_p(Cm.data_ptr()), C.c_int64(ldc), C.c_void_p(s)))

def dtrrk(uplo, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc, rs, rst, cs, cst):
    s = torch.cuda.current_stream().cuda_stream
    check(L.elb200_dtrrk(C.c_char(uplo.encode()), C.c_char(ta.encode()), C.c_char(tb.encode()), C.c_int64(m), C.c_int64(n), C.c_int64(k),
                         C.c_double(alpha), C.c_void_p(A.data_ptr()), C.c_int64(lda), C.c_void_p(B.data_ptr()), C.c_int64(ldb),
                         C.c_double(beta), C.c_void_p(Cm.data_ptr()), C.c_int64(ldc),
                         C.c_int64(rs), C.c_int64(rst), C.c_int64(cs), C.c_int64(cst), C.c_void_p(s)))

# column-major helper: a torch tensor of shape (cols, ld) row-major == column-major (ld x cols)
def colmajor(rows, cols, ld, off=0):
    buf = torch.empty(cols * ld + off + 2, dtype=torch.float64, device=dev).uniform_(-1, 1)
    v = buf[off:off + cols * ld].view(cols, ld)
    return buf, v  # v[j, i] = M(i, j)

ok_all = True
torch.manual_seed(0)
for (m, n, k) in [(128, 128, 16), (256, 384, 64), (1, 1, 1), (7, 5, 3), (130, 257, 45), (513, 129, 257), (1000, 900, 333)]:
    for ta in "NT":
        for tb in "NT":
            for off in (0, 1):
                ar, ac = (m, k) if ta == "N" else (k, m)
                br, bc = (k, n) if tb == "N" else (n, k)
                lda, ldb, ldc = ar + 3 + off, br + 2, m + 1 + off
                _, Av = colmajor(ar, ac, lda, off); _, Bv = colmajor(br, bc, ldb, 0); _, Cv = colmajor(m, n, ldc, off)
                A = Av[:, :ar].t(); B = Bv[:, :br].t(); C0 = Cv[:, :m].t().clone()
                opA = A if ta == "N" else A.t(); opB = B if tb == "N" else B.t()
                ref = 3.0 * opA @ opB + 4.0 * C0
                dgemm(ta, tb, m, n, k, 3.0, Av, lda, Bv, ldb, 4.0, Cv, ldc)
                got = Cv[:, :m].t()
                err = (got - ref).abs().max().item() / max(1.0, ref.abs().max().item())
                pad_ok = True
                ok = err < 1e-13 * max(k, 1)
                ok_all &= ok
                if not ok:
                    print(f"FAIL gemm {ta}{tb} m={m} n={n} k={k} off={off}: err={err:.3e}")
print("gemm correctness:", "OK" if ok_all else "FAILED")
out["gemm_ok"] = ok_all

# trrk with cyclic shifts
ok_t = True
for (m, n, k, rs, rst, cs, cst) in [(300, 300, 40, 0, 1, 0, 1), (257, 131, 33, 1, 2, 3, 4), (131, 257, 64, 0, 2, 1, 4), (640, 384, 256, 1, 2, 2, 4)]:
    for uplo in "LU":
        for ta, tb in (("T", "N"), ("N", "T"), ("N", "N"), ("T", "T")):
            ar, ac = (m, k) if ta == "N" else (k, m)
            br, bc = (k, n) if tb == "N" else (n, k)
            lda, ldb, ldc = ar + 1, br + 2, m + 3
            _, Av = colmajor(ar, ac, lda); _, Bv = colmajor(br, bc, ldb); _, Cv = colmajor(m, n, ldc)
            A = Av[:, :ar].t(); B = Bv[:, :br].t(); C0 = Cv[:, :m].t().clone()
            opA = A if ta == "N" else A.t(); opB = B if tb == "N" else B.t()
            full = -1.0 * opA @ opB + 1.0 * C0
            gi = rs + rst * torch.arange(m, device=dev)[:, None]; gj = cs + cst * torch.arange(n, device=dev)[None, :]
            mask = (gi >= gj) if uplo == "L" else (gi <= gj)
            ref = torch.where(mask, full, C0)
            dtrrk(uplo, ta, tb, m, n, k, -1.0, Av, lda, Bv, ldb, 1.0, Cv, ldc, rs, rst, cs, cst)
            got = Cv[:, :m].t()
            err = (got - ref).abs().max().item()
            ok = err < 1e-12 * k
            ok_t &= ok
            if not ok:
                print(f"FAIL trrk {uplo} {ta}{tb} m={m} n={n} k={k}: err={err:.3e}")
print("trrk correctness:", "OK" if ok_t else "FAILED")
out["trrk_ok"] = ok_t

# speed
for (m, n, k) in [(4096, 4096, 4096), (8192, 8192, 8192), (16384, 8192, 128), (16384, 16384, 256), (16384, 8192, 256)]:
    for ta, tb in (("N", "N"), ("N", "T"), ("T", "N"), ("T", "T")):
        ar, ac = (m, k) if ta == "N" else (k, m)
        br, bc = (k, n) if tb == "N" else (n, k)
        A = torch.empty(ac, ar, dtype=torch.float64, device=dev).uniform_(-1, 1)
        B = torch.empty(bc, br, dtype=torch.float64, device=dev).uniform_(-1, 1)
        Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
        t = time_fn(lambda: dgemm(ta, tb, m, n, k, 1.0, A, ar, B, br, 1.0, Cm, m), reps=3)
        tf = 2 * m * n * k / t / 1e9
        print(f"elb200 dgemm {ta}{tb} {m}x{n}x{k}: {tf:.2f} TFLOP/s ({t:.3f} ms)")
        out[f"dgemm_{ta}{tb}_{m}_{n}_{k}_tflops"] = tf
        del A, B, Cm
# trrk speed (lower, T N) like Cholesky's update
m = n = 16384; k = 256
A = torch.empty(m, k, dtype=torch.float64, device=dev).uniform_(-1, 1)
B = torch.empty(n, k, dtype=torch.float64, device=dev).uniform_(-1, 1)
Cm = torch.zeros(n, m, dtype=torch.float64, device=dev)
t = time_fn(lambda: dtrrk("L", "T", "N", m, n, k, -1.0, A, k, B, k, 1.0, Cm, m, 0, 1, 0, 1), reps=3)
print(f"elb200 dtrrk L TN {m}x{n}x{k}: {m*n*k/t/1e9:.2f} TFLOP/s useful ({t:.3f} ms)")
out["dtrrk_tflops"] = m * n * k / t / 1e9
import os
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/probe1.json", "w"), indent=1)
print(json.dumps(out))
