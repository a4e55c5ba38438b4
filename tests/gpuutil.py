"""ctypes helpers for calling the leaf C-ABI (include/elb200_blas.h) with torch CUDA tensors."""
import ctypes as C

import numpy as np
import torch

from elemental_b200._lib import c32, c64, check, lib

SUF = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c", np.dtype(np.complex128): "z"}
TORCH = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
         np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def sc(dt, x):
    dt = np.dtype(dt)
    if dt == np.float32:
        return C.c_float(float(np.real(x)))
    if dt == np.float64:
        return C.c_double(float(np.real(x)))
    if dt == np.complex64:
        return c32(float(np.real(x)), float(np.imag(x)))
    return c64(float(np.real(x)), float(np.imag(x)))


def rsc(dt, x):
    return C.c_float(float(x)) if np.dtype(dt) in (np.float32, np.complex64) else C.c_double(float(x))


def ch(c):
    return C.c_char(c.encode())


def i64(x):
    return C.c_int64(int(x))


class DevMat:
    """column-major matrix on the device with leading dimension ld (and a guard band)."""

    def __init__(self, a: np.ndarray, ld=None, offset=0):
        a = np.asarray(a)
        self.m, self.n = a.shape
        self.dt = a.dtype
        self.ld = int(ld if ld is not None else max(self.m, 1))
        host = np.full(offset + self.ld * max(self.n, 1) + 8, 777.0, dtype=a.dtype)
        for j in range(self.n):
            host[offset + j * self.ld: offset + j * self.ld + self.m] = a[:, j]
        self.host0 = host.copy()
        self.offset = offset
        self.t = torch.from_numpy(host).cuda()

    @property
    def ptr(self):
        return C.c_void_p(self.t.data_ptr() + self.offset * self.t.element_size())

    def get(self):
        host = self.t.cpu().numpy()
        out = np.empty((self.m, self.n), dtype=self.dt)
        for j in range(self.n):
            out[:, j] = host[self.offset + j * self.ld: self.offset + j * self.ld + self.m]
        return out

    def padding_untouched(self):
        host = self.t.cpu().numpy()
        mask = np.ones(host.shape, dtype=bool)
        for j in range(self.n):
            mask[self.offset + j * self.ld: self.offset + j * self.ld + self.m] = False
        return np.array_equal(host[mask], self.host0[mask])


def gemm(ta, tb, alpha, A: DevMat, B: DevMat, beta, Cm: DevMat, k, fn=None):
    fn = getattr(lib(), fn or f"elb200_{SUF[Cm.dt]}gemm")
    check(fn(ch(ta), ch(tb), i64(Cm.m), i64(Cm.n), i64(k), sc(Cm.dt, alpha), A.ptr, i64(A.ld), B.ptr, i64(B.ld),
             sc(Cm.dt, beta), Cm.ptr, i64(Cm.ld), stream()), "gemm")


def trrk(uplo, ta, tb, alpha, A, B, beta, Cm, k, rs=0, rst=1, cs=0, cst=1):
    fn = getattr(lib(), f"elb200_{SUF[Cm.dt]}trrk")
    check(fn(ch(uplo), ch(ta), ch(tb), i64(Cm.m), i64(Cm.n), i64(k), sc(Cm.dt, alpha), A.ptr, i64(A.ld), B.ptr,
             i64(B.ld), sc(Cm.dt, beta), Cm.ptr, i64(Cm.ld), i64(rs), i64(rst), i64(cs), i64(cst), stream()), "trrk")


def trsm(side, uplo, trans, diag, alpha, A, B):
    fn = getattr(lib(), f"elb200_{SUF[B.dt]}trsm")
    check(fn(ch(side), ch(uplo), ch(trans), ch(diag), i64(B.m), i64(B.n), sc(B.dt, alpha), A.ptr, i64(A.ld), B.ptr,
             i64(B.ld), stream()), "trsm")


def potrf(uplo, A):
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    fn = getattr(lib(), f"elb200_{SUF[A.dt]}potrf")
    check(fn(ch(uplo), i64(A.m), A.ptr, i64(A.ld), C.c_void_p(info.data_ptr()), stream()), "potrf")
    return int(info.item())


def herk(uplo, trans, alpha, A, beta, Cm, k):
    dt = Cm.dt
    if dt.kind == "c":
        fn = getattr(lib(), f"elb200_{SUF[dt]}herk")
        check(fn(ch(uplo), ch(trans), i64(Cm.m), i64(k), rsc(dt, alpha), A.ptr, i64(A.ld), rsc(dt, beta), Cm.ptr,
                 i64(Cm.ld), stream()), "herk")
    else:
        fn = getattr(lib(), f"elb200_{SUF[dt]}syrk")
        check(fn(ch(uplo), ch(trans), i64(Cm.m), i64(k), sc(dt, alpha), A.ptr, i64(A.ld), sc(dt, beta), Cm.ptr,
                 i64(Cm.ld), stream()), "syrk")


def rand(rng, m, n, dt):
    dt = np.dtype(dt)
    a = rng.uniform(-1, 1, (m, n))
    if dt.kind == "c":
        a = a + 1j * rng.uniform(-1, 1, (m, n))
    return np.asfortranarray(a.astype(dt))


def eps(dt):
    return np.finfo(np.dtype(dt).char.lower() if np.dtype(dt).kind == "c" else dt).eps
