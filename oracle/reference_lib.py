"""ctypes wrapper over oracle/_ref/libElRef.so -- the REFERENCE itself (Elemental's own
sources, built by oracle/refbuild/Makefile).  TEST INFRASTRUCTURE ONLY: used to pin
oracle/elemental_oracle.py, to generate tests/golden/, and as bench.py's CPU baseline."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

REF_LIB = Path(__file__).resolve().parent / "_ref" / "libElRef.so"
# the same reference objects linked against libelb200.so's Fortran BLAS symbols with device-addressable buffers
# (oracle/refbuild/Makefile target `dev`): the depth-1 integration of INTEGRATION.md, used by one GPU test
REF_DEV_LIB = Path(__file__).resolve().parent / "_ref" / "libElRefDev.so"
_lib = None


_DEV = False
_keep = []   # pinned staging buffers stay alive until the next call


def use_device_build(on: bool = True):
    """Switch every wrapper in this module between libElRef.so (CPU, the oracle) and libElRefDev.so.  The driver
    ATTACHES the caller's arrays to DistMatrix objects, so in the device build they are staged through page-locked
    (device-mapped) memory first -- a numpy array is pageable host memory the GPU cannot address."""
    global _lib, REF_LIB, _DEV
    _lib = None
    _DEV = on
    REF_LIB = (REF_DEV_LIB if on else Path(__file__).resolve().parent / "_ref" / "libElRef.so")


def _pin(a):
    """F-ordered copy of `a` in page-locked memory (torch's pinned allocator = cudaHostAlloc)."""
    import torch
    a = np.asarray(a)
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
           np.dtype(np.complex64): torch.complex64, np.dtype(np.complex128): torch.complex128}[a.dtype]
    t = torch.empty(tuple(reversed(a.shape)), dtype=tdt, pin_memory=True)
    v = t.numpy().T          # Fortran-ordered view of the pinned block
    v[...] = a
    _keep.append(t)
    return v


class _InOut:
    """the array the reference updates in place: the caller's array itself, or its pinned stand-in"""

    def __init__(self, a):
        self.user = a
        self.buf = _pin(a) if _DEV else a

    def done(self):
        if _DEV:
            self.user[...] = self.buf
            del _keep[:]
        return self.user


class ReferenceUnavailable(RuntimeError):
    pass


class RefNonHPD(Exception):
    pass


def available() -> bool:
    return REF_LIB.exists()


def lib():
    global _lib
    if _lib is None:
        if not REF_LIB.exists():
            raise ReferenceUnavailable(f"{REF_LIB} not built (make -C oracle/refbuild)")
        _lib = C.CDLL(str(REF_LIB))
        _lib.elref_last_error.restype = C.c_char_p
        _lib.elref_blas_corename.restype = C.c_char_p
        _lib.elref_blas_config.restype = C.c_char_p
    return _lib


def _chk(rc):
    if rc == 2:
        raise RefNonHPD(lib().elref_last_error().decode())
    if rc != 0:
        raise RuntimeError("reference call failed: " + lib().elref_last_error().decode())


_SUF = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
        np.dtype(np.complex128): "z"}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f(a, dtype=None):
    a = np.asfortranarray(a, dtype=dtype)
    return _pin(a) if _DEV else a


def _scalar(x, dt):
    """real types pass by value, complex through a 2-element real array"""
    if dt.kind == "c":
        r = np.array([np.real(x), np.imag(x)], dtype=np.float64 if dt == np.complex128 else np.float32)
        return r, _p(r)
    return None, (C.c_double(float(x)) if dt == np.float64 else C.c_float(float(x)))


def info():
    L = lib()
    return {"corename": L.elref_blas_corename().decode(), "threads": L.elref_get_threads(),
            "config": L.elref_blas_config().decode()}


def set_threads(n: int):
    lib().elref_set_threads(int(n))


def gemm(oa, ob, alpha, A, B, beta, Cm, nb=128, alg=0):
    """El::Gemm on DistMatrix<T,MC,MR> over the 1x1 default grid (in place on Cm, Fortran order)."""
    dt = Cm.dtype
    A, B = _f(A, dt), _f(B, dt)
    assert Cm.flags.f_contiguous
    m, n = Cm.shape
    k = A.shape[1] if oa.upper() == "N" else A.shape[0]
    ka, av = _scalar(alpha, dt)
    kb, bv = _scalar(beta, dt)
    fn = getattr(lib(), "elref_gemm_" + _SUF[dt])
    io = _InOut(Cm)
    _chk(fn(C.c_char(oa.encode()), C.c_char(ob.encode()), m, n, k, av, _p(A), A.shape[0], _p(B), B.shape[0],
            bv, _p(io.buf), Cm.shape[0], int(nb), int(alg)))
    return io.done()


def cholesky(uplo, A, nb=128):
    assert A.flags.f_contiguous
    fn = getattr(lib(), "elref_cholesky_" + _SUF[A.dtype])
    io = _InOut(A)
    _chk(fn(C.c_char(uplo.encode()), A.shape[0], _p(io.buf), A.shape[0], int(nb)))
    return io.done()


def hpd_solve(uplo, orient, A, B, nb=128):
    assert B.flags.f_contiguous
    A = _f(A, B.dtype)
    fn = getattr(lib(), "elref_hpdsolve_" + _SUF[B.dtype])
    io = _InOut(B)
    _chk(fn(C.c_char(uplo.encode()), C.c_char(orient.encode()), A.shape[0], B.shape[1], _p(A), A.shape[0],
            _p(io.buf), B.shape[0], int(nb)))
    return io.done()


def trsm(side, uplo, orient, diag, alpha, A, B, nb=128):
    assert B.flags.f_contiguous
    A = _f(A, B.dtype)
    ka, av = _scalar(alpha, B.dtype)
    fn = getattr(lib(), "elref_trsm_" + _SUF[B.dtype])
    io = _InOut(B)
    _chk(fn(C.c_char(side.encode()), C.c_char(uplo.encode()), C.c_char(orient.encode()), C.c_char(diag.encode()),
            B.shape[0], B.shape[1], av, _p(A), A.shape[0], _p(io.buf), B.shape[0], int(nb)))
    return io.done()


def herk(uplo, orient, alpha, A, beta, Cm, nb=128):
    assert Cm.flags.f_contiguous
    A = _f(A, Cm.dtype)
    n = Cm.shape[0]
    k = A.shape[1] if orient.upper() == "N" else A.shape[0]
    fn = getattr(lib(), "elref_herk_" + _SUF[Cm.dtype])
    _chk(fn(C.c_char(uplo.encode()), C.c_char(orient.encode()), n, k, C.c_double(alpha), _p(A), A.shape[0],
            C.c_double(beta), _p(Cm), n, int(nb)))
    return Cm


def trrk(uplo, oa, ob, alpha, A, B, beta, Cm, nb=128):
    assert Cm.flags.f_contiguous and Cm.dtype == np.float64
    A, B = _f(A, np.float64), _f(B, np.float64)
    n = Cm.shape[0]
    k = A.shape[1] if oa.upper() == "N" else A.shape[0]
    _chk(lib().elref_trrk_d(C.c_char(uplo.encode()), C.c_char(oa.encode()), C.c_char(ob.encode()), n, k,
                            C.c_double(alpha), _p(A), A.shape[0], _p(B), B.shape[0], C.c_double(beta), _p(Cm), n,
                            int(nb)))
    return Cm


# ---- siblings SURVEY.md section 8f ranks next (double and complex double) ----
def syr2k(uplo, orient, alpha, A, B, beta, Cm, conjugate=False, nb=128):
    """El::Syr2k / El::Her2k (conjugate=True) on the 1x1 grid, in place on the uplo triangle of Cm."""
    assert Cm.flags.f_contiguous
    dt = Cm.dtype
    A, B = _f(A, dt), _f(B, dt)
    n = Cm.shape[0]
    k = A.shape[1] if orient.upper() == "N" else A.shape[0]
    ka, av = _scalar(alpha, dt)
    kb, bv = _scalar(beta, dt)
    fn = getattr(lib(), "elref_syr2k_" + _SUF[dt])
    _chk(fn(C.c_char(uplo.encode()), C.c_char(orient.encode()), n, k, av, _p(A), A.shape[0], _p(B), B.shape[0], bv,
            _p(Cm), n, int(bool(conjugate)), int(nb)))
    return Cm


def symm(side, uplo, alpha, A, B, beta, Cm, conjugate=False, nb=128):
    """El::Symm / El::Hemm (conjugate=True): only the uplo triangle of A is referenced."""
    assert Cm.flags.f_contiguous
    dt = Cm.dtype
    A, B = _f(A, dt), _f(B, dt)
    m, n = Cm.shape
    ka, av = _scalar(alpha, dt)
    kb, bv = _scalar(beta, dt)
    fn = getattr(lib(), "elref_symm_" + _SUF[dt])
    _chk(fn(C.c_char(side.encode()), C.c_char(uplo.encode()), m, n, av, _p(A), A.shape[0], _p(B), B.shape[0], bv,
            _p(Cm), m, int(bool(conjugate)), int(nb)))
    return Cm


def trmm(side, uplo, orient, diag, alpha, A, B, nb=128):
    """El::Trmm: B := alpha op(tri(A)) B (LEFT) or alpha B op(tri(A)) (RIGHT), in place."""
    assert B.flags.f_contiguous
    A = _f(A, B.dtype)
    ka, av = _scalar(alpha, B.dtype)
    fn = getattr(lib(), "elref_trmm_" + _SUF[B.dtype])
    _chk(fn(C.c_char(side.encode()), C.c_char(uplo.encode()), C.c_char(orient.encode()), C.c_char(diag.encode()),
            B.shape[0], B.shape[1], av, _p(A), A.shape[0], _p(B), B.shape[0], int(nb)))
    return B


class RefSingular(Exception):
    pass


def _chk3(rc):
    if rc == 3:
        raise RefSingular(lib().elref_last_error().decode())
    _chk(rc)


# ---- LU, pivoted Cholesky, CholeskyMod, LinearSolve (reference's own code) ----
def lu(A, nb=128):
    """El::LU(A) without pivoting (src/lapack_like/factor/LU.cpp:21-99), in place: unit-lower L and U packed."""
    assert A.flags.f_contiguous
    fn = getattr(lib(), "elref_lu_" + _SUF[A.dtype])
    io = _InOut(A)
    _chk3(fn(A.shape[0], A.shape[1], _p(io.buf), A.shape[0], int(nb)))
    return io.done()


def lu_piv(A, nb=128):
    """El::LU(A, P) with partial pivoting (LU.cpp:104-220).  Returns (packed LU in place, preimage vector p):
    row i of P A is row p[i] of A, i.e. (P A) = A[p, :] = L U."""
    assert A.flags.f_contiguous
    pre = np.zeros(A.shape[0], dtype=np.int32)
    fn = getattr(lib(), "elref_lu_piv_" + _SUF[A.dtype])
    _chk3(fn(A.shape[0], A.shape[1], _p(A), A.shape[0], pre.ctypes.data_as(C.c_void_p), int(nb)))
    return A, pre.astype(np.int64)


def lu_piv_solve(orient, A, B, nb=128):
    """LU(A, P) on a copy of A, then lu::SolveAfter(orientation, LU, P, B) (LU/SolveAfter.hpp): B := op(A)^{-1} B."""
    assert B.flags.f_contiguous
    A = _f(A, B.dtype)
    fn = getattr(lib(), "elref_lu_piv_solve_" + _SUF[B.dtype])
    _chk3(fn(C.c_char(orient.encode()), A.shape[0], B.shape[1], _p(A), A.shape[0], _p(B), B.shape[0], int(nb)))
    return B


def linear_solve(A, B, nb=128):
    """El::LinearSolve(A, B) (src/lapack_like/solve/Linear.cpp: RowEchelon + back substitution)."""
    assert B.flags.f_contiguous
    A = _f(A, B.dtype)
    fn = getattr(lib(), "elref_linear_solve_" + _SUF[B.dtype])
    _chk3(fn(A.shape[0], B.shape[1], _p(A), A.shape[0], _p(B), B.shape[0], int(nb)))
    return B


def cholesky_piv(uplo, A, nb=128):
    """El::Cholesky(uplo, A, P) with diagonal pivoting (Cholesky/PivotedLowerVariant3.hpp, PivotedUpperVariant3.hpp).
    Returns (factor in the uplo triangle, preimage vector p): A[p][:, p] = L L^H (or U^H U)."""
    assert A.flags.f_contiguous
    pre = np.zeros(A.shape[0], dtype=np.int32)
    fn = getattr(lib(), "elref_cholesky_piv_" + _SUF[A.dtype])
    _chk(fn(C.c_char(uplo.encode()), A.shape[0], _p(A), A.shape[0], pre.ctypes.data_as(C.c_void_p), int(nb)))
    return A, pre.astype(np.int64)


def cholesky_piv_solve(uplo, orient, A, B, nb=128):
    assert B.flags.f_contiguous
    A = _f(A, B.dtype)
    fn = getattr(lib(), "elref_cholesky_piv_solve_" + _SUF[B.dtype])
    _chk(fn(C.c_char(uplo.encode()), C.c_char(orient.encode()), A.shape[0], B.shape[1], _p(A), A.shape[0], _p(B),
            B.shape[0], int(nb)))
    return B


def cholesky_mod(uplo, T, alpha, V, nb=128):
    """El::CholeskyMod(uplo, T, alpha, V) (Cholesky/LowerMod.hpp, UpperMod.hpp): the factor of T T^H + alpha V V^H
    (LOWER) or T^H T + alpha V V^H (UPPER) overwrites T; V is overwritten with workspace."""
    assert T.flags.f_contiguous and V.flags.f_contiguous
    fn = getattr(lib(), "elref_cholesky_mod_" + _SUF[T.dtype])
    _chk(fn(C.c_char(uplo.encode()), T.shape[0], V.shape[1], _p(T), T.shape[0],
            C.c_double(float(alpha)), _p(V), V.shape[0], int(nb)))
    return T
