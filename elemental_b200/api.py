"""Python mirror of the reference's interface for the Gemm / Cholesky / HPDSolve path.

Same names and argument order as the reference's own Python bindings over its C API
(python/core/Grid.py, python/core/DistMatrix.py, python/blas_like/level3.py: `Gemm`,
`Herk`, `Trsm`; python/lapack_like/factor.py: `Cholesky`; python/lapack_like/solve.py:
`HPDSolve`), bound to libelb200.so through include/elb200_El.h.  Local matrices live in
device memory; everything is enqueued on torch's current CUDA stream.  There is no CPU
fallback: without the built library and a B200 the calls raise.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Elb200Error, c32, c64, lib

# ---- enums (values of include/El/core/types.hpp) ----
MC, MD, MR, VC, VR, STAR, CIRC = range(7)
NORMAL, TRANSPOSE, ADJOINT = range(3)
LOWER, UPPER = range(2)
LEFT, RIGHT = range(2)
NON_UNIT, UNIT = range(2)
ROW_MAJOR, COLUMN_MAJOR = range(2)
GEMM_DEFAULT, GEMM_SUMMA_A, GEMM_SUMMA_B, GEMM_SUMMA_C, GEMM_SUMMA_DOT, GEMM_CANNON = range(6)
DIST_NAMES = {MC: "MC", MR: "MR", VC: "VC", VR: "VR", STAR: "STAR"}

EL_NON_HPD_ERROR = 100
EL_SINGULAR_ERROR = 101


class NonHPDMatrixException(Elb200Error):
    pass


class SingularMatrixException(Elb200Error):
    pass


class LogicError(Elb200Error):
    pass


_SUFFIX = {np.dtype(np.float32): "s", np.dtype(np.float64): "d", np.dtype(np.complex64): "c",
           np.dtype(np.complex128): "z"}


def _check(rc: int, what: str):
    if rc == 0:
        return
    L = lib()
    L.ElLastErrorMessage.restype = C.c_char_p
    msg = (L.ElLastErrorMessage() or b"").decode()
    if rc == EL_NON_HPD_ERROR:
        raise NonHPDMatrixException(msg)
    if rc == EL_SINGULAR_ERROR:
        raise SingularMatrixException(msg)
    if rc == 4:
        raise LogicError(f"{what}: {msg}")
    raise Elb200Error(f"{what} failed (ElError {rc}): {msg}")


def _sync_stream():
    """Point the layer at torch's current stream (cheap; done before every call)."""
    import torch

    lib().ElSetStream(C.c_void_p(torch.cuda.current_stream().cuda_stream))


def Initialize():
    _check(lib().ElInitialize(None, None), "ElInitialize")
    _sync_stream()


def Blocksize() -> int:
    v = C.c_int()
    _check(lib().ElBlocksize(C.byref(v)), "ElBlocksize")
    return v.value


def SetBlocksize(b: int):
    _check(lib().ElSetBlocksize(int(b)), "ElSetBlocksize")


def PushBlocksizeStack(b: int):
    _check(lib().ElPushBlocksizeStack(int(b)), "ElPushBlocksizeStack")


def PopBlocksizeStack():
    _check(lib().ElPopBlocksizeStack(), "ElPopBlocksizeStack")


def SetGemmDotBlocksize(b: int):
    """Edge of the C blocks of SUMMA_Dot (reference: 2000, Gemm/NN.hpp:233); 0 = sized for HBM."""
    _check(lib().ElSetGemmDotBlocksize(int(b)), "ElSetGemmDotBlocksize")


def Synchronize():
    _check(lib().ElSynchronize(), "ElSynchronize")


def RedistStats(reset: bool = False) -> dict:
    out = (C.c_uint64 * 8)()
    _check(lib().ElRedistStats(out, C.c_bool(reset)), "ElRedistStats")
    keys = ("copies", "messages", "bytesSent", "packLaunches", "zeroCopySends", "reduceScatters", "allGathers", "p2pPushes")
    return dict(zip(keys, [int(x) for x in out]))


class Grid:
    """r x c process grid, one process per GPU (src/core/Grid.cpp).  With torch.distributed
    initialised and world_size > 1 the ncclUniqueId is broadcast through it."""

    def __init__(self, height: int | None = None, order: int = COLUMN_MAJOR):
        import torch
        import torch.distributed as dist

        Initialize()
        self._h = C.c_void_p()
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            _check(lib().ElGridCreateTrivial(C.byref(self._h)), "ElGridCreateTrivial")
        else:
            rank = dist.get_rank()
            uid = (C.c_ubyte * 128)()
            if rank == 0:
                _check(lib().ElNcclUniqueId(uid), "ElNcclUniqueId")
            # move the id through whatever backend the process group has
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
            t = torch.tensor(list(uid), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)
            uid = (C.c_ubyte * 128)(*t.cpu().tolist())
            if height is None or height <= 0:
                height = 0
            _check(lib().ElGridCreateNccl(uid, rank, world, int(height), int(order), C.byref(self._h)),
                   "ElGridCreateNccl")

    def _geti(self, fn):
        v = C.c_int()
        _check(getattr(lib(), fn)(self._h, C.byref(v)), fn)
        return v.value

    def Height(self): return self._geti("ElGridHeight")
    def Width(self): return self._geti("ElGridWidth")
    def Size(self): return self._geti("ElGridSize")
    def Rank(self): return self._geti("ElGridRank")
    def Row(self): return self._geti("ElGridRow")
    def Col(self): return self._geti("ElGridCol")
    def VCRank(self): return self._geti("ElGridVCRank")
    def VRRank(self): return self._geti("ElGridVRRank")

    def Destroy(self):
        if self._h:
            lib().ElGridDestroy(self._h)
            self._h = C.c_void_p()


def _scalar(dtype, x):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return C.c_float(float(np.real(x)))
    if dtype == np.float64:
        return C.c_double(float(np.real(x)))
    if dtype == np.complex64:
        return c32(float(np.real(x)), float(np.imag(x)))
    return c64(float(np.real(x)), float(np.imag(x)))


def _real(dtype, x):
    return C.c_float(float(x)) if np.dtype(dtype) in (np.float32, np.complex64) else C.c_double(float(x))


class DistMatrix:
    """DistMatrix<T,U,V> over a Grid; `dtype` selects T (float32/float64/complex64/complex128)."""

    def __init__(self, dtype=np.float64, colDist: int = MC, rowDist: int = MR, grid: Grid | None = None,
                 height: int = 0, width: int = 0):
        self.dtype = np.dtype(dtype)
        self.suf = _SUFFIX[self.dtype]
        self.grid = grid if grid is not None else default_grid()
        self.colDist, self.rowDist = colDist, rowDist
        self._h = C.c_void_p()
        self._keep = None  # tensors / parents whose memory this matrix aliases
        _check(self._fn("ElDistMatrixCreateSpecific")(colDist, rowDist, self.grid._h, C.byref(self._h)),
               "ElDistMatrixCreateSpecific")
        if height or width:
            self.Resize(height, width)

    def _fn(self, name):
        return getattr(lib(), f"{name}_{self.suf}")

    def __del__(self):
        try:
            if self._h:
                self._fn("ElDistMatrixDestroy")(self._h)
        except Exception:
            pass

    def _geti(self, name):
        v = C.c_int()
        _check(self._fn(name)(self._h, C.byref(v)), name)
        return v.value

    def Height(self): return self._geti("ElDistMatrixHeight")
    def Width(self): return self._geti("ElDistMatrixWidth")
    def LocalHeight(self): return self._geti("ElDistMatrixLocalHeight")
    def LocalWidth(self): return self._geti("ElDistMatrixLocalWidth")
    def LDim(self): return self._geti("ElDistMatrixLDim")
    def ColAlign(self): return self._geti("ElDistMatrixColAlign")
    def RowAlign(self): return self._geti("ElDistMatrixRowAlign")
    def ColShift(self): return self._geti("ElDistMatrixColShift")
    def RowShift(self): return self._geti("ElDistMatrixRowShift")
    def ColStride(self): return self._geti("ElDistMatrixColStride")
    def RowStride(self): return self._geti("ElDistMatrixRowStride")

    def Resize(self, h, w):
        _sync_stream()
        _check(self._fn("ElDistMatrixResize")(self._h, int(h), int(w)), "ElDistMatrixResize")
        return self

    def Empty(self):
        _check(self._fn("ElDistMatrixEmpty")(self._h), "ElDistMatrixEmpty")

    def Align(self, colAlign, rowAlign, constrain=True):
        _check(self._fn("ElDistMatrixAlign")(self._h, int(colAlign), int(rowAlign), C.c_bool(constrain)),
               "ElDistMatrixAlign")
        return self

    def AlignWith(self, other: "DistMatrix"):
        _check(self._fn("ElDistMatrixAlignWith")(self._h, other._h), "ElDistMatrixAlignWith")
        return self

    def Attach(self, height, width, tensor, ldim, colAlign=0, rowAlign=0):
        """Attach to a torch CUDA tensor holding the local column-major matrix."""
        self._keep = tensor
        _check(self._fn("ElDistMatrixAttach")(self._h, int(height), int(width), self.grid._h, int(colAlign),
                                              int(rowAlign), C.c_void_p(tensor.data_ptr()), int(ldim), 0),
               "ElDistMatrixAttach")
        return self

    def View(self, parent: "DistMatrix", i, j, height, width):
        self._keep = parent
        _check(self._fn("ElDistMatrixView")(self._h, parent._h, int(i), int(j), int(height), int(width)),
               "ElDistMatrixView")
        return self

    def Buffer(self) -> int:
        p = C.c_void_p()
        _check(self._fn("ElDistMatrixLockedBuffer")(self._h, C.byref(p)), "ElDistMatrixLockedBuffer")
        return p.value or 0

    # ---- host transfers (synchronous) ----
    def LocalToHost(self) -> np.ndarray:
        _sync_stream()
        lh, lw = self.LocalHeight(), self.LocalWidth()
        out = np.zeros((lh, lw), dtype=self.dtype, order="F")
        if lh and lw:
            _check(self._fn("ElDistMatrixLocalToHost")(self._h, out.ctypes.data_as(C.c_void_p), max(lh, 1)),
                   "ElDistMatrixLocalToHost")
        return out

    def LocalFromHost(self, a: np.ndarray):
        _sync_stream()
        lh, lw = self.LocalHeight(), self.LocalWidth()
        a = np.asfortranarray(a, dtype=self.dtype)
        if a.shape != (lh, lw):
            raise LogicError(f"local shape {a.shape} != {(lh, lw)}")
        if lh and lw:
            _check(self._fn("ElDistMatrixLocalFromHost")(self._h, a.ctypes.data_as(C.c_void_p), max(lh, 1)),
                   "ElDistMatrixLocalFromHost")
        return self

    def FromGlobal(self, G: np.ndarray):
        """Every rank passes the same global matrix; each keeps its element-cyclic piece."""
        h, w = G.shape
        if (self.Height(), self.Width()) != (h, w):
            self.Resize(h, w)
        rows = np.arange(self.ColShift(), h, self.ColStride())
        cols = np.arange(self.RowShift(), w, self.RowStride())
        return self.LocalFromHost(G[np.ix_(rows, cols)])

    def ToGlobal(self) -> np.ndarray:
        """Gather to every rank (through a [*,*] copy) and return the global matrix."""
        if self.grid.Size() == 1 or (self.colDist == STAR and self.rowDist == STAR):
            if self.colDist == STAR and self.rowDist == STAR:
                return self.LocalToHost()
        S = DistMatrix(self.dtype, STAR, STAR, self.grid)
        Copy(self, S)
        return S.LocalToHost()

    def HashFill(self, kind: int, seed: int, diag: float = 0.0):
        _sync_stream()
        _check(self._fn("ElDistMatrixHashFill")(self._h, int(kind), C.c_uint64(seed), C.c_double(diag)),
               "ElDistMatrixHashFill")
        return self


_default_grid = None


def default_grid() -> Grid:
    global _default_grid
    if _default_grid is None:
        _default_grid = Grid()
    return _default_grid


def _same(*ms):
    dt = ms[0].dtype
    for m in ms:
        if m.dtype != dt:
            raise LogicError("mixed dtypes")
    return ms[0]


# ---- level 1 ----
def Copy(A: DistMatrix, B: DistMatrix):
    _sync_stream(); _check(_same(A, B)._fn("ElCopyDist")(A._h, B._h), "ElCopyDist")


def Transpose(A: DistMatrix, B: DistMatrix, conjugate=False):
    _sync_stream()
    _check(_same(A, B)._fn("ElAdjointDist" if conjugate else "ElTransposeDist")(A._h, B._h), "ElTransposeDist")


def Adjoint(A, B):
    Transpose(A, B, True)


def Axpy(alpha, X: DistMatrix, Y: DistMatrix):
    _sync_stream(); _check(_same(X, Y)._fn("ElAxpyDist")(_scalar(X.dtype, alpha), X._h, Y._h), "ElAxpyDist")


def AxpyContract(alpha, A: DistMatrix, B: DistMatrix):
    _sync_stream()
    _check(_same(A, B)._fn("ElAxpyContractDist")(_scalar(A.dtype, alpha), A._h, B._h), "ElAxpyContractDist")


def Contract(A: DistMatrix, B: DistMatrix):
    _sync_stream(); _check(_same(A, B)._fn("ElContractDist")(A._h, B._h), "ElContractDist")


def Scale(alpha, A: DistMatrix):
    _sync_stream(); _check(A._fn("ElScaleDist")(_scalar(A.dtype, alpha), A._h), "ElScaleDist")


def Zero(A: DistMatrix):
    _sync_stream(); _check(A._fn("ElZeroDist")(A._h), "ElZeroDist")


def ScaleTrapezoid(alpha, uplo, A: DistMatrix, offset=0):
    _sync_stream()
    _check(A._fn("ElScaleTrapezoidDist")(_scalar(A.dtype, alpha), uplo, A._h, int(offset)), "ElScaleTrapezoidDist")


def MakeTrapezoidal(uplo, A: DistMatrix, offset=0):
    _sync_stream(); _check(A._fn("ElMakeTrapezoidalDist")(uplo, A._h, int(offset)), "ElMakeTrapezoidalDist")


def FrobeniusNorm(A: DistMatrix) -> float:
    _sync_stream()
    v = C.c_float() if A.dtype in (np.float32, np.complex64) else C.c_double()
    _check(A._fn("ElFrobeniusNormDist")(A._h, C.byref(v)), "ElFrobeniusNormDist")
    return float(v.value)


def MaxNorm(A: DistMatrix) -> float:
    _sync_stream()
    v = C.c_float() if A.dtype in (np.float32, np.complex64) else C.c_double()
    _check(A._fn("ElMaxNormDist")(A._h, C.byref(v)), "ElMaxNormDist")
    return float(v.value)


# ---- level 3 / factor / solve ----
def Gemm(orientA, orientB, alpha, A: DistMatrix, B: DistMatrix, beta, Cm: DistMatrix, alg=GEMM_DEFAULT):
    """El::Gemm(orientA, orientB, alpha, A, B, beta, C, alg) (src/blas_like/level3/Gemm.cpp:90-118)."""
    _sync_stream()
    dt = _same(A, B, Cm).dtype
    _check(Cm._fn("ElGemmXDist")(orientA, orientB, _scalar(dt, alpha), A._h, B._h, _scalar(dt, beta), Cm._h, alg),
           "ElGemmXDist")


_SUF = {"float32": "s", "float64": "d", "complex64": "c", "complex128": "z"}


def GemmHost(orientA, orientB, alpha, grid: Grid, m, n, k, A, B, beta, Cm, alg=GEMM_DEFAULT):
    """El::Gemm for HOST-resident [MC,MR] local matrices: A, B, Cm are this process's column-major local matrices
    (numpy arrays in Fortran order, or torch CPU tensors holding the transpose; pinned memory for the copy overlap),
    m, n, k the global sizes of op(A) op(B).  Streams C through HBM in column bands with the host copies
    overlapped (ElGemmDistHost_*, csrc/host/stream_gemm.cpp); Cm is complete on return."""
    _sync_stream()

    def ptr_ld(x):
        if isinstance(x, np.ndarray):
            if not x.flags.f_contiguous and x.ndim == 2 and min(x.shape) > 1:
                raise ValueError("GemmHost needs column-major (Fortran-ordered) local matrices")
            return x.ctypes.data_as(C.c_void_p), max(int(x.strides[1] // x.itemsize) if x.ndim == 2 and x.shape[1] > 1 else x.shape[0], 1), x.dtype
        # torch tensor of shape (local width, ld): row j is local column j
        return C.c_void_p(x.data_ptr()), max(int(x.stride(0)), 1), np.dtype(str(x.dtype).replace("torch.", ""))

    (pa, lda, dt), (pb, ldb, _), (pc, ldc, _) = ptr_ld(A), ptr_ld(B), ptr_ld(Cm)
    fn = getattr(lib(), "ElGemmDistHost_" + _SUF[np.dtype(dt).name])
    _check(fn(orientA, orientB, _scalar(dt, alpha), grid._h, int(m), int(n), int(k), pa, lda, pb, ldb,
              _scalar(dt, beta), pc, ldc, alg), "ElGemmDistHost")


BINARY, BINARY_FLAT = 3, 4   # include/El/core/types.hpp:494-510


def ReadBinaryFlat(A: DistMatrix, height, width, filename):
    """El::read::BinaryFlat(A, height, width, filename) (src/io/Read/BinaryFlat.hpp:37-102) into device memory."""
    _sync_stream()
    _check(A._fn("ElReadBinaryFlatDist")(A._h, int(height), int(width), str(filename).encode()), "ElReadBinaryFlatDist")
    return A


def ReadBinary(A: DistMatrix, filename):
    """El::read::Binary(A, filename) (src/io/Read/Binary.hpp): two Int header words, then the entries."""
    _sync_stream()
    _check(A._fn("ElReadBinaryDist")(A._h, str(filename).encode()), "ElReadBinaryDist")
    return A


def Write(A: DistMatrix, basename, fmt=BINARY):
    """El::Write(A, basename, format) for BINARY (basename.bin) and BINARY_FLAT (basename.dat) (src/io/Write.cpp:46-63)."""
    _sync_stream()
    _check(A._fn("ElWriteDist")(A._h, str(basename).encode(), int(fmt)), "ElWriteDist")


def AxpyTrapezoid(uplo, alpha, X: DistMatrix, Y: DistMatrix, offset=0):
    """El::AxpyTrapezoid(uplo, alpha, X, Y, offset) (include/El/blas_like/level1/AxpyTrapezoid.hpp:128-160)."""
    _sync_stream()
    _check(_same(X, Y)._fn("ElAxpyTrapezoidDist")(uplo, _scalar(X.dtype, alpha), X._h, Y._h, int(offset)), "ElAxpyTrapezoidDist")


def Syrk(uplo, orient, alpha, A: DistMatrix, beta, Cm: DistMatrix):
    _sync_stream()
    dt = _same(A, Cm).dtype
    _check(Cm._fn("ElSyrkDist")(uplo, orient, _scalar(dt, alpha), A._h, _scalar(dt, beta), Cm._h), "ElSyrkDist")


def Herk(uplo, orient, alpha, A: DistMatrix, beta, Cm: DistMatrix):
    _sync_stream()
    dt = _same(A, Cm).dtype
    _check(Cm._fn("ElHerkDist")(uplo, orient, _real(dt, alpha), A._h, _real(dt, beta), Cm._h), "ElHerkDist")


def Trrk(uplo, orientA, orientB, alpha, A: DistMatrix, B: DistMatrix, beta, Cm: DistMatrix):
    _sync_stream()
    dt = _same(A, B, Cm).dtype
    _check(Cm._fn("ElTrrkDist")(uplo, orientA, orientB, _scalar(dt, alpha), A._h, B._h, _scalar(dt, beta), Cm._h),
           "ElTrrkDist")


def Syr2k(uplo, orient, alpha, A: DistMatrix, B: DistMatrix, beta, Cm: DistMatrix):
    """El::Syr2k (src/blas_like/level3/Syr2k.cpp): C_tri := alpha (op(A) op(B)^T + op(B) op(A)^T) + beta C_tri."""
    _sync_stream()
    dt = _same(A, B, Cm).dtype
    _check(Cm._fn("ElSyr2kDist")(uplo, orient, _scalar(dt, alpha), A._h, B._h, _scalar(dt, beta), Cm._h), "ElSyr2kDist")


def Her2k(uplo, orient, alpha, A: DistMatrix, B: DistMatrix, beta, Cm: DistMatrix):
    """El::Her2k: alpha op(A) op(B)^H + conj(alpha) op(B) op(A)^H + beta C on the triangle (complex types; beta real)."""
    _sync_stream()
    dt = _same(A, B, Cm).dtype
    if dt.kind != "c":
        return Syr2k(uplo, orient, alpha, A, B, beta, Cm)
    _check(Cm._fn("ElHer2kDist")(uplo, orient, _scalar(dt, alpha), A._h, B._h, _real(dt, beta), Cm._h), "ElHer2kDist")


def Symm(side, uplo, alpha, A: DistMatrix, B: DistMatrix, beta, Cm: DistMatrix):
    """El::Symm (src/blas_like/level3/Symm.cpp): C := alpha A B + beta C (LEFT) / alpha B A + beta C, A = A^T from `uplo`."""
    _sync_stream()
    dt = _same(A, B, Cm).dtype
    _check(Cm._fn("ElSymmDist")(side, uplo, _scalar(dt, alpha), A._h, B._h, _scalar(dt, beta), Cm._h), "ElSymmDist")


def Hemm(side, uplo, alpha, A: DistMatrix, B: DistMatrix, beta, Cm: DistMatrix):
    _sync_stream()
    dt = _same(A, B, Cm).dtype
    if dt.kind != "c":
        return Symm(side, uplo, alpha, A, B, beta, Cm)
    _check(Cm._fn("ElHemmDist")(side, uplo, _scalar(dt, alpha), A._h, B._h, _scalar(dt, beta), Cm._h), "ElHemmDist")


def Trmm(side, uplo, orient, diag, alpha, A: DistMatrix, B: DistMatrix):
    """El::Trmm (src/blas_like/level3/Trmm.cpp): B := alpha op(tri(A)) B (LEFT) / alpha B op(tri(A))."""
    _sync_stream()
    dt = _same(A, B).dtype
    _check(B._fn("ElTrmmDist")(side, uplo, orient, diag, _scalar(dt, alpha), A._h, B._h), "ElTrmmDist")


TRSM_DEFAULT, TRSM_LARGE, TRSM_MEDIUM, TRSM_SMALL = 0, 1, 2, 3


def TwoSidedTrsm(uplo, diag, A: DistMatrix, B: DistMatrix):
    """El::TwoSidedTrsm (src/blas_like/level3/TwoSidedTrsm.cpp): A := inv(L) A inv(L)^H / inv(U)^H A inv(U)."""
    _sync_stream()
    _check(_same(A, B)._fn("ElTwoSidedTrsmDist")(uplo, diag, A._h, B._h), "ElTwoSidedTrsmDist")


def TwoSidedTrmm(uplo, diag, A: DistMatrix, B: DistMatrix):
    """El::TwoSidedTrmm (src/blas_like/level3/TwoSidedTrmm.cpp): A := L^H A L / U A U^H."""
    _sync_stream()
    _check(_same(A, B)._fn("ElTwoSidedTrmmDist")(uplo, diag, A._h, B._h), "ElTwoSidedTrmmDist")


def Trr2k(uplo, orientA, orientB, orientC, orientD, alpha, A, B, beta, Cm, D, gamma, E):
    """El::Trr2k (src/blas_like/level3/Trr2k.cpp:34-...): E_tri := alpha op(A) op(B) + beta op(C) op(D) + gamma E_tri."""
    _sync_stream()
    dt = _same(A, B, Cm, D, E).dtype
    _check(E._fn("ElTrr2kDist")(uplo, orientA, orientB, orientC, orientD, _scalar(dt, alpha), A._h, B._h,
                                _scalar(dt, beta), Cm._h, D._h, _scalar(dt, gamma), E._h), "ElTrr2kDist")


def Trsm(side, uplo, orient, diag, alpha, A: DistMatrix, B: DistMatrix, checkIfSingular=False, alg=TRSM_DEFAULT):
    """El::Trsm(side, uplo, orientation, diag, alpha, A, B, checkIfSingular, alg) (src/blas_like/level3/Trsm.cpp:67-375);
    raises SingularMatrixException when checkIfSingular finds a zero diagonal entry."""
    _sync_stream()
    dt = _same(A, B).dtype
    _check(B._fn("ElTrsmXDist")(side, uplo, orient, diag, _scalar(dt, alpha), A._h, B._h, C.c_bool(checkIfSingular),
                                int(alg)), "ElTrsmXDist")


def Trsv(uplo, orient, diag, A: DistMatrix, x: DistMatrix):
    """El::Trsv(uplo, orientation, diag, A, x) (src/blas_like/level2/Trsv.cpp:47-68)."""
    _sync_stream()
    _check(_same(A, x)._fn("ElTrsvDist")(uplo, orient, diag, A._h, x._h), "ElTrsvDist")


def Cholesky(uplo, A: DistMatrix):
    """El::Cholesky(uplo, A) (src/lapack_like/factor/Cholesky.cpp:95-110); raises NonHPDMatrixException."""
    _sync_stream()
    _check(A._fn("ElCholeskyDist")(uplo, A._h), "ElCholeskyDist")


def ReverseCholesky(uplo, A: DistMatrix):
    """El::ReverseCholesky(uplo, A) (src/lapack_like/factor/Cholesky.cpp:129-137): A = L^H L (LOWER) / U U^H (UPPER)."""
    _sync_stream()
    _check(A._fn("ElReverseCholeskyDist")(uplo, A._h), "ElReverseCholeskyDist")


def CholeskyVariant2(uplo, A: DistMatrix):
    """cholesky::LowerVariant2Blocked / UpperVariant2Blocked (Cholesky/LowerVariant2.hpp:43-110): left-looking."""
    _sync_stream()
    _check(A._fn("ElCholeskyVariant2Dist")(uplo, A._h), "ElCholeskyVariant2Dist")


def CholeskySolveAfter(uplo, orient, A: DistMatrix, B: DistMatrix):
    _sync_stream()
    _check(_same(A, B)._fn("ElCholeskySolveAfterDist")(uplo, orient, A._h, B._h), "ElCholeskySolveAfterDist")


class DistPermutation:
    """El::DistPermutation (include/El/core/DistPermutation.hpp): a swap sequence whose list lives in device memory.
    Row i of P A is row Preimage(i) of A."""

    def __init__(self, grid: Grid | None = None):
        self.grid = grid if grid is not None else default_grid()
        self._h = C.c_void_p()
        _check(lib().ElDistPermutationCreate(C.byref(self._h), self.grid._h), "ElDistPermutationCreate")

    def __del__(self):
        try:
            if self._h:
                lib().ElDistPermutationDestroy(self._h)
        except Exception:
            pass

    def _geti(self, fn, *args):
        v = C.c_int()
        _check(getattr(lib(), fn)(self._h, *args, C.byref(v)), fn)
        return v.value

    def _getb(self, fn):
        v = C.c_bool()
        _check(getattr(lib(), fn)(self._h, C.byref(v)), fn)
        return bool(v.value)

    def Empty(self): _check(lib().ElDistPermutationEmpty(self._h), "ElDistPermutationEmpty")
    def MakeIdentity(self, size): _sync_stream(); _check(lib().ElDistPermutationMakeIdentity(self._h, int(size)), "MakeIdentity")
    def ReserveSwaps(self, n): _sync_stream(); _check(lib().ElDistPermutationReserveSwaps(self._h, int(n)), "ReserveSwaps")
    def Swap(self, origin, dest): _sync_stream(); _check(lib().ElDistPermutationSwap(self._h, int(origin), int(dest)), "Swap")

    def SwapSequence(self, P: "DistPermutation", offset=0):
        _sync_stream()
        _check(lib().ElDistPermutationSwapSequence(self._h, P._h, int(offset)), "SwapSequence")

    def Height(self): return self._geti("ElDistPermutationHeight")
    def Width(self): return self._geti("ElDistPermutationWidth")
    def Parity(self): _sync_stream(); return self._getb("ElDistPermutationParity")
    def IsSwapSequence(self): return self._getb("ElDistPermutationIsSwapSequence")
    def IsImplicitSwapSequence(self): return self._getb("ElDistPermutationIsImplicitSwapSequence")
    def Image(self, origin): _sync_stream(); return self._geti("ElDistPermutationImage", int(origin))
    def Preimage(self, dest): _sync_stream(); return self._geti("ElDistPermutationPreimage", int(dest))

    def Preimages(self) -> np.ndarray:
        _sync_stream()
        n = self.Height()
        out = np.zeros(max(n, 1), dtype=np.int32)
        _check(lib().ElDistPermutationPreimages(self._h, out.ctypes.data_as(C.c_void_p)), "ElDistPermutationPreimages")
        return out[:n].astype(np.int64)

    def _apply(self, name, A: DistMatrix, offset):
        _sync_stream()
        _check(A._fn(name)(self._h, A._h, int(offset)), name)

    def PermuteRows(self, A, offset=0): self._apply("ElDistPermutationPermuteRowsDist", A, offset)
    def InversePermuteRows(self, A, offset=0): self._apply("ElDistPermutationInversePermuteRowsDist", A, offset)
    def PermuteCols(self, A, offset=0): self._apply("ElDistPermutationPermuteColsDist", A, offset)
    def InversePermuteCols(self, A, offset=0): self._apply("ElDistPermutationInversePermuteColsDist", A, offset)


def LU(A: DistMatrix, P: DistPermutation | None = None):
    """El::LU(A) without pivoting / El::LU(A, P) with partial pivoting (src/lapack_like/factor/LU.cpp:21-220)."""
    _sync_stream()
    if P is None:
        _check(A._fn("ElLUDist")(A._h), "ElLUDist")
    else:
        _check(A._fn("ElLUPartialPivDist")(A._h, P._h), "ElLUPartialPivDist")


def LUSolveAfter(orient, A: DistMatrix, B: DistMatrix, P: DistPermutation | None = None):
    """lu::SolveAfter(orientation, A, [P,] B) (LU/SolveAfter.hpp)."""
    _sync_stream()
    if P is None:
        _check(_same(A, B)._fn("ElSolveAfterLUDist")(orient, A._h, B._h), "ElSolveAfterLUDist")
    else:
        _check(_same(A, B)._fn("ElSolveAfterLUPartialPivDist")(orient, A._h, P._h, B._h), "ElSolveAfterLUPartialPivDist")


def CholeskyPiv(uplo, A: DistMatrix, P: DistPermutation):
    """El::Cholesky(uplo, A, P) (src/lapack_like/factor/Cholesky.cpp:112-121): diagonally pivoted, P A P^T = L L^H."""
    _sync_stream()
    _check(A._fn("ElCholeskyPivDist")(uplo, A._h, P._h), "ElCholeskyPivDist")


def CholeskyPivSolveAfter(uplo, orient, A: DistMatrix, P: DistPermutation, B: DistMatrix):
    """cholesky::SolveAfter(uplo, orientation, A, P, B) (Cholesky/SolveAfter.hpp:108-141)."""
    _sync_stream()
    _check(_same(A, B)._fn("ElSolveAfterCholeskyPivDist")(uplo, orient, A._h, P._h, B._h), "ElSolveAfterCholeskyPivDist")


def CholeskyMod(uplo, T: DistMatrix, alpha: float, V: DistMatrix):
    """El::CholeskyMod(uplo, T, alpha, V) (src/lapack_like/factor/Cholesky.cpp:143-173): T becomes the factor of
    T T^H + alpha V V^H (LOWER) / T^H T + alpha V V^H (UPPER); V is overwritten with workspace."""
    _sync_stream()
    _check(_same(T, V)._fn("ElCholeskyModDist")(uplo, T._h, _real(T.dtype, alpha), V._h), "ElCholeskyModDist")


def LinearSolve(A: DistMatrix, B: DistMatrix):
    """El::LinearSolve(A, B) (src/lapack_like/solve/Linear.cpp): B := inv(A) B, A unchanged."""
    _sync_stream()
    _check(_same(A, B)._fn("ElLinearSolveDist")(A._h, B._h), "ElLinearSolveDist")


def HPDSolve(uplo, orient, A: DistMatrix, B: DistMatrix):
    """El::HPDSolve(uplo, orientation, A, B) (src/lapack_like/solve/HPD.cpp:59-69)."""
    _sync_stream()
    _check(_same(A, B)._fn("ElHPDSolveDist")(uplo, orient, A._h, B._h), "ElHPDSolveDist")
