"""GPU parity of the leaf kernels (called through the C-ABI of include/elb200_blas.h) against
numpy on the same seeded inputs.  Tolerances are the north_star residuals:
  GEMM/TRRK : ||C - C_ref||_F <= 4 * k * eps * ||A||_F ||B||_F   (+ bit-exact masks / untouched padding)
  TRSM      : ||op(A) X - alpha B||_F <= 50 * n * eps * ||A||_F ||X||_F
  POTRF     : ||A - L L^H||_F <= 10 * n * eps * ||A||_F, other triangle bit-identical."""
import numpy as np
import pytest

from oracle import elemental_oracle as O

pytestmark = pytest.mark.gpu

DTYPES = [np.float64, np.complex128, np.float32, np.complex64]


def _op(X, t):
    return X if t == "N" else (X.T if t == "T" else X.conj().T)


@pytest.mark.parametrize("dt", DTYPES)
def test_gemm_all_orientations_and_ragged_sizes(dt):
    import gpuutil as G
    rng = np.random.default_rng(0)
    alpha, beta = (3.0, 4.0) if np.dtype(dt).kind != "c" else (3.0 - 1.0j, 4.0 + 0.5j)
    for (m, n, k) in [(128, 128, 16), (1, 1, 1), (7, 5, 3), (130, 257, 45), (257, 129, 200), (64, 300, 0)]:
        for ta in "NTC":
            for tb in "NTC":
                A = G.rand(rng, *((m, k) if ta == "N" else (k, m)), dt)
                B = G.rand(rng, *((k, n) if tb == "N" else (n, k)), dt)
                C0 = G.rand(rng, m, n, dt)
                dA = G.DevMat(A, A.shape[0] + 3, offset=1); dB = G.DevMat(B, B.shape[0] + 1); dC = G.DevMat(C0, m + 5, offset=1)
                G.gemm(ta, tb, alpha, dA, dB, beta, dC, k)
                ref = alpha * (_op(A, ta) @ _op(B, tb)) + beta * C0 if k else beta * C0
                got = dC.get()
                tol = 4 * max(k, 1) * G.eps(dt) * max(np.linalg.norm(A) * np.linalg.norm(B), 1) + 8 * G.eps(dt) * np.linalg.norm(C0) * abs(beta)
                assert np.linalg.norm(got - ref) <= tol, (dt, ta, tb, m, n, k)
                assert dC.padding_untouched()


@pytest.mark.parametrize("dt", DTYPES)
def test_trrk_global_staircase_mask(dt):
    import gpuutil as G
    rng = np.random.default_rng(1)
    for (m, n, k, rs, rst, cs, cst) in [(300, 300, 40, 0, 1, 0, 1), (257, 131, 33, 1, 2, 3, 4), (131, 257, 64, 0, 2, 1, 4),
                                       (129, 129, 16, 1, 3, 0, 2)]:
        for uplo in "LU":
            for ta, tb in (("T", "N"), ("N", "C"), ("N", "N"), ("C", "T")):
                A = G.rand(rng, *((m, k) if ta == "N" else (k, m)), dt)
                B = G.rand(rng, *((k, n) if tb == "N" else (n, k)), dt)
                C0 = G.rand(rng, m, n, dt)
                dA, dB, dC = G.DevMat(A), G.DevMat(B, B.shape[0] + 2), G.DevMat(C0, m + 3)
                G.trrk(uplo, ta, tb, -1.0, dA, dB, 1.0, dC, k, rs, rst, cs, cst)
                full = -(_op(A, ta) @ _op(B, tb)) + C0
                gi = rs + rst * np.arange(m)[:, None]; gj = cs + cst * np.arange(n)[None, :]
                mask = gi >= gj if uplo == "L" else gi <= gj
                got = dC.get()
                # outside the triangle: bit-identical to the input
                assert np.array_equal(got[~mask], C0[~mask])
                tol = 4 * k * G.eps(dt) * np.linalg.norm(A) * np.linalg.norm(B)
                assert np.linalg.norm((got - full)[mask]) <= tol
                assert dC.padding_untouched()


@pytest.mark.parametrize("dt", DTYPES)
def test_herk_syrk(dt):
    import gpuutil as G
    rng = np.random.default_rng(2)
    n, k = 150, 70
    for uplo in "LU":
        for tr in ("N", "C" if np.dtype(dt).kind == "c" else "T"):
            A = G.rand(rng, *((n, k) if tr == "N" else (k, n)), dt)
            C0 = G.rand(rng, n, n, dt)
            dA, dC = G.DevMat(A), G.DevMat(C0, n + 1)
            G.herk(uplo, tr, -1.0, dA, 1.0, dC, k)
            ref = O.blas_herk(uplo, tr, -1.0, A, 1.0, C0.copy())
            got = dC.get()
            assert np.linalg.norm(got - ref) <= 4 * k * G.eps(dt) * np.linalg.norm(A) ** 2


@pytest.mark.parametrize("dt", DTYPES)
def test_trsm_all_variants(dt):
    import gpuutil as G
    rng = np.random.default_rng(3)
    for (m, n) in [(100, 37), (33, 64), (1, 5), (70, 1), (256, 130)]:
        for side in "LR":
            na = m if side == "L" else n
            # strict triangle scaled by 1/na so that the unit-diagonal variants are well conditioned too (with
            # O(1) off-diagonals they are exponentially ill-conditioned and ||X|| overflows in float)
            A = G.rand(rng, na, na, dt) / max(na, 1) + 2 * np.eye(na, dtype=dt)
            for uplo in "LU":
                for tr in "NTC":
                    for diag in "NU":
                        B0 = G.rand(rng, m, n, dt)
                        alpha = 2.0
                        dA, dB = G.DevMat(A, na + 1), G.DevMat(B0, m + 2)
                        G.trsm(side, uplo, tr, diag, alpha, dA, dB)
                        X = dB.get()
                        T = O._tri(A, uplo, diag)
                        lhs = _op(T, tr) @ X if side == "L" else X @ _op(T, tr)
                        res = np.linalg.norm(lhs - alpha * B0)
                        assert np.all(np.isfinite(X))
                        assert res <= 50 * na * G.eps(dt) * np.linalg.norm(T) * max(np.linalg.norm(X), 1e-30), \
                            (dt, side, uplo, tr, diag, m, n, res)
                        assert dB.padding_untouched()


@pytest.mark.parametrize("dt", DTYPES)
def test_potrf_blocks(dt):
    import gpuutil as G
    for n in (1, 5, 32, 33, 100, 128, 256, 300):
        A = O.fill(1, n, n, 11, diag=float(n), dtype=dt)
        for uplo in "LU":
            dA = G.DevMat(A, n + 1)
            info = G.potrf(uplo, dA)
            assert info == 0
            F = dA.get()
            other = np.triu(np.ones((n, n), bool), 1) if uplo == "L" else np.tril(np.ones((n, n), bool), -1)
            assert np.array_equal(F[other], A[other]), "the other triangle must not be touched"
            assert O.cholesky_residual(uplo, F, A) <= 10
            # against the unblocked reference restatement
            Fo = O.cholesky_unblocked(uplo, A.copy(order="F"))
            tri = ~other
            assert np.linalg.norm((F - Fo)[tri]) <= 50 * n * G.eps(dt) * np.linalg.norm(Fo)


def test_potrf_reports_first_bad_pivot():
    import gpuutil as G
    n = 96
    A = O.fill(1, n, n, 3, diag=float(n))
    A[40, 40] = -5.0
    for uplo in "LU":
        info = G.potrf(uplo, G.DevMat(A))
        assert info == 41
    A = O.fill(1, 64, 64, 3, diag=64.0)
    A[0, 0] = float("nan")
    assert G.potrf("L", G.DevMat(A)) == 1


def test_potrf_large_block_host_blocked_path():
    import gpuutil as G
    n = 1100  # exceeds one CTA's shared memory -> blocked sweep over the same leaf
    A = O.fill(1, n, n, 5, diag=float(n))
    dA = G.DevMat(A)
    assert G.potrf("L", dA) == 0
    assert O.cholesky_residual("L", dA.get(), A) <= 10


def test_fortran_abi_dgemm_on_current_stream():
    import ctypes as C
    import torch
    import gpuutil as G
    from elemental_b200._lib import lib
    rng = np.random.default_rng(4)
    m, n, k = 65, 33, 17
    A, B, C0 = G.rand(rng, m, k, np.float64), G.rand(rng, k, n, np.float64), G.rand(rng, m, n, np.float64)
    dA, dB, dC = G.DevMat(A), G.DevMat(B), G.DevMat(C0)
    L = lib()
    L.elb200_set_stream(G.stream())
    i = lambda v: C.byref(C.c_int(v))
    d = lambda v: C.byref(C.c_double(v))
    L.dgemm_(C.c_char_p(b"N"), C.c_char_p(b"N"), i(m), i(n), i(k), d(3.0), dA.ptr, i(dA.ld), dB.ptr, i(dB.ld), d(4.0),
             dC.ptr, i(dC.ld))
    torch.cuda.synchronize()
    assert np.linalg.norm(dC.get() - (3 * A @ B + 4 * C0)) <= 1e-12 * k


def _even(x):
    return x + (x & 1)


def test_dgemm_tma_kernel_all_orientations_ragged():
    """The persistent TMA kernel (16-byte aligned operands, even ld): every orientation, ragged
    sizes around the 128x64 tile and the 16-deep k-stage, many tiles per CTA (persistence), k
    long enough to wrap the 4-stage ring several times."""
    import gpuutil as G
    from elemental_b200._lib import lib
    L = lib()
    rng = np.random.default_rng(7)
    dt = np.float64
    shapes = [(128, 64, 16), (1, 1, 1), (7, 5, 3), (130, 257, 45), (257, 129, 200), (64, 300, 0), (1000, 900, 333),
              (384, 2112, 130), (4100, 3100, 70)]
    for (m, n, k) in shapes:
        for ta in "NT":
            for tb in "NT":
                A = G.rand(rng, *((m, k) if ta == "N" else (k, m)), dt)
                B = G.rand(rng, *((k, n) if tb == "N" else (n, k)), dt)
                C0 = G.rand(rng, m, n, dt)
                dA = G.DevMat(A, _even(A.shape[0] + 2)); dB = G.DevMat(B, _even(B.shape[0] + 4), offset=2)
                dC = G.DevMat(C0, m + 5, offset=1)
                for alpha, beta in ((3.0, 4.0), (1.0, 0.0), (-2.5, 1.0)):
                    dC.t.copy_(__import__("torch").from_numpy(dC.host0))
                    G.gemm(ta, tb, alpha, dA, dB, beta, dC, k)
                    if k > 0:
                        assert L.elb200_dgemm_last_kernel() == 2, "TMA kernel was not selected"
                    ref = alpha * (_op(A, ta) @ _op(B, tb)) + beta * C0 if k else beta * C0
                    got = dC.get()
                    tol = 4 * max(k, 1) * G.eps(dt) * max(np.linalg.norm(A) * np.linalg.norm(B), 1) + 8 * G.eps(dt) * np.linalg.norm(C0) * abs(beta)
                    assert np.linalg.norm(got - ref) <= tol, (ta, tb, m, n, k, alpha, beta)
                    assert dC.padding_untouched()


def test_dtrrk_tma_kernel_staircase():
    import gpuutil as G
    from elemental_b200._lib import lib
    L = lib()
    rng = np.random.default_rng(8)
    dt = np.float64
    for (m, n, k, rs, rst, cs, cst) in [(300, 300, 40, 0, 1, 0, 1), (257, 131, 33, 1, 2, 3, 4), (131, 257, 64, 0, 2, 1, 4),
                                       (1029, 517, 256, 1, 2, 0, 4), (700, 700, 128, 0, 1, 0, 1)]:
        for uplo in "LU":
            for ta, tb in (("T", "N"), ("N", "T"), ("N", "N"), ("T", "T")):
                A = G.rand(rng, *((m, k) if ta == "N" else (k, m)), dt)
                B = G.rand(rng, *((k, n) if tb == "N" else (n, k)), dt)
                C0 = G.rand(rng, m, n, dt)
                dA, dB, dC = G.DevMat(A, _even(A.shape[0])), G.DevMat(B, _even(B.shape[0] + 2)), G.DevMat(C0, m + 3)
                G.trrk(uplo, ta, tb, -1.0, dA, dB, 1.0, dC, k, rs, rst, cs, cst)
                assert L.elb200_dgemm_last_kernel() == 2
                full = -(_op(A, ta) @ _op(B, tb)) + C0
                gi = rs + rst * np.arange(m)[:, None]; gj = cs + cst * np.arange(n)[None, :]
                mask = gi >= gj if uplo == "L" else gi <= gj
                got = dC.get()
                assert np.array_equal(got[~mask], C0[~mask])
                tol = 4 * k * G.eps(dt) * np.linalg.norm(A) * np.linalg.norm(B)
                assert np.linalg.norm((got - full)[mask]) <= tol
                assert dC.padding_untouched()


def test_dgemm_kernels_agree():
    """cp.async and TMA kernels accumulate each C entry over k ascending in steps of 4 inside one
    accumulator; they must agree to rounding level (a few ulp of the accumulated magnitude)."""
    import gpuutil as G
    from elemental_b200._lib import lib
    L = lib()
    rng = np.random.default_rng(9)
    m, n, k = 515, 390, 777
    A = G.rand(rng, m, k, np.float64); B = G.rand(rng, k, n, np.float64); C0 = G.rand(rng, m, n, np.float64)
    outs = []
    for cfg in (2, 3):
        L.elb200_dgemm_set_config(cfg)
        dA, dB, dC = G.DevMat(A, _even(m)), G.DevMat(B, _even(k)), G.DevMat(C0, m)
        G.gemm("N", "N", 1.5, dA, dB, -0.5, dC, k)
        outs.append(dC.get())
    L.elb200_dgemm_set_config(0)
    scale = np.abs(A) @ np.abs(B) * 1.5 + np.abs(C0) * 0.5
    assert np.all(np.abs(outs[0] - outs[1]) <= 8 * np.finfo(np.float64).eps * scale)


def test_dgemm_l2_reduction_epilogue_is_bit_identical():
    """beta == 1 uses red.global.add.f64 (C += alpha*acc at L2); it must equal the load-add-store
    epilogue bit for bit (one add per entry, one writer per entry)."""
    import gpuutil as G
    from elemental_b200._lib import lib
    L = lib()
    rng = np.random.default_rng(10)
    m, n, k = 1100, 700, 150
    A = G.rand(rng, m, k, np.float64); B = G.rand(rng, k, n, np.float64); C0 = G.rand(rng, m, n, np.float64)
    outs = []
    for flags in (0, 16):
        L.elb200_dgemm_set_debug_flags(flags)
        dA, dB, dC = G.DevMat(A, _even(m)), G.DevMat(B, _even(k)), G.DevMat(C0, m + 1)
        G.gemm("N", "N", -1.0, dA, dB, 1.0, dC, k)
        assert L.elb200_dgemm_last_kernel() == 2
        outs.append(dC.get())
        for uplo in "LU":
            dT = G.DevMat(C0[:, :n], m + 1)
            G.trrk(uplo, "N", "N", -1.0, dA, dB, 1.0, dT, k, 1, 2, 0, 4)
            outs.append(dT.get())
    L.elb200_dgemm_set_debug_flags(0)
    for a, b in zip(outs[:3], outs[3:]):
        assert np.array_equal(a, b)


def test_sgemm_3xtf32_tcgen05_all_orientations():
    """elb200_sgemm_3xtf32 (tcgen05 kind::tf32, 3-pass operand split, k-chunk promotion) against an FP64
    product.  Stated tolerance: ||C - C_ref||_F <= 8 * 2^-22 * |alpha| ||A||_F ||B||_F + 2 eps32 |beta| ||C0||_F
    (every a*b term carries <= 3 * 2^-22 from the dropped lo*lo and the truncated lo; the fp32 accumulation
    adds the usual random-walk term) -- the same order as the exact-FFMA kernel, checked beside it."""
    import gpuutil as G
    rng = np.random.default_rng(7)
    dt = np.float32
    for (m, n, k, alpha, beta) in [(128, 128, 32, 1.0, 0.0), (256, 384, 96, 3.0, 4.0), (100, 60, 40, 3.0, 4.0),
                                   (1000, 900, 1000, 1.0, 1.0), (130, 257, 4100, -1.0, 1.0), (257, 129, 7, 2.0, 0.0)]:
        for ta in "NT":
            for tb in "NT":
                A = G.rand(rng, *((m, k) if ta == "N" else (k, m)), dt)
                B = G.rand(rng, *((k, n) if tb == "N" else (n, k)), dt)
                C0 = G.rand(rng, m, n, dt)
                # TMA needs 16-byte aligned bases and leading dimensions that are multiples of 4
                lda = (A.shape[0] + 3) // 4 * 4 + 4; ldb = (B.shape[0] + 3) // 4 * 4
                dA, dB, dC = G.DevMat(A, lda), G.DevMat(B, ldb), G.DevMat(C0, m + 5, offset=1)
                G.gemm(ta, tb, alpha, dA, dB, beta, dC, k, fn="elb200_sgemm_3xtf32")
                ref = alpha * (_op(A, ta).astype(np.float64) @ _op(B, tb).astype(np.float64)) + beta * C0.astype(np.float64)
                got = dC.get().astype(np.float64)
                tol = 8 * 2.0 ** -22 * abs(alpha) * np.linalg.norm(A) * np.linalg.norm(B) + 2 * G.eps(dt) * abs(beta) * np.linalg.norm(C0)
                assert np.linalg.norm(got - ref) <= tol, (ta, tb, m, n, k, np.linalg.norm(got - ref) / tol)
                assert dC.padding_untouched()


def test_sgemm_3xtf32_rejects_unaligned_operands_loudly():
    import gpuutil as G
    rng = np.random.default_rng(8)
    A, B, C0 = (G.rand(rng, 64, 64, np.float32) for _ in range(3))
    dA, dB, dC = G.DevMat(A, 65), G.DevMat(B, 64), G.DevMat(C0, 64)   # lda % 4 != 0
    with pytest.raises(Exception):
        G.gemm("N", "N", 1.0, dA, dB, 0.0, dC, 64, fn="elb200_sgemm_3xtf32")


def test_trsm_alpha_zero_does_not_reference_a():
    """alpha == 0: B := 0 without touching A (the BLAS definition), even when A holds NaN."""
    import gpuutil as G
    A = np.full((40, 40), np.nan)
    B0 = np.asfortranarray(np.random.default_rng(1).uniform(-1, 1, (40, 17)))
    B0[3, 4] = np.inf
    dA, dB = G.DevMat(A), G.DevMat(B0, 42)
    G.trsm("L", "L", "N", "N", 0.0, dA, dB)
    assert np.array_equal(dB.get(), np.zeros_like(B0)) and dB.padding_untouched()


def test_fortran_abi_level3_on_current_stream():
    """The Fortran symbols the reference binds besides ?gemm_ (src/core/imports/blas/Trsm.hpp:12-33, Syrk.hpp:12-50):
    dtrsm_, ztrsm_, dsyrk_, ssyrk_, zherk_, cherk_, zsyrk_ with by-reference scalars and device arrays."""
    import ctypes as C
    import torch
    import gpuutil as G
    from elemental_b200._lib import lib, c32, c64
    L = lib()
    L.elb200_set_stream(G.stream())
    rng = np.random.default_rng(5)
    i = lambda v: C.byref(C.c_int(v))
    ch = lambda c: C.c_char_p(c.encode())
    scal = {np.float32: lambda v: C.byref(C.c_float(v)), np.float64: lambda v: C.byref(C.c_double(v)),
            np.complex64: lambda v: C.byref(c32(v.real, v.imag)), np.complex128: lambda v: C.byref(c64(v.real, v.imag))}
    # ?trsm_
    for dt, name in ((np.float64, "dtrsm_"), (np.complex128, "ztrsm_"), (np.float32, "strsm_"), (np.complex64, "ctrsm_")):
        m, n = 70, 33
        A = G.rand(rng, m, m, dt) / m + 2 * np.eye(m, dtype=dt)
        B0 = G.rand(rng, m, n, dt)
        dA, dB = G.DevMat(A, m + 2), G.DevMat(B0, m + 1)
        alpha = dt(2.0) if dt in (np.float32, np.float64) else dt(2.0 - 1.0j)
        getattr(L, name)(ch("L"), ch("U"), ch("C"), ch("N"), i(m), i(n), scal[dt](alpha), dA.ptr, i(dA.ld), dB.ptr, i(dB.ld))
        torch.cuda.synchronize()
        T = np.triu(A)
        assert np.linalg.norm(T.conj().T @ dB.get() - alpha * B0) <= 50 * m * G.eps(dt) * np.linalg.norm(T) * np.linalg.norm(B0)
    # ?syrk_ / ?herk_
    n, k = 65, 40
    for dt, name, herm in ((np.float64, "dsyrk_", False), (np.float32, "ssyrk_", False), (np.complex128, "zsyrk_", False),
                           (np.complex128, "zherk_", True), (np.complex64, "cherk_", True)):
        A = G.rand(rng, n, k, dt)
        C0 = G.rand(rng, n, n, dt)
        if herm:
            C0 = C0 + C0.conj().T
        dA, dC = G.DevMat(A, n + 1), G.DevMat(C0, n + 3)
        rdt = np.float32 if dt in (np.float32, np.complex64) else np.float64
        if herm:
            a, b = rdt(-1.5), rdt(0.5)
            getattr(L, name)(ch("L"), ch("N"), i(n), i(k), scal[rdt](a), dA.ptr, i(dA.ld), scal[rdt](b), dC.ptr, i(dC.ld))
            P = A @ A.conj().T
        else:
            a, b = dt(-1.5), dt(0.5)
            getattr(L, name)(ch("L"), ch("N"), i(n), i(k), scal[dt](a), dA.ptr, i(dA.ld), scal[dt](b), dC.ptr, i(dC.ld))
            P = A @ A.T
        torch.cuda.synchronize()
        got = dC.get()
        low = np.tril(np.ones((n, n), bool))
        want = np.where(low, a * P + b * C0, C0)
        assert np.array_equal(got[~low], C0[~low]), name
        assert np.linalg.norm(got - want) <= 8 * k * G.eps(dt) * (np.linalg.norm(A) ** 2 + np.linalg.norm(C0)), name


@pytest.mark.parametrize("dt", DTYPES)
def test_fortran_abi_level1_level2(dt):
    """?scal_, ?axpy_, ?lacpy_, ?syr_ / ?her_ (src/core/imports/blas/Scal.hpp:12-19, Axpy.hpp:12-29, Syr.hpp:12-30,
    src/core/imports/lapack.cpp:20-31) as device kernels: bit-exact against numpy for the copies, rounding-level
    for the arithmetic; strided and negative increments as BLAS defines them."""
    import ctypes as C
    import torch
    import gpuutil as G
    from elemental_b200._lib import lib, c32, c64
    L = lib()
    L.elb200_set_stream(G.stream())
    rng = np.random.default_rng(6)
    p = G.SUF[np.dtype(dt)]
    cplx = np.dtype(dt).kind == "c"
    rdt = np.float32 if np.dtype(dt) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64
    i = lambda v: C.byref(C.c_int(v))
    ch = lambda c: C.c_char_p(c.encode())
    def sc(v):
        if np.dtype(dt) == np.float32: return C.byref(C.c_float(v))
        if np.dtype(dt) == np.float64: return C.byref(C.c_double(v))
        return C.byref((c32 if np.dtype(dt) == np.complex64 else c64)(complex(v).real, complex(v).imag))
    alpha = dt(1.5 - 0.5j) if cplx else dt(1.5)
    e = G.eps(dt)
    n = 1000
    for incx, incy in ((1, 1), (3, 2), (-2, 1)):
        x = G.rand(rng, n * abs(incx), 1, dt)[:, 0]
        y = G.rand(rng, n * abs(incy), 1, dt)[:, 0]
        tx, ty = torch.from_numpy(x.copy()).cuda(), torch.from_numpy(y.copy()).cuda()
        getattr(L, p + "axpy_")(i(n), sc(alpha), C.c_void_p(tx.data_ptr()), i(incx), C.c_void_p(ty.data_ptr()), i(incy))
        torch.cuda.synchronize()
        xs = x[::incx][:n] if incx > 0 else x[: (n - 1) * (-incx) + 1][::-1][::-incx][:n]
        want = y.copy()
        idx = np.arange(n) * incy
        want[idx] = y[idx] + alpha * xs
        assert np.linalg.norm(ty.cpu().numpy() - want) <= 4 * e * np.linalg.norm(want)
        if incx > 0:
            getattr(L, p + "scal_")(i(n), sc(alpha), C.c_void_p(tx.data_ptr()), i(incx))
            torch.cuda.synchronize()
            want = x.copy()
            want[np.arange(n) * incx] = alpha * x[np.arange(n) * incx]
            assert np.linalg.norm(tx.cpu().numpy() - want) <= 4 * e * np.linalg.norm(want)
    m, nn = 45, 37
    A = G.rand(rng, m, nn, dt)
    for uplo in "ULA":
        B0 = G.rand(rng, m, nn, dt)
        dA, dB = G.DevMat(A, m + 3), G.DevMat(B0, m + 1)
        getattr(L, p + "lacpy_")(ch(uplo), i(m), i(nn), dA.ptr, i(dA.ld), dB.ptr, i(dB.ld))
        torch.cuda.synchronize()
        ii, jj = np.indices((m, nn))
        sel = (ii <= jj) if uplo == "U" else ((ii >= jj) if uplo == "L" else np.ones((m, nn), bool))
        assert np.array_equal(dB.get(), np.where(sel, A, B0)) and dB.padding_untouched()
    nn = 130
    x = G.rand(rng, nn, 1, dt)[:, 0]
    tx = torch.from_numpy(x.copy()).cuda()
    name = {"s": "ssyr_", "d": "dsyr_", "c": "cher_", "z": "zher_"}[p]
    for uplo in "LU":
        A0 = G.rand(rng, nn, nn, dt)
        dA = G.DevMat(A0, nn + 1)
        ra = rdt(-0.75)
        ralpha = C.byref(C.c_float(ra)) if rdt == np.float32 else C.byref(C.c_double(ra))
        getattr(L, name)(ch(uplo), i(nn), ralpha, C.c_void_p(tx.data_ptr()), i(1), dA.ptr, i(dA.ld))
        torch.cuda.synchronize()
        ii, jj = np.indices((nn, nn))
        tri = (ii >= jj) if uplo == "L" else (ii <= jj)
        upd = A0 + ra * np.outer(x, x.conj() if cplx else x)
        if cplx:
            upd[np.arange(nn), np.arange(nn)] = upd[np.arange(nn), np.arange(nn)].real
        want = np.where(tri, upd, A0)
        got = dA.get()
        assert np.array_equal(got[~tri], A0[~tri]) and dA.padding_untouched()
        assert np.linalg.norm(got - want) <= 8 * e * np.linalg.norm(want)


@pytest.mark.parametrize("dt", DTYPES)
def test_trapezoid_kernels_on_the_global_staircase(dt):
    """elb200_scale_trapezoid / elb200_make_trapezoidal (ScaleTrapezoid.hpp:14-86, MakeTrapezoidal) with alpha != 1,
    nonzero offsets and [MC,MR]-style shifts / strides: bit-exact against the same predicate in numpy."""
    import ctypes as C
    import gpuutil as G
    from elemental_b200._lib import lib, check
    L = lib()
    rng = np.random.default_rng(8)
    m, n = 77, 53
    code = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.complex64): 2, np.dtype(np.complex128): 3}[np.dtype(dt)]
    alpha = dt(0.5)   # a power of two: the scaled entries are exact
    for (rs, rst, cs, cst) in ((0, 1, 0, 1), (1, 2, 3, 4), (2, 3, 0, 2)):
        for uplo in "LU":
            for off in (0, 1, -2, 5):
                A0 = G.rand(rng, m, n, dt)
                gi = rs + np.arange(m)[:, None] * rst
                gj = cs + np.arange(n)[None, :] * cst
                inside = (gj - gi <= off) if uplo == "L" else (gj - gi >= off)
                dA = G.DevMat(A0, m + 1)
                a = G.sc(dt, alpha)
                check(L.elb200_scale_trapezoid(code, C.byref(a), G.ch(uplo), G.i64(m), G.i64(n), dA.ptr, G.i64(dA.ld),
                                               G.i64(rs), G.i64(rst), G.i64(cs), G.i64(cst), G.i64(off), G.stream()), "scale")
                assert np.array_equal(dA.get(), np.where(inside, alpha * A0, A0)) and dA.padding_untouched()
                dA = G.DevMat(A0, m + 1)
                check(L.elb200_make_trapezoidal(code, G.ch(uplo), G.i64(m), G.i64(n), dA.ptr, G.i64(dA.ld),
                                                G.i64(rs), G.i64(rst), G.i64(cs), G.i64(cst), G.i64(off), G.stream()), "make")
                assert np.array_equal(dA.get(), np.where(inside, A0, 0)) and dA.padding_untouched()


def test_zgemm_real_kernel_path_agrees_with_complex_kernel():
    """Complex<double> products above the size threshold run on the real persistent kernel through the (re, im) row-pair
    identity (gemm_c64_real.cu); the dedicated complex kernel (gemm_c64.cu) is the second implementation.  Both must
    meet the GEMM tolerance against numpy, on ragged sizes, every orientation, the masked TRRK form and ZHERK's real
    diagonal, and elb200_zgemm_last_kernel must show that the intended kernel ran."""
    import gpuutil as G
    L = G.lib()
    L.elb200_zgemm_last_kernel.restype = int
    dt = np.complex128
    rng = np.random.default_rng(21)
    try:
        for (m, n, k) in [(257, 129, 200), (130, 300, 77), (512, 256, 128)]:
            for ta in "NTC":
                for tb in "NTC":
                    for alpha, beta in ((1.0, 1.0), (-1.0, 1.0), (0.5 - 2.0j, 0.25), (2.0, 0.5 + 1.0j)):
                        A = G.rand(rng, *((m, k) if ta == "N" else (k, m)), dt)
                        B = G.rand(rng, *((k, n) if tb == "N" else (n, k)), dt)
                        C0 = G.rand(rng, m, n, dt)
                        ref = alpha * (_op(A, ta) @ _op(B, tb)) + beta * C0
                        tol = 4 * k * G.eps(dt) * np.linalg.norm(A) * np.linalg.norm(B) + 8 * G.eps(dt) * np.linalg.norm(C0) * abs(beta)
                        for path, kern in ((0, 2), (1, 1)):
                            L.elb200_zgemm_set_path(path)
                            dA, dB, dC = G.DevMat(A, A.shape[0] + 3, offset=1), G.DevMat(B, B.shape[0] + 1), G.DevMat(C0, m + 5, offset=1)
                            G.gemm(ta, tb, alpha, dA, dB, beta, dC, k)
                            assert L.elb200_zgemm_last_kernel() == kern, (path, m, n, k)
                            assert np.linalg.norm(dC.get() - ref) <= tol, (path, ta, tb, m, n, k, alpha, beta)
                            assert dC.padding_untouched()
        # masked rank-k update on a global staircase and the Hermitian update (real diagonal), both paths
        m, n, k = 300, 260, 64
        for uplo in "LU":
            A = G.rand(rng, m, k, dt); B = G.rand(rng, k, n, dt); C0 = G.rand(rng, m, n, dt)
            gi = 1 + 2 * np.arange(m)[:, None]; gj = 3 + 4 * np.arange(n)[None, :]
            mask = gi >= gj if uplo == "L" else gi <= gj
            full = -(A @ B) + C0
            for path, kern in ((0, 2), (1, 1)):
                L.elb200_zgemm_set_path(path)
                dA, dB, dC = G.DevMat(A), G.DevMat(B, k + 2), G.DevMat(C0, m + 3)
                G.trrk(uplo, "N", "N", -1.0, dA, dB, 1.0, dC, k, 1, 2, 3, 4)
                assert L.elb200_zgemm_last_kernel() == kern
                got = dC.get()
                assert np.array_equal(got[~mask], C0[~mask])
                assert np.linalg.norm((got - full)[mask]) <= 4 * k * G.eps(dt) * np.linalg.norm(A) * np.linalg.norm(B)
            nn, kk = 300, 96
            for tr in "NC":
                A = G.rand(rng, *((nn, kk) if tr == "N" else (kk, nn)), dt)
                C0 = G.rand(rng, nn, nn, dt)
                ref = O.blas_herk(uplo, tr, -1.0, A, 1.0, C0.copy())
                for path, kern in ((0, 2), (1, 1)):
                    L.elb200_zgemm_set_path(path)
                    dA, dC = G.DevMat(A), G.DevMat(C0, nn + 1)
                    G.herk(uplo, tr, -1.0, dA, 1.0, dC, kk)
                    assert L.elb200_zgemm_last_kernel() == kern
                    got = dC.get()
                    assert np.linalg.norm(got - ref) <= 4 * kk * G.eps(dt) * np.linalg.norm(A) ** 2
                    tri = np.tril(np.ones((nn, nn), bool)) if uplo == "L" else np.triu(np.ones((nn, nn), bool))
                    assert np.array_equal(got[~tri], C0[~tri])
                    assert np.all(np.diag(got).imag == 0.0)
    finally:
        L.elb200_zgemm_set_path(0)


def test_sgemm_register_tiled_ffma_kernel_agrees_with_generic_kernel():
    """Exact-FFMA float products with 16-byte-aligned operands run on the register-tiled kernel (gemm_f32_ffma.cu); the
    generic SIMT kernel (gemm_simt.cu) takes everything else.  Both against numpy in FP64, ragged sizes, all four
    storage orientations, general alpha / beta, the masked form; unaligned operands must fall back (kernel 1)."""
    import gpuutil as G
    L = G.lib()
    L.elb200_sgemm_ffma_last_kernel.restype = int
    dt = np.float32
    rng = np.random.default_rng(22)
    try:
        for (m, n, k) in [(128, 128, 8), (257, 131, 45), (130, 260, 200), (1, 1, 1), (5, 300, 17)]:
            for ta in "NT":
                for tb in "NT":
                    for alpha, beta in ((1.0, 0.0), (-1.0, 1.0), (3.0, 4.0)):
                        A = G.rand(rng, *((m, k) if ta == "N" else (k, m)), dt)
                        B = G.rand(rng, *((k, n) if tb == "N" else (n, k)), dt)
                        C0 = G.rand(rng, m, n, dt)
                        ref = alpha * (_op(A, ta).astype(np.float64) @ _op(B, tb).astype(np.float64)) + beta * C0
                        tol = 4 * k * G.eps(dt) * max(np.linalg.norm(A) * np.linalg.norm(B), 1) + 8 * G.eps(dt) * np.linalg.norm(C0) * abs(beta)
                        lda = (A.shape[0] + 3) // 4 * 4 + 4; ldb = (B.shape[0] + 3) // 4 * 4
                        for path, kern in ((0, 2), (1, 1)):
                            L.elb200_sgemm_set_ffma_path(path)
                            dA, dB, dC = G.DevMat(A, lda), G.DevMat(B, ldb), G.DevMat(C0, m + 5, offset=1)
                            G.gemm(ta, tb, alpha, dA, dB, beta, dC, k)
                            assert L.elb200_sgemm_ffma_last_kernel() == kern, (path, m, n, k)
                            assert np.linalg.norm(dC.get() - ref) <= tol, (path, ta, tb, m, n, k, alpha, beta)
                            assert dC.padding_untouched()
        L.elb200_sgemm_set_ffma_path(0)
        # unaligned leading dimension / base pointer: the generic kernel, silently and correctly
        A = G.rand(rng, 130, 40, dt); B = G.rand(rng, 40, 70, dt); C0 = G.rand(rng, 130, 70, dt)
        dA, dB, dC = G.DevMat(A, 133, offset=1), G.DevMat(B, 41), G.DevMat(C0, 131)
        G.gemm("N", "N", 1.0, dA, dB, 1.0, dC, 40)
        assert L.elb200_sgemm_ffma_last_kernel() == 1
        assert np.linalg.norm(dC.get() - (A.astype(np.float64) @ B + C0)) <= 4 * 40 * G.eps(dt) * np.linalg.norm(A) * np.linalg.norm(B)
        # masked form on the register-tiled kernel
        m, n, k = 260, 200, 32
        for uplo in "LU":
            A = G.rand(rng, m, k, dt); B = G.rand(rng, k, n, dt); C0 = G.rand(rng, m, n, dt)
            gi = 2 * np.arange(m)[:, None]; gj = 1 + 3 * np.arange(n)[None, :]
            mask = gi >= gj if uplo == "L" else gi <= gj
            dA, dB, dC = G.DevMat(A), G.DevMat(B), G.DevMat(C0, m + 4)
            G.trrk(uplo, "N", "N", -1.0, dA, dB, 1.0, dC, k, 0, 2, 1, 3)
            assert L.elb200_sgemm_ffma_last_kernel() == 2
            got = dC.get()
            assert np.array_equal(got[~mask], C0[~mask])
            assert np.linalg.norm((got - (C0 - A.astype(np.float64) @ B))[mask]) <= 4 * k * G.eps(dt) * np.linalg.norm(A) * np.linalg.norm(B)
    finally:
        L.elb200_sgemm_set_ffma_path(0)
