"""GPU parity of the El-level API (elemental_b200.api, through the C API of include/elb200_El.h)
against the reference itself (oracle/_ref/libElRef.so, Elemental's own sources) when it is
present, else against the numpy restatement -- same inputs (grid-independent hash fill), same
Blocksize().  These mirror the reference's own drivers: tests/blas_like/Gemm.cpp (alpha=3,
beta=4, all SUMMA variants), tests/lapack_like/Cholesky.cpp (solve check <= 100),
tests/blas_like/Trsm.cpp, tests/blas_like/Syrk.cpp, tests/core/DistMatrix.cpp.

Tolerances (BASELINE.json north_star):
   ||C - C_ref||_F / (k eps ||A||_F ||B||_F) <= 1     ||A - L L^H||_F / (n eps ||A||_F) <= 10
"""
import numpy as np
import pytest

from oracle import elemental_oracle as O
from oracle import reference_lib as R

pytestmark = pytest.mark.gpu

DT = [np.float64, np.complex128, np.float32, np.complex64]
UL = {"L": 0, "U": 1}
DG = {"N": 0, "U": 1}
ORI = {"N": 0, "T": 1, "C": 2}


def _ref_gemm(oa, ob, alpha, A, B, beta, C, nb, alg):
    if R.available():
        return R.gemm(oa, ob, alpha, A, B, beta, C, nb=nb, alg=alg)
    return O.gemm(oa, ob, alpha, A, B, beta, C, nb=nb, alg=alg)


@pytest.fixture(scope="module")
def El():
    from elemental_b200 import api
    api.Initialize()
    return api


def _dm(El, a, dist=(0, 2)):
    M = El.DistMatrix(a.dtype, dist[0], dist[1])
    M.FromGlobal(a)
    return M


@pytest.mark.parametrize("dt", DT)
def test_gemm_matches_reference_all_variants(El, dt):
    m, n, k, nb = 200, 160, 144, 32
    alpha, beta = (3.0, 4.0)
    for oa in "NTC":
        for ob in "NTC":
            A = O.fill(0, *((m, k) if oa == "N" else (k, m)), 1, dtype=dt)
            B = O.fill(0, *((k, n) if ob == "N" else (n, k)), 2, dtype=dt)
            C0 = O.fill(0, m, n, 3, dtype=dt)
            for alg in (El.GEMM_SUMMA_A, El.GEMM_SUMMA_B, El.GEMM_SUMMA_C, El.GEMM_SUMMA_DOT, El.GEMM_DEFAULT):
                El.PushBlocksizeStack(nb)
                dA, dB, dC = _dm(El, A), _dm(El, B), _dm(El, C0)
                El.Gemm(ORI[oa], ORI[ob], alpha, dA, dB, beta, dC, alg)
                El.PopBlocksizeStack()
                got = dC.ToGlobal()
                ref = _ref_gemm(oa, ob, alpha, A, B, beta, C0.copy(order="F"), nb, alg)
                assert O.gemm_residual(got, ref, k, A, B) <= 1.0, (dt, oa, ob, alg)


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_gemm_host_streamed_matches_device_gemm(El, dt, monkeypatch):
    """El.GemmHost (ElGemmDistHost_*: host-resident local matrices, C streamed through HBM in column bands, A in
    chunks of the summation index) against El.Gemm on device-resident copies of the same inputs.  alpha = -1:
    bit-identical (same rank-nb updates in the same order); general alpha: the Gemm tolerance.  The band count is
    forced to 8 for the orientation sweep (the default picks 2 bands for matrices this small); the default and a
    single band are run as well -- the result does not depend on the schedule."""
    m, n, k, nb = 210, 300, 200, 32     # 8 wanted: 7 bands of 48 columns, 7 chunks of 32 summation indices
    g = El.Grid()
    A, B, C0 = O.fill(0, m, k, 1, dtype=dt), O.fill(0, k, n, 2, dtype=dt), O.fill(0, m, n, 3, dtype=dt)
    El.PushBlocksizeStack(nb)
    dC = _dm(El, C0)
    El.Gemm(0, 0, -1.0, _dm(El, A), _dm(El, B), 1.0, dC, El.GEMM_SUMMA_C)
    want = dC.ToGlobal()
    for bands in (None, "1", "3"):
        if bands is None:
            monkeypatch.delenv("ELB200_GEMMHOST_BANDS", raising=False)
        else:
            monkeypatch.setenv("ELB200_GEMMHOST_BANDS", bands)
        hC = np.asfortranarray(C0.copy())
        El.GemmHost(0, 0, -1.0, g, m, n, k, np.asfortranarray(A), np.asfortranarray(B), 1.0, hC, El.GEMM_SUMMA_C)
        assert np.array_equal(hC, want), (dt, bands)
    El.PopBlocksizeStack()
    monkeypatch.setenv("ELB200_GEMMHOST_BANDS", "8")
    for oa in "NTC":
        for ob in "NTC":
            A = O.fill(0, *((m, k) if oa == "N" else (k, m)), 1, dtype=dt)
            B = O.fill(0, *((k, n) if ob == "N" else (n, k)), 2, dtype=dt)
            C0 = O.fill(0, m, n, 3, dtype=dt)
            for alpha, beta, alg in ((-1.0, 1.0, El.GEMM_SUMMA_C), (3.0, 4.0, El.GEMM_SUMMA_C), (-1.0, 0.5, El.GEMM_DEFAULT)):
                El.PushBlocksizeStack(nb)
                dA, dB, dC = _dm(El, A), _dm(El, B), _dm(El, C0)
                El.Gemm(ORI[oa], ORI[ob], alpha, dA, dB, beta, dC, alg)
                want = dC.ToGlobal()
                hA, hB, hC = np.asfortranarray(A), np.asfortranarray(B), np.asfortranarray(C0.copy())
                El.GemmHost(ORI[oa], ORI[ob], alpha, g, m, n, k, hA, hB, beta, hC, alg)
                El.PopBlocksizeStack()
                if alpha == -1.0 and alg == El.GEMM_SUMMA_C:
                    assert np.array_equal(hC, want), (dt, oa, ob, alpha)
                else:
                    assert O.gemm_residual(hC, want, k, A, B) <= 1.0, (dt, oa, ob, alpha, alg)


def test_gemm_cannon_nn(El):
    """GEMM_CANNON (Gemm/NN.hpp:15-89) on the square grid at hand (1x1 here; tests/mgpu_worker.py runs it on 2x2): the
    ring-shift product must agree with the reference's result; the other orientations reject the option as the
    reference does (NT.hpp / TN.hpp / TT.hpp: "Unsupported Gemm option")."""
    m, n, k = 96, 80, 64
    A, B, C0 = O.fill(0, m, k, 1), O.fill(0, k, n, 2), O.fill(0, m, n, 3)
    dC = _dm(El, C0)
    El.Gemm(El.NORMAL, El.NORMAL, 3.0, _dm(El, A), _dm(El, B), 4.0, dC, El.GEMM_CANNON)
    assert O.gemm_residual(dC.ToGlobal(), 3.0 * A @ B + 4.0 * C0, k, A, B) <= 1.0
    with pytest.raises(El.Elb200Error):
        El.Gemm(El.TRANSPOSE, El.NORMAL, 1.0, _dm(El, np.asfortranarray(A.T)), _dm(El, B), 0.0, _dm(El, C0), El.GEMM_CANNON)


def test_gemm_float_3xtf32_mode_summa_dot(El):
    """BASELINE.json configs[4] in miniature: El::Gemm float, tall-skinny k (auto-selects SUMMA_Dot, NN.hpp:305),
    exact-FFMA mode vs 3xTF32 mode, both against the FP64 product.  Tolerance (north_star):
    ||C - C_ref||_F / (k eps32 ||A||_F ||B||_F) <= 1, and the mode switch must really change the kernel."""
    from elemental_b200._lib import lib
    L = lib()
    m, n, k = 256, 192, 8192
    A, B, C0 = O.fill(0, m, k, 1, dtype=np.float32), O.fill(0, k, n, 2, dtype=np.float32), O.fill(0, m, n, 3, dtype=np.float32)
    ref = 3.0 * (A.astype(np.float64) @ B.astype(np.float64)) + 4.0 * C0
    errs = {}
    try:
        for mode, kern in ((0, 1), (1, 2)):
            L.elb200_sgemm_set_mode(mode)
            dA, dB, dC = _dm(El, A), _dm(El, B), _dm(El, C0)
            El.Gemm(El.NORMAL, El.NORMAL, 3.0, dA, dB, 4.0, dC)
            assert L.elb200_sgemm_last_kernel() == kern
            errs[mode] = O.gemm_residual(dC.ToGlobal(), ref, k, A, B)   # eps of float32
            assert errs[mode] <= 1.0, errs
    finally:
        L.elb200_sgemm_set_mode(0)


def test_gemm_dot_blocksize_does_not_change_the_result(El):
    """SUMMA_Dot forms C block by block (reference blockSizeDot = 2000, NN.hpp:233); here the block edge is sized
    for HBM.  Every entry is a full-k product, so small blocks, the reference's 2000 and one block must agree
    with the reference library and -- up to the tile the entry falls in -- with each other."""
    m, n, k = 300, 260, 2100
    A, B, C0 = O.fill(0, m, k, 1), O.fill(0, k, n, 2), O.fill(0, m, n, 3)
    ref = _ref_gemm("N", "N", 3.0, A, B, 4.0, C0.copy(order="F"), 128, El.GEMM_SUMMA_DOT)
    try:
        for bs in (128, 2000, 0):
            El.SetGemmDotBlocksize(bs)
            dA, dB, dC = _dm(El, A), _dm(El, B), _dm(El, C0)
            El.Gemm(El.NORMAL, El.NORMAL, 3.0, dA, dB, 4.0, dC, El.GEMM_SUMMA_DOT)
            assert O.gemm_residual(dC.ToGlobal(), ref, k, A, B) <= 1.0, bs
    finally:
        El.SetGemmDotBlocksize(0)


def test_gemm_config1_2048_nb128(El):
    """BASELINE.json configs[0]: Gemm NN double m=n=k=2048 nb=128 on a 1x1 Grid."""
    n = 2048
    A, B, C0 = O.fill(0, n, n, 1), O.fill(0, n, n, 2), O.fill(0, n, n, 3)
    dA = El.DistMatrix(np.float64, height=n, width=n).HashFill(0, 1)
    dB = El.DistMatrix(np.float64, height=n, width=n).HashFill(0, 2)
    dC = El.DistMatrix(np.float64, height=n, width=n).HashFill(0, 3)
    assert np.array_equal(dA.ToGlobal(), A), "device hash fill must equal the oracle generator bit for bit"
    El.PushBlocksizeStack(128)
    El.Gemm(El.NORMAL, El.NORMAL, 3.0, dA, dB, 4.0, dC, El.GEMM_SUMMA_C)
    El.PopBlocksizeStack()
    ref = _ref_gemm("N", "N", 3.0, A, B, 4.0, C0.copy(order="F"), 128, 3)
    assert O.gemm_residual(dC.ToGlobal(), ref, n, A, B) <= 1.0


def test_gemm_misaligned_and_empty(El):
    A = O.fill(0, 50, 0, 1)
    dA = El.DistMatrix(np.float64, height=50, width=0)
    dB = El.DistMatrix(np.float64, height=0, width=40)
    C0 = O.fill(0, 50, 40, 3)
    dC = _dm(El, C0)
    El.Gemm(El.NORMAL, El.NORMAL, 3.0, dA, dB, 4.0, dC)   # k == 0: C := beta C (Gemm.cpp:68-71)
    assert np.array_equal(dC.ToGlobal(), 4.0 * C0)
    with pytest.raises(El.LogicError):
        El.Gemm(El.NORMAL, El.NORMAL, 1.0, _dm(El, O.fill(0, 5, 6, 1)), _dm(El, O.fill(0, 7, 5, 1)), 0.0, _dm(El, O.fill(0, 5, 5, 1)))


@pytest.mark.parametrize("dt", DT)
@pytest.mark.parametrize("uplo", "LU")
def test_cholesky_matches_reference(El, dt, uplo):
    for n, nb in [(300, 64), (257, 96), (64, 128), (1, 32)]:
        A = O.fill(1, n, n, 5, diag=float(n), dtype=dt)
        dA = _dm(El, A)
        El.PushBlocksizeStack(nb)
        El.Cholesky(El.LOWER if uplo == "L" else El.UPPER, dA)
        El.PopBlocksizeStack()
        F = dA.ToGlobal()
        other = np.triu(np.ones((n, n), bool), 1) if uplo == "L" else np.tril(np.ones((n, n), bool), -1)
        assert np.array_equal(F[other], A[other]), "other triangle must be left untouched"
        assert O.cholesky_residual(uplo, F, A) <= 10
        ref = R.cholesky(uplo, A.copy(order="F"), nb=nb) if R.available() else O.cholesky(uplo, A.copy(order="F"), nb)
        e = np.finfo(np.dtype(dt).char.lower() if np.dtype(dt).kind == "c" else dt).eps
        assert np.linalg.norm((F - ref)[~other]) <= 20 * n * e * np.linalg.norm(ref)
        # the reference test's own criterion (tests/lapack_like/Cholesky.cpp:47-82), threshold 100
        if np.dtype(dt).itemsize >= 8 and np.dtype(dt) != np.complex64:
            assert O.cholesky_solve_check(uplo, F.astype(A.dtype), A) <= 100


def test_cholesky_non_hpd_raises(El):
    A = O.fill(1, 200, 200, 5, diag=0.0)
    dA = _dm(El, A)
    with pytest.raises(El.NonHPDMatrixException):
        El.Cholesky(El.LOWER, dA)
    B = O.fill(1, 200, 200, 5, diag=200.0)
    B[150, 150] = -1.0
    with pytest.raises(El.NonHPDMatrixException):
        El.Cholesky(El.UPPER, _dm(El, B))


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_hpdsolve_matches_reference(El, dt):
    n, nrhs, nb = 260, 70, 64
    A = O.fill(1, n, n, 7, diag=float(n), dtype=dt)
    B = O.fill(0, n, nrhs, 8, dtype=dt)
    for uplo in "LU":
        for orient in "NT":
            dA, dB = _dm(El, A), _dm(El, B)
            El.PushBlocksizeStack(nb)
            El.HPDSolve(El.LOWER if uplo == "L" else El.UPPER, ORI[orient], dA, dB)
            El.PopBlocksizeStack()
            X = dB.ToGlobal()
            assert np.array_equal(dA.ToGlobal(), A), "HPDSolve must not modify A (HPD.cpp:67 copies it)"
            ref = (R.hpd_solve(uplo, orient, A, B.copy(order="F"), nb=nb) if R.available()
                   else O.hpd_solve(uplo, orient, A, B.copy(order="F"), nb))
            Aeff = A if orient == "N" else A.T
            e = np.finfo(np.float64).eps
            assert np.linalg.norm(Aeff @ X - B) <= 10 * n * e * np.linalg.norm(A) * np.linalg.norm(X)
            assert np.linalg.norm(X - ref) <= 100 * n * e * np.linalg.norm(ref)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_trsm_dist_all_variants(El, dt):
    m, n, nb = 150, 90, 32
    for side in "LR":
        na = m if side == "L" else n
        # strict triangle scaled by 1/na: the unit-diagonal variants stay well conditioned too
        A = O.fill(0, na, na, 9, dtype=dt) / na + 2 * np.eye(na)
        for uplo in "LU":
            for tr in "NTC":
                for diag in "NU":
                    B0 = O.fill(0, m, n, 10, dtype=dt)
                    dA, dB = _dm(El, np.asfortranarray(A.astype(dt))), _dm(El, B0)
                    El.PushBlocksizeStack(nb)
                    El.Trsm(0 if side == "L" else 1, 0 if uplo == "L" else 1, ORI[tr], 0 if diag == "N" else 1, 2.0, dA, dB)
                    El.PopBlocksizeStack()
                    X = dB.ToGlobal()
                    T = O._tri(A.astype(dt), uplo, diag)
                    opT = T if tr == "N" else (T.T if tr == "T" else T.conj().T)
                    lhs = opT @ X if side == "L" else X @ opT
                    e = np.finfo(np.float64).eps
                    assert np.linalg.norm(lhs - 2.0 * B0) <= 50 * na * e * np.linalg.norm(T) * np.linalg.norm(X), (side, uplo, tr, diag)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_trsm_dist_algorithms_and_trsv(El, dt):
    """TrsmAlgorithm Large / Medium / Small (Trsm/LLN.hpp:18-175, LLT.hpp:20-252 and the upper mirrors), the default
    rule (Large iff width > 5 p, Trsm.cpp:126-128), and the width-1 dispatch to Trsv (Trsm.cpp:97-101), against the
    reference library at the same blocksize (residual-level: the reference's flops are an external BLAS)."""
    m, nb = 150, 32
    A = np.asfortranarray((O.fill(0, m, m, 9, dtype=dt) / m + 2 * np.eye(m)).astype(dt))
    e = np.finfo(np.float64).eps
    for n in (40, 4, 1):
        for uplo in "LU":
            for tr in "NTC":
                for diag in "NU":
                    B0 = O.fill(0, m, n, 10, dtype=dt)
                    ref = (R.trsm("L", uplo, tr, diag, 2.0, A, B0.copy(order="F"), nb=nb) if R.available()
                           else O.trsm("L", uplo, tr, diag, 2.0, A, B0.copy(order="F"), nb=nb))
                    for alg in (El.TRSM_DEFAULT, El.TRSM_LARGE, El.TRSM_MEDIUM, El.TRSM_SMALL):
                        dA, dB = _dm(El, A), _dm(El, B0)
                        El.PushBlocksizeStack(nb)
                        El.Trsm(0, UL[uplo], ORI[tr], DG[diag], 2.0, dA, dB, False, alg)
                        El.PopBlocksizeStack()
                        X = dB.ToGlobal()
                        assert np.linalg.norm(X - ref) <= 50 * m * e * np.linalg.norm(ref), (dt, n, uplo, tr, diag, alg)
    # El::Trsv on a column and on a row vector
    for uplo in "LU":
        for tr in "NC":
            x0 = O.fill(0, m, 1, 12, dtype=dt)
            T = O._tri(A, uplo, "N")
            opT = T if tr == "N" else T.conj().T
            want = np.linalg.solve(opT, x0)
            dA, dx = _dm(El, A), _dm(El, x0)
            El.Trsv(UL[uplo], ORI[tr], 0, dA, dx)
            assert np.linalg.norm(dx.ToGlobal() - want) <= 50 * m * e * np.linalg.norm(want)
            dr = _dm(El, np.asfortranarray(x0.T))
            El.Trsv(UL[uplo], ORI[tr], 0, dA, dr)
            assert np.linalg.norm(dr.ToGlobal().T - want) <= 50 * m * e * np.linalg.norm(want)


def test_trsm_check_if_singular_raises(El):
    """checkIfSingular: a zero on the diagonal raises SingularMatrixException (Trsm.cpp:54-60) -- from the
    distributed solve (any block) and not for a unit-diagonal solve; without the flag inf/nan come out silently."""
    m, n, nb = 100, 30, 32
    A = np.asfortranarray(O.fill(0, m, m, 9) / m + 2 * np.eye(m))
    A[70, 70] = 0.0
    B0 = O.fill(0, m, n, 10)
    for alg in (El.TRSM_LARGE, El.TRSM_MEDIUM, El.TRSM_SMALL):
        dA, dB = _dm(El, A), _dm(El, B0)
        El.PushBlocksizeStack(nb)
        with pytest.raises(El.SingularMatrixException):
            El.Trsm(0, 0, 0, 0, 1.0, dA, dB, True, alg)
        dB = _dm(El, B0)
        El.Trsm(0, 0, 0, 1, 1.0, dA, dB, True, alg)      # UNIT: the stored diagonal is never read
        assert np.all(np.isfinite(dB.ToGlobal()))
        El.PopBlocksizeStack()
    dA, dB = _dm(El, A), _dm(El, np.asfortranarray(B0.T))
    with pytest.raises(El.SingularMatrixException):
        El.Trsm(1, 0, 0, 0, 1.0, dA, dB, True)


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_syrk_herk_trrk_dist_with_beta(El, dt):
    """beta != 1 exercises ScaleTrapezoid on the local staircase (level1/ScaleTrapezoid.hpp:14-86) before the panel
    loop; the strictly-other triangle must stay bit-identical."""
    n, k, nb = 170, 90, 48
    e = np.finfo(np.dtype(dt).type(0).real.dtype).eps
    for uplo in "LU":
        for o in "NC":
            A = O.fill(0, *((n, k) if o == "N" else (k, n)), 3, dtype=dt)
            C0 = O.fill(0, n, n, 4, dtype=dt)
            mask = O._tri_mask(n, n, uplo)
            El.SetBlocksize(nb)
            dC = _dm(El, C0)
            El.Herk(UL[uplo], ORI[o], -0.75, _dm(El, A), 0.5, dC)
            ref = (R.herk(uplo, o, -0.75, A, 0.5, C0.copy(order="F"), nb=nb) if R.available() and dt != np.float32
                   else O.herk(uplo, o, -0.75, A, 0.5, C0.copy(order="F")))   # the reference driver has no float herk
            got = dC.ToGlobal()
            assert np.array_equal(got[~mask], C0[~mask])
            assert np.linalg.norm(got - ref) <= 4 * k * e * (np.linalg.norm(A) ** 2 + np.linalg.norm(C0))
            dC = _dm(El, C0)
            oo = "N" if o == "N" else "T"
            El.Syrk(UL[uplo], ORI[oo], 1.5, _dm(El, A), -2.0, dC)
            P = (A @ A.T) if o == "N" else (A.T @ A)
            want = np.where(mask, 1.5 * P - 2.0 * C0, C0)
            got = dC.ToGlobal()
            assert np.array_equal(got[~mask], C0[~mask])
            assert np.linalg.norm(got - want) <= 4 * k * e * (1.5 * np.linalg.norm(A) ** 2 + 2 * np.linalg.norm(C0))
            for oa, ob in (("N", "N"), ("C", "N"), ("N", "T")):
                if (oa == "N") != (o == "N"):
                    continue
                Bm = O.fill(0, *((k, n) if ob == "N" else (n, k)), 5, dtype=dt)
                opA = A if oa == "N" else A.conj().T
                opB = Bm if ob == "N" else Bm.T
                dC = _dm(El, C0)
                El.Trrk(UL[uplo], ORI[oa], ORI[ob], 2.0, _dm(El, A), _dm(El, Bm), 0.25, dC)
                want = np.where(mask, 2.0 * (opA @ opB) + 0.25 * C0, C0)
                got = dC.ToGlobal()
                assert np.array_equal(got[~mask], C0[~mask])
                assert np.linalg.norm(got - want) <= 4 * k * e * (2 * np.linalg.norm(A) * np.linalg.norm(Bm) + np.linalg.norm(C0))
    El.SetBlocksize(128)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_syrk_dot_variant_long_summation_index(El, dt):
    """k > 10 n selects syrk::{LN,LT,UN,UT}_Dot (Syrk/LN.hpp:88-157): summation index over all processes, one local
    product and one sum-scatter per block of the triangle.  Small Dot blocksize so that several blocks are formed."""
    n, k = 70, 900
    e = np.finfo(np.float64).eps
    El.SetGemmDotBlocksize(32)
    try:
        for uplo in "LU":
            for o in "NTC":
                conj = o == "C"
                A = O.fill(0, *((n, k) if o == "N" else (k, n)), 3, dtype=dt)
                C0 = O.fill(0, n, n, 4, dtype=dt)
                dC = _dm(El, C0)
                if conj:
                    El.Herk(UL[uplo], ORI[o], 1.25, _dm(El, A), 0.5, dC)
                    P = A.conj().T @ A
                elif o == "T":
                    El.Syrk(UL[uplo], ORI[o], 1.25, _dm(El, A), 0.5, dC)
                    P = A.T @ A
                else:
                    El.Syrk(UL[uplo], ORI[o], 1.25, _dm(El, A), 0.5, dC)
                    P = A @ A.T
                mask = O._tri_mask(n, n, uplo)
                want = np.where(mask, 1.25 * P + 0.5 * C0, C0)
                got = dC.ToGlobal()
                assert np.array_equal(got[~mask], C0[~mask]), (dt, uplo, o)
                assert np.linalg.norm(got - want) <= 4 * k * e * (np.linalg.norm(A) ** 2 + np.linalg.norm(C0)), (dt, uplo, o)
    finally:
        El.SetGemmDotBlocksize(0)


@pytest.mark.parametrize("dt", [np.float64, np.complex64])
def test_axpy_trapezoid_dist(El, dt):
    """AxpyTrapezoid (level1/AxpyTrapezoid.hpp:128-160): same distribution -> local staircase kernel; different
    distribution -> X is redistributed first.  Bit-exact for alpha a power of two."""
    n = 61
    X, Y0 = O.fill(0, n, n, 1, dtype=dt), O.fill(0, n, n, 2, dtype=dt)
    ii, jj = np.indices((n, n))
    for uplo in "LU":
        for off in (0, 2, -3):
            inside = (jj - ii <= off) if uplo == "L" else (jj - ii >= off)
            want = np.where(inside, Y0 + dt(0.5) * X, Y0)
            for xdist in ((El.MC, El.MR), (El.VC, El.STAR), (El.STAR, El.STAR)):
                dX = El.DistMatrix(dt, xdist[0], xdist[1]); dX.FromGlobal(X)
                dY = _dm(El, Y0)
                El.AxpyTrapezoid(UL[uplo], 0.5, dX, dY, off)
                assert np.array_equal(dY.ToGlobal(), want), (uplo, off, xdist)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_herk_trrk_dist(El, dt):
    n, k, nb = 170, 90, 48
    for uplo in "LU":
        for o in "NC":
            A = O.fill(0, *((n, k) if o == "N" else (k, n)), 3, dtype=dt)
            C0 = O.fill(0, n, n, 4, dtype=dt)
            dA, dC = _dm(El, A), _dm(El, C0)
            El.PushBlocksizeStack(nb)
            El.Herk(0 if uplo == "L" else 1, ORI[o], -1.0, dA, 1.0, dC)
            El.PopBlocksizeStack()
            ref = (R.herk(uplo, o, -1.0, A, 1.0, C0.copy(order="F"), nb=nb) if R.available()
                   else O.herk(uplo, o, -1.0, A, 1.0, C0.copy(order="F")))
            got = dC.ToGlobal()
            mask = O._tri_mask(n, n, uplo)
            assert np.array_equal(got[~mask], C0[~mask])
            assert np.linalg.norm(got - ref) <= 4 * k * np.finfo(np.float64).eps * np.linalg.norm(A) ** 2


def test_redistribution_all_pairs_single_rank(El):
    """tests/core/DistMatrix.cpp on one rank: every B[U',V'] = A[U,V] must reproduce A entrywise."""
    from planutil import LEGAL
    G = O.fill(0, 37, 29, 1)
    for (u, v) in LEGAL:
        A = El.DistMatrix(np.float64, u, v)
        A.FromGlobal(G)
        for (u2, v2) in LEGAL:
            B = El.DistMatrix(np.float64, u2, v2)
            El.Copy(A, B)
            assert np.array_equal(B.ToGlobal(), G)
            Bt = El.DistMatrix(np.float64, u2, v2)
            El.Transpose(A, Bt)
            assert np.array_equal(Bt.ToGlobal(), G.T)


def test_views_and_blocksize_stack(El):
    G = O.fill(0, 40, 30, 1)
    A = _dm(El, G)
    V = El.DistMatrix(np.float64).View(A, 5, 7, 20, 11)
    assert np.array_equal(V.ToGlobal(), G[5:25, 7:18])
    b0 = El.Blocksize()
    assert b0 == 128  # default pushed at init (src/core/environment.cpp:181-183)
    El.PushBlocksizeStack(77); assert El.Blocksize() == 77
    El.SetBlocksize(55); assert El.Blocksize() == 55
    El.PopBlocksizeStack(); assert El.Blocksize() == b0
    # shrinking keeps the leading dimension and the leading block (Matrix/impl.hpp:698-717: "simply shrink our view")
    ld = A.LDim()
    A.Resize(25, 20)
    assert A.LDim() == ld and np.array_equal(A.ToGlobal(), G[:25, :20])
    A.Resize(41, 20)   # growing reallocates: contents undefined, size right
    assert (A.Height(), A.Width()) == (41, 20) and A.LDim() >= 41


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_binary_and_binaryflat_files_device_roundtrip(El, dt, tmp_path):
    """BINARY / BINARY_FLAT (src/io/Write/Binary.hpp:16-36, Write/BinaryFlat.hpp:16-33, Read/Binary.hpp,
    Read/BinaryFlat.hpp:37-102): the bytes on disk are the reference's -- column-major entries, BINARY prefixed by two
    32-bit Ints -- and files in that layout read back bit-exactly into any distribution."""
    h, w = 93, 57
    A = O.fill(0, h, w, 21, dtype=dt)
    dA = _dm(El, A)
    El.Write(dA, str(tmp_path / "m"), El.BINARY)
    El.Write(dA, str(tmp_path / "m"), El.BINARY_FLAT)
    raw = np.asfortranarray(A).tobytes(order="F")
    assert (tmp_path / "m.dat").read_bytes() == raw
    assert (tmp_path / "m.bin").read_bytes() == np.array([h, w], dtype=np.int32).tobytes() + raw
    for dist in ((El.MC, El.MR), (El.STAR, El.VC), (El.VC, El.STAR), (El.STAR, El.STAR), (El.MR, El.MC)):
        B = El.DistMatrix(dt, dist[0], dist[1])
        El.ReadBinaryFlat(B, h, w, str(tmp_path / "m.dat"))
        assert np.array_equal(B.ToGlobal(), A), dist
        B2 = El.DistMatrix(dt, dist[0], dist[1])
        El.ReadBinary(B2, str(tmp_path / "m.bin"))
        assert (B2.Height(), B2.Width()) == (h, w) and np.array_equal(B2.ToGlobal(), A), dist
    with pytest.raises(El.Elb200Error):
        El.ReadBinaryFlat(El.DistMatrix(dt), h + 1, w, str(tmp_path / "m.dat"))   # size check (BinaryFlat.hpp:24-28)
