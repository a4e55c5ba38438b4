"""GPU parity of the level-3 siblings SURVEY.md section 8f ranks next -- El::Syr2k / Her2k, Symm / Hemm, Trmm
(elemental_b200/csrc/host/level3.cpp, compositions over Trrk / Gemm / the redistribution engine) -- against the
reference library (or the numpy restatement that tests/test_oracle_cpu.py pins to it).

First run on a B200 in round 2 (gpurun_out/r2_siblings.log: 7 passed); the round-1 environment gate is gone.

Tolerance: ||C - C_ref||_F <= 4 k eps ||A||_F ||B||_F as for Gemm; entries outside the triangle bit-identical."""
import numpy as np
import pytest

from oracle import elemental_oracle as O
from oracle import reference_lib as R

pytestmark = pytest.mark.gpu

ORI = {"N": 0, "T": 1, "C": 2}
UL = {"L": 0, "U": 1}
LR = {"L": 0, "R": 1}
DG = {"N": 0, "U": 1}


@pytest.fixture(scope="module")
def El():
    from elemental_b200 import api
    api.Initialize()
    return api


def _dm(El, a):
    M = El.DistMatrix(a.dtype, 0, 2)
    M.FromGlobal(a)
    return M


def _tol(k, dt, *mats):
    e = np.finfo(np.dtype(dt).type(0).real.dtype).eps
    prod = 1.0
    for m in mats:
        prod *= np.linalg.norm(m)
    return 4 * k * e * prod


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_syr2k_her2k(El, dt):
    n, k, nb = 150, 70, 32
    alpha = 0.7 if dt == np.float64 else 0.7 - 0.3j
    for conj in (False, True):
        beta = 1.5 if (conj or dt == np.float64) else 1.5 + 0.25j
        for uplo in "LU":
            for orient in ("N", "C" if conj else "T"):
                A = O.fill(0, *((n, k) if orient == "N" else (k, n)), 1, dtype=dt)
                B = O.fill(0, *((n, k) if orient == "N" else (k, n)), 2, dtype=dt)
                C0 = O.fill(0, n, n, 3, dtype=dt)
                ref = (R.syr2k(uplo, orient, alpha, A, B, beta, C0.copy(order="F"), conjugate=conj, nb=nb)
                       if R.available() else O.syr2k(uplo, orient, alpha, A, B, beta, C0.copy(order="F"), conjugate=conj))
                dA, dB, dC = _dm(El, A), _dm(El, B), _dm(El, C0)
                El.PushBlocksizeStack(nb)
                (El.Her2k if conj else El.Syr2k)(UL[uplo], ORI[orient], alpha, dA, dB, beta, dC)
                El.PopBlocksizeStack()
                got = dC.ToGlobal()
                mask = O._tri_mask(n, n, uplo)
                assert np.array_equal(got[~mask], C0[~mask]), "the other triangle must not be touched"
                assert np.linalg.norm((got - ref)[mask]) <= 2 * _tol(k, dt, A, B), (dt, conj, uplo, orient)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_symm_hemm(El, dt):
    m, n, nb = 130, 90, 32
    alpha, beta = (0.7, 1.5) if dt == np.float64 else (0.7 - 0.3j, 1.5 + 0.25j)
    for conj in (False, True):
        for side in "LR":
            for uplo in "LU":
                ka = m if side == "L" else n
                A = O.fill(0, ka, ka, 1, dtype=dt)
                if conj and dt == np.complex128:
                    A[np.arange(ka), np.arange(ka)] = A[np.arange(ka), np.arange(ka)].real
                B, C0 = O.fill(0, m, n, 2, dtype=dt), O.fill(0, m, n, 3, dtype=dt)
                ref = (R.symm(side, uplo, alpha, A, B, beta, C0.copy(order="F"), conjugate=conj, nb=nb)
                       if R.available() else O.symm(side, uplo, alpha, A, B, beta, C0.copy(order="F"), conjugate=conj))
                dA, dB, dC = _dm(El, A), _dm(El, B), _dm(El, C0)
                El.PushBlocksizeStack(nb)
                (El.Hemm if conj else El.Symm)(LR[side], UL[uplo], alpha, dA, dB, beta, dC)
                El.PopBlocksizeStack()
                assert np.array_equal(dA.ToGlobal(), A), "A is read-only"
                assert np.linalg.norm(dC.ToGlobal() - ref) <= _tol(ka, dt, A, B), (dt, conj, side, uplo)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_trmm(El, dt):
    m, n, nb = 120, 75, 32
    alpha = 0.7 if dt == np.float64 else 0.7 - 0.3j
    for side in "LR":
        for uplo in "LU":
            for orient in "NTC":
                for diag in "NU":
                    ka = m if side == "L" else n
                    A, B0 = O.fill(0, ka, ka, 1, dtype=dt), O.fill(0, m, n, 2, dtype=dt)
                    ref = (R.trmm(side, uplo, orient, diag, alpha, A, B0.copy(order="F"), nb=nb)
                           if R.available() else O.trmm(side, uplo, orient, diag, alpha, A, B0.copy(order="F")))
                    dA, dB = _dm(El, A), _dm(El, B0)
                    El.PushBlocksizeStack(nb)
                    El.Trmm(LR[side], UL[uplo], ORI[orient], DG[diag], alpha, dA, dB)
                    El.PopBlocksizeStack()
                    assert np.linalg.norm(dB.ToGlobal() - ref) <= _tol(ka, dt, A, B0), (dt, side, uplo, orient, diag)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_trr2k_all_orientation_cases(El, dt):
    """El::Trr2k (src/blas_like/level3/Trr2k.cpp:34-...): E_tri := alpha op(A) op(B) + beta op(C) op(D) + gamma E_tri
    for the 16 NORMAL / (conjugate-)transposed cases; here one masked GEMM over the stacked panels per step.  The
    strictly-other triangle of E must stay bit-identical."""
    n, k, nb = 130, 70, 32
    alpha = 0.7 if dt == np.float64 else 0.7 - 0.3j
    beta = -1.25 if dt == np.float64 else -1.25 + 0.5j
    gamma = 0.5
    e = np.finfo(np.float64).eps
    tr = "C" if dt == np.complex128 else "T"
    op = lambda M, o: M if o == "N" else (M.T if o == "T" else M.conj().T)
    for uplo in "LU":
        for oa in ("N", tr):
            for ob in ("N", tr):
                for oc in ("N", tr):
                    for od in ("N", tr):
                        A = O.fill(0, *((n, k) if oa == "N" else (k, n)), 1, dtype=dt)
                        B = O.fill(0, *((k, n) if ob == "N" else (n, k)), 2, dtype=dt)
                        Cc = O.fill(0, *((n, k) if oc == "N" else (k, n)), 3, dtype=dt)
                        D = O.fill(0, *((k, n) if od == "N" else (n, k)), 4, dtype=dt)
                        E0 = O.fill(0, n, n, 5, dtype=dt)
                        dE = _dm(El, E0)
                        El.PushBlocksizeStack(nb)
                        El.Trr2k(UL[uplo], ORI[oa], ORI[ob], ORI[oc], ORI[od], alpha, _dm(El, A), _dm(El, B), beta,
                                 _dm(El, Cc), _dm(El, D), gamma, dE)
                        El.PopBlocksizeStack()
                        mask = O._tri_mask(n, n, uplo)
                        full = alpha * (op(A, oa) @ op(B, ob)) + beta * (op(Cc, oc) @ op(D, od)) + gamma * E0
                        want = np.where(mask, full, E0)
                        got = dE.ToGlobal()
                        assert np.array_equal(got[~mask], E0[~mask]), (dt, uplo, oa, ob, oc, od)
                        bound = 8 * k * e * (np.linalg.norm(A) * np.linalg.norm(B) + np.linalg.norm(Cc) * np.linalg.norm(D) + np.linalg.norm(E0))
                        assert np.linalg.norm(got - want) <= bound, (dt, uplo, oa, ob, oc, od)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_two_sided_trsm_trmm(El, dt):
    """El::TwoSidedTrsm / TwoSidedTrmm (TwoSidedTrsm/{LVar4,UVar4}.hpp, TwoSidedTrmm/*.hpp): only the `uplo` triangle
    of the Hermitian A is read and written; the other triangle stays bit-identical."""
    n, nb = 140, 32
    S = O.fill(1, n, n, 5, diag=float(n), dtype=dt)          # Hermitian, well conditioned
    Tfull = np.asfortranarray((O.fill(0, n, n, 7, dtype=dt) / n + 2 * np.eye(n)).astype(dt))
    junk = O.fill(0, n, n, 9, dtype=dt)
    e = np.finfo(np.float64).eps
    for uplo in "LU":
        mask = O._tri_mask(n, n, uplo)
        A0 = np.where(mask, S, junk)                          # garbage in the triangle that must not be referenced
        for diag in "NU":
            T = O._tri(Tfull, uplo, diag)
            Ti = np.linalg.inv(T)
            for solve in (True, False):
                dA = _dm(El, np.asfortranarray(A0))
                El.PushBlocksizeStack(nb)
                (El.TwoSidedTrsm if solve else El.TwoSidedTrmm)(UL[uplo], DG[diag], dA, _dm(El, Tfull))
                El.PopBlocksizeStack()
                if solve:
                    full = Ti @ S @ Ti.conj().T if uplo == "L" else Ti.conj().T @ S @ Ti
                else:
                    full = T.conj().T @ S @ T if uplo == "L" else T @ S @ T.conj().T
                got = dA.ToGlobal()
                assert np.array_equal(got[~mask], A0[~mask]), (dt, uplo, diag, solve)
                err = np.linalg.norm((got - full)[mask])
                assert err <= 200 * n * e * np.linalg.norm(full) * (np.linalg.cond(T) if solve else 1.0), (dt, uplo, diag, solve, err)


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_reverse_cholesky_and_variant2(El, dt):
    """El::ReverseCholesky (A = L^H L / U U^H, Cholesky/ReverseLowerVariant3.hpp:73-126,
    ReverseUpperVariant3.hpp:75-123) and the left-looking cholesky::{Lower,Upper}Variant2Blocked
    (LowerVariant2.hpp:43-110).  Residual ||A - F^H F|| / (n eps ||A||) <= 10 as for Cholesky; the triangle that is
    not referenced stays bit-identical; variant 2 agrees with variant 3."""
    n, nb = 150, 32
    A = O.fill(1, n, n, 5, diag=float(n), dtype=dt)
    junk = O.fill(0, n, n, 9, dtype=dt)
    e = np.finfo(np.dtype(dt).type(0).real.dtype).eps
    for uplo in "LU":
        mask = O._tri_mask(n, n, uplo)
        A0 = np.asfortranarray(np.where(mask, A, junk))
        dA = _dm(El, A0)
        El.PushBlocksizeStack(nb)
        El.ReverseCholesky(UL[uplo], dA)
        got = dA.ToGlobal()
        assert np.array_equal(got[~mask], A0[~mask]), (dt, uplo)
        F = np.where(mask, got, 0)
        rec = F.conj().T @ F if uplo == "L" else F @ F.conj().T
        assert np.linalg.norm(rec - A) / (n * e * np.linalg.norm(A)) <= 10, (dt, uplo)
        d3, d2 = _dm(El, A0), _dm(El, A0)
        El.Cholesky(UL[uplo], d3)
        El.CholeskyVariant2(UL[uplo], d2)
        El.PopBlocksizeStack()
        g3, g2 = d3.ToGlobal(), d2.ToGlobal()
        # variant 2 updates the FULL diagonal block (AxpyContract of X11, LowerVariant2.hpp:88-90), as the reference
        # does, so the other triangle of the diagonal blocks is not preserved; off-diagonal blocks are
        other_blocks = ~mask & (np.add.outer(np.arange(n) // nb, -(np.arange(n) // nb)) != 0)
        assert np.array_equal(g2[other_blocks], A0[other_blocks])
        assert O.cholesky_residual(uplo, g2, A) <= 10
        assert np.linalg.norm((g2 - g3)[mask]) <= 50 * n * e * np.linalg.norm(g3[mask])
    bad = A.copy(order="F")
    bad[40, 40] = -1.0
    with pytest.raises(El.NonHPDMatrixException):
        El.ReverseCholesky(0, _dm(El, bad))
