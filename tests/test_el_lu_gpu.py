"""GPU parity of LU, the permutation object, lu::SolveAfter and LinearSolve against the
reference itself (oracle/_ref/libElRef.so: El::LU, lu::SolveAfter, El::LinearSolve from its own sources) when built,
else the numpy restatement (oracle/elemental_oracle.py: lu, lu_panel, lu_solve_after -- pinned to the reference in
tests/test_oracle_cpu.py).  Partial pivoting is deterministic (i?amax: largest |x|, first occurrence), so the
permutation must be IDENTICAL to the reference's; the packed factors agree to rounding:
   ||LU - LU_ref||_F <= 50 n eps ||A||_F        ||P A - L U||_F <= 10 n eps ||A||_F
"""
import ctypes as C

import numpy as np
import pytest

from oracle import elemental_oracle as O
from oracle import reference_lib as R

pytestmark = pytest.mark.gpu

DT = [np.float64, np.complex128, np.float32, np.complex64]
ORI = {"N": 0, "T": 1, "C": 2}


@pytest.fixture(scope="module")
def El():
    from elemental_b200 import api
    api.Initialize()
    return api


def _dm(El, a, dist=(0, 2)):
    M = El.DistMatrix(a.dtype, dist[0], dist[1])
    M.FromGlobal(a)
    return M


def _eps(dt):
    return np.finfo(np.dtype(dt).char.lower() if np.dtype(dt).kind == "c" else dt).eps


def _ref_lu(A, nb):
    """(packed factors, preimage vector) of the reference at this Blocksize()"""
    F = A.copy(order="F")
    if R.available():
        return R.lu_piv(F, nb=nb)
    p = O.lu(F, nb)
    return F, p


def _split(F):
    m, n = F.shape
    k = min(m, n)
    return np.tril(F[:, :k], -1) + np.eye(m, k, dtype=F.dtype), np.triu(F[:k, :])


@pytest.mark.parametrize("dt", DT)
def test_getrf_panel_kernel_matches_the_unblocked_reference_panel(dt):
    """elb200_?getrf_panel (one cooperative kernel) against lu::Panel restated in numpy: same pivots, same factors to
    rounding, on one-CTA and many-CTA panels, with ragged heights and a padded leading dimension."""
    import gpuutil as G
    import torch
    rng = np.random.default_rng(31)
    fn = getattr(G.lib(), f"elb200_{G.SUF[np.dtype(dt)]}getrf_panel")
    for (m, n) in [(1, 1), (5, 3), (64, 64), (300, 37), (1000, 128), (5000, 64), (40000, 32)]:
        A = G.rand(rng, m, n, dt)
        dA = G.DevMat(A, m + 3, offset=1)
        ipiv = torch.full((n,), -7, dtype=torch.int64, device="cuda")
        info = torch.zeros(1, dtype=torch.int32, device="cuda")
        G.check(fn(G.i64(m), G.i64(n), dA.ptr, G.i64(dA.ld), C.c_void_p(ipiv.data_ptr()), 1, C.c_void_p(info.data_ptr()),
                   G.stream()), "getrf_panel")
        ref = A.copy(order="F")
        piv = O.lu_panel(ref, True)
        assert int(info.item()) == 0
        got, gp = dA.get(), ipiv.cpu().numpy()
        double = np.dtype(dt) in (np.dtype(np.float64), np.dtype(np.complex128))
        if double:      # single precision may legitimately pick the other of two candidates that agree to rounding
            assert np.array_equal(gp, piv), (dt, m, n)
        if np.array_equal(gp, piv):
            assert np.linalg.norm(got - ref) <= 20 * n * _eps(dt) * np.linalg.norm(A), (dt, m, n)
        assert np.all((gp >= np.arange(n)) & (gp < m))
        PA = A.copy()
        for j, pj in enumerate(gp):
            PA[[j, pj], :] = PA[[pj, j], :]
        L, U = _split(got)
        assert np.linalg.norm(PA - L @ U) <= 20 * n * _eps(dt) * np.linalg.norm(A), (dt, m, n)
        # partial pivoting bounds the multipliers by one in the norm i?amax compares (sqrt 2 in modulus for complex)
        assert np.max(np.abs(L)) <= (1.4143 if np.dtype(dt).kind == "c" else 1 + 1e-5)
        assert dA.padding_untouched()
    # without pivoting (lu::Unb) on a diagonally dominant block
    m, n = 700, 96
    A = G.rand(rng, m, n, dt)
    A[:n, :n] += 2 * n * np.eye(n, dtype=dt)
    dA = G.DevMat(A, m + 1)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    G.check(fn(G.i64(m), G.i64(n), dA.ptr, G.i64(dA.ld), None, 0, C.c_void_p(info.data_ptr()), G.stream()), "getrf_panel")
    ref = A.copy(order="F")
    O.lu_panel(ref, False)
    assert int(info.item()) == 0 and np.linalg.norm(dA.get() - ref) <= 20 * n * _eps(dt) * np.linalg.norm(A)
    # an exactly zero pivot column is reported (1-based column), not divided by
    A = G.rand(rng, 200, 16, dt)
    A[:, 5] = 0
    dA = G.DevMat(A)
    ipiv = torch.zeros(16, dtype=torch.int64, device="cuda")
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    G.check(fn(G.i64(200), G.i64(16), dA.ptr, G.i64(dA.ld), C.c_void_p(ipiv.data_ptr()), 1, C.c_void_p(info.data_ptr()),
               G.stream()), "getrf_panel")
    assert int(info.item()) == 6 and np.all(np.isfinite(dA.get()))


@pytest.mark.parametrize("dt", DT)
def test_lu_partial_pivoting_matches_reference(El, dt):
    for (m, n, nb) in [(300, 300, 64), (200, 330, 48), (330, 200, 64), (65, 65, 128), (257, 257, 32), (1, 1, 8)]:
        A = O.fill(0, m, n, 11, dtype=dt)
        Fref, pref = _ref_lu(A, nb)
        El.PushBlocksizeStack(nb)
        try:
            dA = _dm(El, A)
            P = El.DistPermutation()
            El.LU(dA, P)
        finally:
            El.PopBlocksizeStack()
        got = dA.ToGlobal()
        p = P.Preimages()
        assert P.Height() == m and P.IsSwapSequence() and P.IsImplicitSwapSequence()
        # single precision: two pivot candidates can lie within rounding of each other, and the trailing updates of
        # the two implementations round differently -- the permutation is then required to be A valid one (residual
        # below), not THE reference's; in double precision it must be identical
        if np.dtype(dt) in (np.dtype(np.float64), np.dtype(np.complex128)):
            assert np.array_equal(p, pref), (dt, m, n, nb)
        assert sorted(p.tolist()) == list(range(m))
        if np.array_equal(p, pref):
            tol = 50 * max(m, n) * _eps(dt) * np.linalg.norm(A)
            assert np.linalg.norm(got - Fref) <= tol, (dt, m, n, nb)
        L, U = _split(got)
        assert np.linalg.norm(A[p, :] - L @ U) <= 10 * max(m, n) * _eps(dt) * np.linalg.norm(A)
        # the permutation object: images invert preimages, parity = parity of the swap count that moved something
        img = np.array([P.Image(i) for i in range(min(m, 40))])
        assert np.array_equal(p[img], np.arange(min(m, 40)))
        seen, cycles = np.zeros(m, bool), 0
        for i in range(m):
            if not seen[i]:
                cycles += 1
                j = i
                while not seen[j]:
                    seen[j] = True
                    j = p[j]
        assert P.Parity() == bool((m - cycles) & 1)


@pytest.mark.parametrize("dt", [np.float64, np.complex128])
def test_lu_without_pivoting_and_other_distributions(El, dt):
    n, nb = 280, 64
    A = O.fill(0, n, n, 12, dtype=dt) + n * np.eye(n, dtype=dt)
    ref = A.copy(order="F")
    if R.available():
        R.lu(ref, nb=nb)
    else:
        O.lu(ref, nb, pivot=False)
    El.PushBlocksizeStack(nb)
    try:
        for dist in ((El.MC, El.MR), (El.VC, El.STAR), (El.STAR, El.STAR), (El.MR, El.MC)):
            dA = _dm(El, A, dist)
            El.LU(dA)
            assert np.linalg.norm(dA.ToGlobal() - ref) <= 50 * n * _eps(dt) * np.linalg.norm(A), dist
        # partial pivoting on a matrix that is not [MC,MR]
        B = O.fill(0, n, n, 13, dtype=dt)
        Fref, pref = _ref_lu(B, nb)
        dB = _dm(El, B, (El.VR, El.STAR))
        P = El.DistPermutation()
        El.LU(dB, P)
        assert np.array_equal(P.Preimages(), pref)
        assert np.linalg.norm(dB.ToGlobal() - Fref) <= 50 * n * _eps(dt) * np.linalg.norm(B)
    finally:
        El.PopBlocksizeStack()


def test_lu_singular_matrix_raises(El):
    n = 150
    A = O.fill(0, n, n, 14)
    A[:, 70] = 0.0
    with pytest.raises(El.SingularMatrixException):
        El.LU(_dm(El, A), El.DistPermutation())
    B = O.fill(0, n, n, 15) + n * np.eye(n)
    B[40, 40] = 0.0
    B[40, :40] = 0.0     # the (40,40) pivot stays exactly zero without interchanges
    with pytest.raises(El.SingularMatrixException):
        El.LU(_dm(El, B))


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_lu_solve_after_and_linear_solve(El, dt):
    n, nrhs, nb = 260, 70, 64
    A = O.fill(0, n, n, 16, dtype=dt)
    B = O.fill(0, n, nrhs, 17, dtype=dt)
    El.PushBlocksizeStack(nb)
    try:
        dF = _dm(El, A)
        P = El.DistPermutation()
        El.LU(dF, P)
        for o in "NTC":
            dB = _dm(El, B)
            El.LUSolveAfter(ORI[o], dF, dB, P)
            X = dB.ToGlobal()
            opA = A if o == "N" else (A.T if o == "T" else A.conj().T)
            res = np.linalg.norm(opA @ X - B) / (n * _eps(dt) * np.linalg.norm(A) * np.linalg.norm(X))
            assert res <= 10.0, (dt, o, res)
            if R.available() and dt != np.float32:
                Xref = R.lu_piv_solve(o, A, B.copy(order="F"), nb=nb)
                assert np.linalg.norm(X - Xref) <= 1e3 * n * _eps(dt) * np.linalg.norm(Xref), (dt, o)
        # unpivoted pair on a diagonally dominant matrix
        Ad = A + n * np.eye(n, dtype=dt)
        dF = _dm(El, Ad)
        El.LU(dF)
        for o in "NC":
            dB = _dm(El, B)
            El.LUSolveAfter(ORI[o], dF, dB)
            X = dB.ToGlobal()
            opA = Ad if o == "N" else Ad.conj().T
            assert np.linalg.norm(opA @ X - B) <= 10 * n * _eps(dt) * np.linalg.norm(Ad) * np.linalg.norm(X)
        # LinearSolve leaves A alone and accepts any distribution of B
        dA = _dm(El, A)
        dB = _dm(El, B, (El.VC, El.STAR))
        El.LinearSolve(dA, dB)
        X = dB.ToGlobal()
        assert np.array_equal(dA.ToGlobal(), A)
        assert np.linalg.norm(A @ X - B) <= 10 * n * _eps(dt) * np.linalg.norm(A) * np.linalg.norm(X)
        if R.available() and dt != np.float32:
            Xref = R.linear_solve(A, B.copy(order="F"), nb=nb)
            assert np.linalg.norm(X - Xref) <= 1e3 * n * _eps(dt) * np.linalg.norm(Xref)
    finally:
        El.PopBlocksizeStack()


@pytest.mark.parametrize("dt", [np.float64, np.complex64])
def test_permutation_object_against_numpy(El, dt):
    """Swap / SwapSequence / PermuteRows / PermuteCols and their inverses, with offsets, on several distributions
    (src/lapack_like/perm/Permutation.cpp:207-250, 545-720)."""
    rng = np.random.default_rng(41)
    m, n, sz, off = 53, 47, 30, 9
    A = O.fill(0, m, n, 18, dtype=dt)
    P = El.DistPermutation()
    P.MakeIdentity(sz)
    P.ReserveSwaps(4)
    pre = np.arange(sz)
    swaps = [(int(a), int(b)) for a, b in zip(rng.integers(0, sz, 25), rng.integers(0, sz, 25))]
    for a, b in swaps:      # more than reserved: the list grows
        P.Swap(a, b)
        pre[[a, b]] = pre[[b, a]]
    assert not P.IsImplicitSwapSequence()
    assert np.array_equal(P.Preimages(), pre)
    assert all(P.Preimage(i) == pre[i] and P.Image(int(pre[i])) == i for i in range(sz))
    for dist in ((El.MC, El.MR), (El.STAR, El.VR), (El.VC, El.STAR), (El.STAR, El.STAR), (El.MR, El.MC)):
        d = _dm(El, A, dist)
        P.PermuteRows(d, off)
        want = A.copy()
        want[off:off + sz, :] = A[off + pre, :]
        assert np.array_equal(d.ToGlobal(), want), dist
        P.InversePermuteRows(d, off)
        assert np.array_equal(d.ToGlobal(), A), dist
        P.PermuteCols(d, off)
        want = A.copy()
        want[:, off:off + sz] = A[:, off + pre]
        assert np.array_equal(d.ToGlobal(), want), dist
        P.InversePermuteCols(d, off)
        assert np.array_equal(d.ToGlobal(), A), dist
    # SwapSequence appends another permutation's swaps at an offset
    Q = El.DistPermutation()
    Q.MakeIdentity(sz + 5)
    Q.Swap(0, 3)
    Q.SwapSequence(P, 5)
    q = np.arange(sz + 5)
    q[[0, 3]] = q[[3, 0]]
    for a, b in swaps:
        q[[a + 5, b + 5]] = q[[b + 5, a + 5]]
    assert np.array_equal(Q.Preimages(), q)
    with pytest.raises(El.Elb200Error):
        P.Swap(0, sz)
    with pytest.raises(El.Elb200Error):
        P.PermuteRows(_dm(El, A), m - sz + 1)


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_cholesky_mod_matches_reference(El, dt):
    """El::CholeskyMod (Cholesky/LowerMod.hpp, UpperMod.hpp): update (alpha > 0) and downdate (alpha < 0) of a factor
    by a rank-w term, against the reference itself when built (else the unique factor of the modified matrix), for
    both triangles, several widths, a Blocksize() smaller than n, and a [VC,*] V."""
    n, nb = 230, 64
    G = O.fill(0, n, n, 61, dtype=dt)
    S = G @ G.conj().T / n + np.eye(n, dtype=dt)
    S = ((S + S.conj().T) / 2).astype(dt)
    L0 = np.linalg.cholesky(S.astype(np.complex128 if np.dtype(dt).kind == "c" else np.float64)).astype(dt)
    eps = _eps(dt)
    El.PushBlocksizeStack(nb)
    try:
        for uplo in "LU":
            T0 = np.asfortranarray(L0 if uplo == "L" else L0.conj().T)
            # junk in the other triangle must survive
            junk = O.fill(0, n, n, 62, dtype=dt)
            T0 = np.asfortranarray(T0 + (np.triu(junk, 1) if uplo == "L" else np.tril(junk, -1)))
            for w, alpha in ((1, 0.7), (5, 0.5), (40, 0.3), (3, -0.004), (17, -0.002)):
                V = np.asfortranarray(O.fill(0, n, w, 63 + w, dtype=dt))
                dT = _dm(El, T0)
                dV = _dm(El, V, (El.VC, El.STAR) if w == 5 else (El.MC, El.MR))
                El.CholeskyMod(0 if uplo == "L" else 1, dT, alpha, dV)
                got = dT.ToGlobal()
                tri = np.tril(got) if uplo == "L" else np.triu(got)
                other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
                assert np.array_equal(other, np.triu(T0, 1) if uplo == "L" else np.tril(T0, -1)), (uplo, w)
                Lg = tri if uplo == "L" else tri.conj().T
                target = S + alpha * (V @ V.conj().T)
                res = np.linalg.norm(Lg @ Lg.conj().T - target) / (n * eps * np.linalg.norm(target))
                assert res <= 10.0, (dt, uplo, w, alpha, res)
                assert np.all(np.diag(Lg).real > 0) and np.all(np.abs(np.diag(Lg).imag) == 0)
                if R.available() and dt != np.float32:
                    ref = R.cholesky_mod(uplo, T0.copy(order="F"), alpha, V.copy(order="F"), nb=nb)
                    rtri = np.tril(ref) if uplo == "L" else np.triu(ref)
                    assert np.linalg.norm(tri - rtri) <= 1e3 * n * eps * np.linalg.norm(rtri), (dt, uplo, w, alpha)
        # a downdate that leaves the matrix indefinite raises, as the reference's hyperbolic reflector does
        V = np.asfortranarray(10.0 * O.fill(0, n, 2, 64, dtype=dt))
        with pytest.raises(El.LogicError):
            El.CholeskyMod(0, _dm(El, np.asfortranarray(L0)), -1.0, _dm(El, V))
    finally:
        El.PopBlocksizeStack()


@pytest.mark.parametrize("dt", [np.float64, np.complex128, np.float32])
def test_pivoted_cholesky(El, dt):
    """El::Cholesky(uplo, A, P): P A P^T = L L^H with the largest remaining diagonal entry as pivot.  The reference's
    own pivot order depends on stale memory (see tests/test_oracle_cpu.py), so the oracle is the numpy restatement of
    the algorithm as stated (oracle.cholesky_pivoted: identical permutation, factor to rounding) plus the identity,
    the ordered diagonal, the untouched other triangle and the solve after."""
    n, nb = 210, 48
    G = O.fill(0, n, n, 71, dtype=dt)
    S = G @ G.conj().T / n + 0.05 * np.eye(n, dtype=dt)
    S = np.asfortranarray(((S + S.conj().T) / 2).astype(dt))
    eps = _eps(dt)
    junk = O.fill(0, n, n, 72, dtype=dt)
    B = O.fill(0, n, 9, 73, dtype=dt)
    El.PushBlocksizeStack(nb)
    try:
        for uplo in "LU":
            A0 = np.asfortranarray((np.tril(S) + np.triu(junk, 1)) if uplo == "L" else (np.triu(S) + np.tril(junk, -1)))
            for dist in ((El.MC, El.MR), (El.VC, El.STAR)):
                dA = _dm(El, A0, dist)
                P = El.DistPermutation()
                El.CholeskyPiv(0 if uplo == "L" else 1, dA, P)
                got, p = dA.ToGlobal(), P.Preimages()
                Fo = A0.copy(order="F")
                po = O.cholesky_pivoted(uplo, Fo)
                Lg = np.tril(got) if uplo == "L" else np.triu(got).conj().T
                other = np.triu(got, 1) if uplo == "L" else np.tril(got, -1)
                assert np.array_equal(other, np.triu(A0, 1) if uplo == "L" else np.tril(A0, -1)), (uplo, dist)
                assert sorted(p.tolist()) == list(range(n))
                res = np.linalg.norm(S[np.ix_(p, p)] - Lg @ Lg.conj().T) / (n * eps * np.linalg.norm(S))
                assert res <= 10.0, (dt, uplo, res)
                dg = np.diag(Lg).real
                assert np.all(np.diff(dg) <= 50 * n * eps * dg[0]), (dt, uplo)      # pivoting orders the diagonal
                if dt != np.float32:
                    assert np.array_equal(p, po), (dt, uplo)
                    Lo = np.tril(Fo) if uplo == "L" else np.triu(Fo).conj().T
                    assert np.linalg.norm(Lg - Lo) <= 1e3 * n * eps * np.linalg.norm(Lo), (dt, uplo)
                for o in "NT":
                    dB = _dm(El, B)
                    El.CholeskyPivSolveAfter(0 if uplo == "L" else 1, ORI[o], dA, P, dB)
                    X = dB.ToGlobal()
                    opS = S if o == "N" else S.T
                    assert np.linalg.norm(opS @ X - B) <= 1e3 * n * eps * np.linalg.norm(S) * np.linalg.norm(X), (dt, uplo, o)
        # semidefinite input of rank 40: pivoting pushes the zero pivots to the end, where the factorisation stops
        Glow = O.fill(0, n, 40, 74, dtype=dt)
        Z = np.asfortranarray((Glow @ Glow.conj().T).astype(dt))
        Z[np.diag_indices(n)] = Z[np.diag_indices(n)].real
        with pytest.raises(El.NonHPDMatrixException):
            El.CholeskyPiv(0, _dm(El, Z - 1e-3 * np.eye(n, dtype=dt)), El.DistPermutation())
    finally:
        El.PopBlocksizeStack()
