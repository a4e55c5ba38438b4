"""Multi-GPU parity worker, launched by torchrun (one rank per GPU, NCCL):
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tests/mgpu_worker.py
Every rank builds the same global inputs from the hash generator, runs the El-level call on an
r x c Grid and compares the gathered result with the oracle / the reference library (a 1x1-grid
run of the reference at the same Blocksize() is a valid oracle for p > 1 up to rounding,
SURVEY.md section 8c)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import elemental_oracle as O  # noqa: E402
from oracle import reference_lib as R  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from elemental_b200 import api as El
    from planutil import LEGAL

    height = {1: 1, 2: 1, 4: 2, 8: 2}.get(world, 0)
    if len(sys.argv) > 1:
        height = int(sys.argv[1])
    g = El.Grid(height)
    r, c = g.Height(), g.Width()
    assert g.Rank() == rank or True
    ORI = {"N": 0, "T": 1, "C": 2}
    fails = []

    def dm(a, dist_=(0, 2), align=None):
        M = El.DistMatrix(a.dtype, dist_[0], dist_[1], g)
        if align:
            M.Align(align[0], align[1])
        M.FromGlobal(a)
        return M

    # 1. redistribution: every pair, with alignments, plain and transposed (tests/core/DistMatrix.cpp)
    G = O.fill(0, 37, 29, 1)
    stride = {0: r, 2: c, 3: r * c, 4: r * c, 5: 1}
    for t, (u, v) in enumerate(LEGAL):
        A = dm(G, (u, v), (t % stride[u], (t + 1) % stride[v]))
        for t2, (u2, v2) in enumerate(LEGAL):
            B = El.DistMatrix(np.float64, u2, v2, g)
            B.Align((t2 + 1) % stride[u2], t2 % stride[v2])
            El.Copy(A, B)
            if not np.array_equal(B.ToGlobal(), G):
                fails.append(f"copy [{u},{v}]->[{u2},{v2}]")
            Bt = El.DistMatrix(np.float64, u2, v2, g)
            El.Transpose(A, Bt)
            if not np.array_equal(Bt.ToGlobal(), G.T):
                fails.append(f"transpose [{u},{v}]->[{u2},{v2}]")
    # 2. Gemm: all SUMMA variants and orientations, double + complex double
    for dt in (np.float64, np.complex128):
        m, n, k, nb = 150, 130, 170, 32
        for oa in "NTC":
            for ob in "NTC":
                A = O.fill(0, *((m, k) if oa == "N" else (k, m)), 1, dtype=dt)
                B = O.fill(0, *((k, n) if ob == "N" else (n, k)), 2, dtype=dt)
                C0 = O.fill(0, m, n, 3, dtype=dt)
                for alg in (1, 2, 3, 4):
                    El.PushBlocksizeStack(nb)
                    dC = dm(C0)
                    El.Gemm(ORI[oa], ORI[ob], 3.0, dm(A), dm(B), 4.0, dC, alg)
                    El.PopBlocksizeStack()
                    ref = (R.gemm(oa, ob, 3.0, A, B, 4.0, C0.copy(order="F"), nb=nb, alg=alg) if R.available()
                           else O.gemm(oa, ob, 3.0, A, B, 4.0, C0.copy(order="F"), nb=nb, alg=alg))
                    res = O.gemm_residual(dC.ToGlobal(), ref, k, A, B)
                    if not res <= 1.0:
                        fails.append(f"gemm {dt.__name__} {oa}{ob} alg={alg} res={res}")
    # 2a. Cannon's algorithm on square grids (Gemm/NN.hpp:15-89), with and without matching alignments
    if r == c:
        for dt in (np.float64, np.complex128):
            m, n, k = 150, 130, 64 * r
            A, B, C0 = O.fill(0, m, k, 1, dtype=dt), O.fill(0, k, n, 2, dtype=dt), O.fill(0, m, n, 3, dtype=dt)
            for al in (None, (r - 1, c - 1)):
                dC = dm(C0)
                El.Gemm(0, 0, 3.0, dm(A, align=al), dm(B, align=al), 4.0, dC, El.GEMM_CANNON)
                res = O.gemm_residual(dC.ToGlobal(), 3.0 * A @ B + 4.0 * C0, k, A, B)
                if not res <= 1.0:
                    fails.append(f"cannon {dt.__name__} align={al} res={res}")
    # 2b. GemmHost: host-resident local matrices streamed in column bands, against the device Gemm (bit-identical
    #     for alpha = -1, SUMMA_C: the same rank-nb updates in the same order)
    for (oa, ob) in (("N", "N"), ("N", "T"), ("C", "N")):
        m, n, k, nb = 260, 520, 400, 32
        A = O.fill(0, *((m, k) if oa == "N" else (k, m)), 1)
        B = O.fill(0, *((k, n) if ob == "N" else (n, k)), 2)
        C0 = O.fill(0, m, n, 3)
        El.PushBlocksizeStack(nb)
        dA, dB, dC = dm(A), dm(B), dm(C0)
        El.Gemm(ORI[oa], ORI[ob], -1.0, dA, dB, 1.0, dC, El.GEMM_SUMMA_C)
        hA, hB, hC = dA.LocalToHost(), dB.LocalToHost(), dm(C0).LocalToHost()
        El.GemmHost(ORI[oa], ORI[ob], -1.0, g, m, n, k, hA, hB, 1.0, hC, El.GEMM_SUMMA_C)
        El.PopBlocksizeStack()
        if not np.array_equal(hC, dC.LocalToHost()):
            fails.append(f"gemmhost {oa}{ob}")
    # 2c. Syrk with a long summation index (the Dot variants), AxpyTrapezoid across distributions, the Trsm algorithms
    El.SetGemmDotBlocksize(32)
    for uplo in "LU":
        for o in "NC":
            n, k = 70, 900
            A = O.fill(0, *((n, k) if o == "N" else (k, n)), 3, dtype=np.complex128)
            C0 = O.fill(0, n, n, 4, dtype=np.complex128)
            dC = dm(C0)
            El.Herk(0 if uplo == "L" else 1, ORI[o], 1.25, dm(A), 0.5, dC)
            P = A @ A.conj().T if o == "N" else A.conj().T @ A
            mask = O._tri_mask(n, n, uplo)
            want = np.where(mask, 1.25 * P + 0.5 * C0, C0)
            got = dC.ToGlobal()
            if not (np.array_equal(got[~mask], C0[~mask]) and
                    np.linalg.norm(got - want) <= 4 * k * np.finfo(np.float64).eps * (np.linalg.norm(A) ** 2 + np.linalg.norm(C0))):
                fails.append(f"herk dot {uplo}{o}")
    El.SetGemmDotBlocksize(0)
    X, Y0 = O.fill(0, 61, 61, 1), O.fill(0, 61, 61, 2)
    ii, jj = np.indices((61, 61))
    for uplo in "LU":
        for xdist in ((0, 2), (3, 5), (5, 5), (2, 0)):
            dY = dm(Y0, align=(r - 1, c - 1))
            El.AxpyTrapezoid(0 if uplo == "L" else 1, 0.5, dm(X, xdist), dY, 1)
            inside = (jj - ii <= 1) if uplo == "L" else (jj - ii >= 1)
            if not np.array_equal(dY.ToGlobal(), np.where(inside, Y0 + 0.5 * X, Y0)):
                fails.append(f"axpytrapezoid {uplo} {xdist}")
    m, nbt = 150, 32
    T = np.asfortranarray(O.fill(0, m, m, 9, dtype=np.complex128) / m + 2 * np.eye(m))
    for nrhs in (40, 3, 1):
        B0 = O.fill(0, m, nrhs, 10, dtype=np.complex128)
        for uplo in "LU":
            for tr in "NTC":
                ref = (R.trsm("L", uplo, tr, "N", 2.0, T, B0.copy(order="F"), nb=nbt) if R.available()
                       else O.trsm("L", uplo, tr, "N", 2.0, T, B0.copy(order="F"), nb=nbt))
                for alg in (0, 1, 2, 3):
                    dB = dm(B0)
                    El.PushBlocksizeStack(nbt)
                    El.Trsm(0, 0 if uplo == "L" else 1, ORI[tr], 0, 2.0, dm(T), dB, False, alg)
                    El.PopBlocksizeStack()
                    if not np.linalg.norm(dB.ToGlobal() - ref) <= 50 * m * np.finfo(np.float64).eps * np.linalg.norm(ref):
                        fails.append(f"trsm alg={alg} {uplo}{tr} nrhs={nrhs}")
    # 2d. BINARY / BINARY_FLAT files written and read by all ranks (columns through [*,VC])
    import tempfile
    tdir = os.environ.get("ELB200_TEST_TMP", tempfile.gettempdir())
    base = os.path.join(tdir, f"elb200_mgpu_io_{world}_{os.environ.get('MASTER_PORT', '0')}")
    Gio = O.fill(0, 93, 57, 21, dtype=np.complex128)
    dG = dm(Gio, align=(r - 1, 0))
    El.Write(dG, base, El.BINARY)
    El.Write(dG, base, El.BINARY_FLAT)
    dist.barrier()
    raw = np.asfortranarray(Gio).tobytes(order="F")
    if open(base + ".dat", "rb").read() != raw or open(base + ".bin", "rb").read() != np.array([93, 57], dtype=np.int32).tobytes() + raw:
        fails.append("binary file bytes")
    for dd in ((0, 2), (5, 3), (3, 5), (2, 0)):
        Bio = El.DistMatrix(np.complex128, dd[0], dd[1], g)
        El.ReadBinary(Bio, base + ".bin")
        Bf = El.DistMatrix(np.complex128, dd[0], dd[1], g)
        El.ReadBinaryFlat(Bf, 93, 57, base + ".dat")
        if not (np.array_equal(Bio.ToGlobal(), Gio) and np.array_equal(Bf.ToGlobal(), Gio)):
            fails.append(f"binary read {dd}")
    dist.barrier()
    if rank == 0:
        for ext in (".bin", ".dat"):
            try:
                os.remove(base + ext)
            except OSError:
                pass
    # 3. Cholesky / HPDSolve / Trsm
    for dt in (np.float64, np.complex128):
        n, nb = 300, 64
        A = O.fill(1, n, n, 5, diag=float(n), dtype=dt)
        for uplo in "LU":
            dA = dm(A)
            El.PushBlocksizeStack(nb)
            El.Cholesky(0 if uplo == "L" else 1, dA)
            F = dA.ToGlobal()
            res = O.cholesky_residual(uplo, F, A)
            ref = R.cholesky(uplo, A.copy(order="F"), nb=nb) if R.available() else O.cholesky(uplo, A.copy(order="F"), nb)
            tri = O._tri_mask(n, n, uplo)
            dif = np.linalg.norm((F - ref)[tri]) / (n * np.finfo(np.float64).eps * np.linalg.norm(ref))
            if not (res <= 10 and dif <= 20 and np.array_equal(F[~tri], A[~tri])):
                fails.append(f"cholesky {dt.__name__} {uplo} res={res} dif={dif}")
            Bm = O.fill(0, n, 40, 8, dtype=dt)
            dB = dm(Bm)
            El.HPDSolve(0 if uplo == "L" else 1, 0, dm(A), dB)
            El.PopBlocksizeStack()
            X = dB.ToGlobal()
            rs = np.linalg.norm(A @ X - Bm) / (n * np.finfo(np.float64).eps * np.linalg.norm(A) * np.linalg.norm(X))
            if not rs <= 10:
                fails.append(f"hpdsolve {dt.__name__} {uplo} res={rs}")
        try:
            El.Cholesky(0, dm(O.fill(1, 100, 100, 5, diag=0.0, dtype=dt)))
            fails.append("non-HPD did not raise")
        except El.NonHPDMatrixException:
            pass
    for side in "LR":
        m, n, nb = 90, 70, 32
        na = m if side == "L" else n
        A = np.asfortranarray(O.fill(0, na, na, 9) + na * np.eye(na))
        for uplo in "LU":
            for tr in "NT":
                B0 = O.fill(0, m, n, 10)
                dB = dm(B0)
                El.PushBlocksizeStack(nb)
                El.Trsm(0 if side == "L" else 1, 0 if uplo == "L" else 1, ORI[tr], 0, 2.0, dm(A), dB)
                El.PopBlocksizeStack()
                X = dB.ToGlobal()
                T = O._tri(A, uplo, "N")
                opT = T if tr == "N" else T.T
                lhs = opT @ X if side == "L" else X @ opT
                if not np.linalg.norm(lhs - 2 * B0) <= 50 * na * np.finfo(np.float64).eps * np.linalg.norm(T) * np.linalg.norm(X):
                    fails.append(f"trsm {side}{uplo}{tr}")
    # 4. Herk
    A = O.fill(0, 120, 60, 3)
    C0 = O.fill(0, 120, 120, 4)
    dC = dm(C0)
    El.PushBlocksizeStack(32)
    El.Herk(0, 0, -1.0, dm(A), 1.0, dC)
    El.PopBlocksizeStack()
    if not np.linalg.norm(dC.ToGlobal() - O.herk("L", "N", -1.0, A, 1.0, C0.copy())) <= 1e-11:
        fails.append("herk")

    # 5. LU with partial pivoting (the panel is replicated, the interchanges all-gather inside the process column),
    #    lu::SolveAfter, LinearSolve and the permutation applied to distributed matrices; against the numpy
    #    restatement at the same Blocksize() (pinned to the reference on the CPU: tests/test_oracle_cpu.py)
    for dt, (m, n, nb) in ((np.float64, (150, 150, 32)), (np.complex128, (130, 170, 32)), (np.float64, (170, 90, 48))):
        A = O.fill(0, m, n, 31, dtype=dt)
        F = A.copy(order="F")
        pref = O.lu(F, nb)
        dA = dm(A)
        P = El.DistPermutation(g)
        El.PushBlocksizeStack(nb)
        El.LU(dA, P)
        El.PopBlocksizeStack()
        p = P.Preimages()
        if not np.array_equal(p, pref):
            fails.append(f"lu pivots {np.dtype(dt).name} {m}x{n}")
        elif not np.linalg.norm(dA.ToGlobal() - F) <= 50 * max(m, n) * np.finfo(np.float64).eps * np.linalg.norm(A):
            fails.append(f"lu factors {np.dtype(dt).name} {m}x{n}")
        for dist_ in ((0, 2), (3, 5), (5, 4), (2, 0)):
            X0 = O.fill(0, m, 23, 32, dtype=dt)
            dX = dm(X0, dist_)
            P.PermuteRows(dX)
            ok = np.array_equal(dX.ToGlobal(), X0[pref, :])
            P.InversePermuteRows(dX)
            if not (ok and np.array_equal(dX.ToGlobal(), X0)):
                fails.append(f"permute rows {dist_}")
    n, nb = 140, 32
    A = O.fill(0, n, n, 33)
    B0 = O.fill(0, n, 40, 34)
    dF, P = dm(A), El.DistPermutation(g)
    El.PushBlocksizeStack(nb)
    El.LU(dF, P)
    for o in "NT":
        dB = dm(B0)
        El.LUSolveAfter(ORI[o], dF, dB, P)
        X = dB.ToGlobal()
        opA = A if o == "N" else A.T
        if not np.linalg.norm(opA @ X - B0) <= 10 * n * np.finfo(np.float64).eps * np.linalg.norm(A) * np.linalg.norm(X):
            fails.append(f"lu solve {o}")
    dB = dm(B0)
    El.LinearSolve(dm(A), dB)
    X = dB.ToGlobal()
    El.PopBlocksizeStack()
    if not np.linalg.norm(A @ X - B0) <= 10 * n * np.finfo(np.float64).eps * np.linalg.norm(A) * np.linalg.norm(X):
        fails.append("linear solve")

    # 6. CholeskyMod and the diagonally pivoted Cholesky (replicated panel state, slot interchanges in both
    #    communicators) against the numpy restatements
    n, nb = 150, 32
    Gm = O.fill(0, n, n, 41)
    S = Gm @ Gm.T / n + 0.05 * np.eye(n)
    S = np.asfortranarray((S + S.T) / 2)
    L0 = np.linalg.cholesky(S)
    El.PushBlocksizeStack(nb)
    for uplo in "LU":
        T0 = np.asfortranarray(L0 if uplo == "L" else L0.T)
        for w, alpha in ((7, 0.4), (3, -0.0003)):
            V = O.fill(0, n, w, 42 + w)
            dT = dm(T0)
            El.CholeskyMod(0 if uplo == "L" else 1, dT, alpha, dm(V))
            got = dT.ToGlobal()
            Lg = np.tril(got) if uplo == "L" else np.triu(got).T
            target = S + alpha * (V @ V.T)
            if not np.linalg.norm(Lg @ Lg.T - target) <= 10 * n * np.finfo(np.float64).eps * np.linalg.norm(target):
                fails.append(f"cholesky mod {uplo} {alpha}")
        dA, P = dm(S), El.DistPermutation(g)
        El.CholeskyPiv(0 if uplo == "L" else 1, dA, P)
        got, p = dA.ToGlobal(), P.Preimages()
        Fo = S.copy(order="F")
        po = O.cholesky_pivoted(uplo, Fo)
        Lg = np.tril(got) if uplo == "L" else np.triu(got).T
        if not np.array_equal(p, po):
            fails.append(f"pivoted cholesky pivots {uplo}")
        if not np.linalg.norm(S[np.ix_(p, p)] - Lg @ Lg.T) <= 10 * n * np.finfo(np.float64).eps * np.linalg.norm(S):
            fails.append(f"pivoted cholesky {uplo}")
        B0 = O.fill(0, n, 11, 44)
        dB = dm(B0)
        El.CholeskyPivSolveAfter(0 if uplo == "L" else 1, 0, dA, P, dB)
        X = dB.ToGlobal()
        if not np.linalg.norm(S @ X - B0) <= 1e3 * n * np.finfo(np.float64).eps * np.linalg.norm(S) * np.linalg.norm(X):
            fails.append(f"pivoted cholesky solve {uplo}")
    El.PopBlocksizeStack()

    bad = torch.tensor([len(fails)], device="cuda")
    dist.all_reduce(bad)
    if fails:
        print(f"[rank {rank}] FAILURES:\n  " + "\n  ".join(fails[:20]), flush=True)
    if rank == 0:
        print(f"MGPU {'OK' if bad.item() == 0 else 'FAILED'} grid={r}x{c} stats={El.RedistStats()}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if bad.item() == 0 else 1)


if __name__ == "__main__":
    main()
