"""CPU tests of the redistribution planner (host logic, no GPU): every rank of an r x c grid
is simulated in one process and the plans returned by the C-ABI are executed with numpy.
The expected result is the oracle's definition of each distribution
(oracle.elemental_oracle.local_part <- src/core/DistMatrix/Element.cpp:544-600), the same
entrywise check as the reference's tests/core/DistMatrix.cpp:12-75."""
import itertools

import numpy as np
import pytest

from oracle import elemental_oracle as O
from planutil import (LEGAL, MC, MR, NAMES, STAR, VC, VR, Layout, contract_plan, flat_local, gather_lattice,
                      local_flat, redist_plan, scatter_lattice)


def _pieces(G, lay, r, c):
    return {(i, j): O.local_part(G, NAMES[lay.colDist], NAMES[lay.rowDist], lay.colAlign, lay.rowAlign, r, c, i, j)
            for i in range(r) for j in range(c)}


def _run_redist(G, la, lb, r, c, transpose, conj=False):
    h, w = (G.shape[1], G.shape[0]) if transpose else G.shape
    srcs = _pieces(G, la, r, c)
    ld_a = {k: max(v.shape[0], 1) + 1 for k, v in srcs.items()}          # padded on purpose
    flats_a = {k: local_flat(v, ld_a[k]) for k, v in srcs.items()}
    Gd = G.T if transpose else G
    want = _pieces(Gd, lb, r, c)
    ld_b = {k: max(v.shape[0], 1) + 2 for k, v in want.items()}
    flats_b = {k: np.full(ld_b[k] * max(v.shape[1], 1), np.nan) for k, v in want.items()}
    mailbox = {}
    plans = {}
    for i in range(r):
        for j in range(c):
            plans[(i, j)] = redist_plan(r, c, i, j, h, w, la, ld_a[(i, j)], lb, ld_b[(i, j)], transpose)
            for m in plans[(i, j)]:
                if m.kind == 0:
                    mailbox[((i, j), (m.peerRow, m.peerCol))] = gather_lattice(flats_a[(i, j)], m, src=True)
                elif m.kind == 2:
                    scatter_lattice(flats_b[(i, j)], m, gather_lattice(flats_a[(i, j)], m, src=True))
    n_recv = 0
    for (i, j), plan in plans.items():
        for m in plan:
            if m.kind == 1:
                vals = mailbox.pop(((m.peerRow, m.peerCol), (i, j)))
                assert vals.shape == (m.nrows, m.ncols)
                scatter_lattice(flats_b[(i, j)], m, vals)
                n_recv += 1
    assert not mailbox, "a message was sent that nobody receives"
    for k, v in want.items():
        got = flat_local(flats_b[k], v.shape[0], v.shape[1], ld_b[k])
        np.testing.assert_array_equal(got, v)
    return n_recv


GRIDS = [(1, 1), (1, 2), (2, 1), (2, 2), (2, 4), (3, 2)]


@pytest.mark.parametrize("r,c", GRIDS)
def test_all_pairs_aligned(r, c):
    rng = np.random.default_rng(1)
    G = rng.standard_normal((13, 11))
    for (u, v), (u2, v2) in itertools.product(LEGAL, LEGAL):
        la, lb = Layout(u, v, 0, 0), Layout(u2, v2, 0, 0)
        _run_redist(G, la, lb, r, c, False)


@pytest.mark.parametrize("r,c", [(2, 2), (2, 4), (3, 2)])
def test_all_pairs_random_alignments_and_transpose(r, c):
    rng = np.random.default_rng(2)
    p = r * c
    stride = {MC: r, MR: c, VC: p, VR: p, STAR: 1}
    for trial in range(3):
        G = rng.standard_normal((int(rng.integers(1, 20)), int(rng.integers(1, 20))))
        for (u, v), (u2, v2) in itertools.product(LEGAL, LEGAL):
            la = Layout(u, v, int(rng.integers(stride[u])), int(rng.integers(stride[v])))
            lb = Layout(u2, v2, int(rng.integers(stride[u2])), int(rng.integers(stride[v2])))
            _run_redist(G, la, lb, r, c, transpose=bool(trial % 2))


def test_filters_are_local():
    """[*,*]->[MC,MR], [MC,*]->[MC,MR], [*,MR]->[MC,MR], [MC,*]->[VC,*] need no communication when
    aligned (copy::Filter / RowFilter / ColFilter / PartialColFilter in the reference)."""
    G = np.arange(17 * 9, dtype=float).reshape(17, 9)
    for la, lb in [(Layout(STAR, STAR, 0, 0), Layout(MC, MR, 1, 2)), (Layout(MC, STAR, 1, 0), Layout(MC, MR, 1, 3)),
                   (Layout(STAR, MR, 0, 2), Layout(MC, MR, 0, 2)), (Layout(MC, STAR, 1, 0), Layout(VC, STAR, 1, 0))]:
        assert _run_redist(G, la, lb, 2, 4, False) == 0


def test_empty_and_tiny():
    for shape in [(0, 5), (5, 0), (1, 1), (1, 9), (9, 1)]:
        G = np.ones(shape)
        for (u, v), (u2, v2) in itertools.product(LEGAL[:6], LEGAL[:6]):
            _run_redist(G, Layout(u, v, 0, 0), Layout(u2, v2, 0, 0), 2, 2, False)


def _run_contract(r, c, la, lb, h, w, seed=0):
    """every replica of A holds a DIFFERENT partial sum; B must end up with the sum, distributed as lb"""
    rng = np.random.default_rng(seed)
    parts = {(i, j): rng.standard_normal((h, w)) for i in range(r) for j in range(c)}
    # which ranks hold copies of the same entries: those sharing the pinned coordinates
    pins_i = la.colDist in (MC, VC, VR) or la.rowDist in (MC, VC, VR)
    pins_j = la.colDist in (MR, VC, VR) or la.rowDist in (MR, VC, VR)
    total = {}
    for (i, j) in parts:
        group = [(a, b) for (a, b) in parts if (not pins_i or a == i) and (not pins_j or b == j)]
        total[(i, j)] = sum(parts[g] for g in group)
    srcs = {k: O.local_part(parts[k], NAMES[la.colDist], NAMES[la.rowDist], la.colAlign, la.rowAlign, r, c, *k)
            for k in parts}
    ld_a = {k: max(v.shape[0], 1) for k, v in srcs.items()}
    flats_a = {k: local_flat(v, ld_a[k]) for k, v in srcs.items()}
    sendbufs, metas = {}, {}
    for k in parts:
        kind, T, chunk, packs = contract_plan(r, c, k[0], k[1], h, w, la, ld_a[k], lb)
        assert kind in (0, 1, 2)
        buf = np.zeros(chunk * len(packs))
        for q, m in enumerate(packs):
            if m.nrows:
                sub = buf[q * chunk:(q + 1) * chunk]
                scatter_lattice(sub, m, gather_lattice(flats_a[k], m, src=True))
        sendbufs[k], metas[k] = buf, (kind, T, chunk, len(packs))
    for k in parts:
        kind, T, chunk, np_ = metas[k]
        if kind == 0:
            members = [(k[0], q) for q in range(c)]; me = k[1]
        elif kind == 1:
            members = [(q, k[1]) for q in range(r)]; me = k[0]
        else:
            members = [(q % r, q // r) for q in range(r * c)]; me = k[0] + r * k[1]
        red = sum(sendbufs[m] for m in members)[me * chunk:(me + 1) * chunk]
        # my piece of T of the group's total
        want = O.local_part(total[k], NAMES[T.colDist], NAMES[T.rowDist], T.colAlign, T.rowAlign, r, c, *k)
        got = flat_local(red, want.shape[0], want.shape[1], max(want.shape[0], 1))
        np.testing.assert_allclose(got, want, rtol=0, atol=1e-12)


@pytest.mark.parametrize("r,c", [(1, 2), (2, 2), (2, 4), (3, 2)])
def test_contract_plans(r, c):
    for (u, v) in [(MC, STAR), (STAR, MR), (MR, STAR), (STAR, MC), (STAR, STAR)]:
        for (bu, bv) in [(MC, MR), (MR, MC)]:
            for h, w in [(13, 7), (1, 5), (4, 4)]:
                ca = 1 % (r if u == MC else (c if u == MR else 1))
                ra = 1 % (c if v == MR else (r if v == MC else 1))
                _run_contract(r, c, Layout(u, v, ca, ra), Layout(bu, bv, (1 % r) if bu == MC else (1 % c), 0), h, w)


def test_permutation_bookkeeping_matches_sequential_swaps():
    """elb200_perm_compose / elb200_perm_parity (the host logic of El::DistPermutation, Permutation.cpp:333-347):
    swaps applied in order to the identity give the preimage vector; images invert it; parity = sign of the
    permutation (determinant of the permutation matrix)."""
    import ctypes as C
    from elemental_b200._lib import lib
    L = lib()
    rng = np.random.default_rng(77)
    for trial in range(40):
        n = int(rng.integers(1, 40))
        k = int(rng.integers(0, 60))
        o = rng.integers(0, n, k).astype(np.int64)
        d = rng.integers(0, n, k).astype(np.int64)
        pre = np.zeros(n, dtype=np.int64)
        img = np.zeros(n, dtype=np.int64)
        P64 = C.POINTER(C.c_int64)
        rc = L.elb200_perm_compose(C.c_int64(n), C.c_int64(k), o.ctypes.data_as(P64), d.ctypes.data_as(P64),
                                   pre.ctypes.data_as(P64), img.ctypes.data_as(P64))
        assert rc == 0
        want = np.arange(n)
        for a, b in zip(o, d):
            want[[a, b]] = want[[b, a]]
        assert np.array_equal(pre, want)
        assert np.array_equal(img[pre], np.arange(n))
        Pm = np.eye(n)[pre]                       # row i of P A is row pre[i] of A
        assert L.elb200_perm_parity(C.c_int64(n), pre.ctypes.data_as(P64)) == (0 if round(np.linalg.det(Pm)) == 1 else 1)
    bad = np.array([5], dtype=np.int64)
    assert L.elb200_perm_compose(C.c_int64(3), C.c_int64(1), bad.ctypes.data_as(P64), bad.ctypes.data_as(P64),
                                 pre.ctypes.data_as(P64), img.ctypes.data_as(P64)) == 1
