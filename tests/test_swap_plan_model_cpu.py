"""CPU model check of the slot scheme that applies a panel's row interchanges in one pack / all-gather / unpack
(elemental_b200/csrc/kernels/lu.cu: swap_plan_kernel, pack_rows_kernel, unpack_rows_kernel; the column form in
cholpiv.cu).  The kernels are restated line by line in numpy and compared with the definition -- swap rows j and
ipiv[j] for j = 0 .. nb-1, in order -- on random swap sequences that contain repeated destinations, destinations
inside the block and fixed points, for rows dealt out cyclically over r processes."""
import numpy as np
import pytest


def swap_plan(nb, ipiv, k):
    """slotRow[2 nb] (global row or -1) and srcSlot[2 nb] (the slot whose ORIGINAL row ends up in this slot's row)"""
    partner = np.zeros(nb, dtype=np.int64)
    slot_row = np.full(2 * nb, -1, dtype=np.int64)
    for j in range(nb):
        p = ipiv[j]
        if p < nb:
            slot = p
        else:
            first = j
            for q in range(j):
                if ipiv[q] == p:
                    first = q
                    break
            slot = nb + first
        partner[j] = slot
        slot_row[j] = k + j
        slot_row[nb + j] = k + p if (p >= nb and slot == nb + j) else -1
    lab = np.arange(2 * nb)
    for j in range(nb):
        b = partner[j]
        lab[j], lab[b] = lab[b], lab[j]
    return slot_row, lab


@pytest.mark.parametrize("r", [1, 2, 3])
def test_slot_interchange_equals_sequential_swaps(r):
    rng = np.random.default_rng(5 + r)
    for trial in range(60):
        nb = int(rng.integers(1, 12))
        m = nb + int(rng.integers(0, 30))
        k = int(rng.integers(0, 7))
        total = k + m
        # destinations as an LU panel produces them: ipiv[j] >= j, relative to row k
        ipiv = np.array([int(rng.integers(j, m)) for j in range(nb)], dtype=np.int64)
        if trial % 3 == 0 and m > nb:      # force repeats and far rows
            ipiv[: nb // 2 + 1] = m - 1
        A = rng.standard_normal((total, 5))
        want = A.copy()
        for j in range(nb):
            a, b = k + j, k + int(ipiv[j])
            want[[a, b]] = want[[b, a]]
        slot_row, src_slot = swap_plan(nb, ipiv, k)
        S = 2 * nb
        align = int(rng.integers(0, r))
        owner = lambda row: (row + align) % r
        # pack: every process contributes the ORIGINAL content of the slot rows it owns
        bufs = np.full((r, S, A.shape[1]), np.nan)
        for rank in range(r):
            for slot in range(S):
                row = slot_row[slot]
                if row >= 0 and owner(row) == rank:
                    bufs[rank, slot] = A[row]
        got = A.copy()
        # unpack: the owner of a slot's row takes the source slot's row from its owner's contribution
        for rank in range(r):
            for slot in range(S):
                row = slot_row[slot]
                if row < 0 or owner(row) != rank or src_slot[slot] == slot:
                    continue
                src = src_slot[slot]
                got[row] = bufs[owner(slot_row[src]), src]
        assert not np.isnan(got).any()
        assert np.array_equal(got, want), (trial, nb, m, k, ipiv)
        # every global row appears in at most one valid slot
        valid = slot_row[slot_row >= 0]
        assert len(set(valid.tolist())) == len(valid)
