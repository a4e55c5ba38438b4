"""Helpers shared by the CPU tests of the redistribution planner (test infrastructure)."""
import ctypes as C

import numpy as np

from elemental_b200._lib import lib

MC, MD, MR, VC, VR, STAR = 0, 1, 2, 3, 4, 5
NAMES = {MC: "MC", MR: "MR", VC: "VC", VR: "VR", STAR: "STAR"}
# the distribution pairs the hot path touches
LEGAL = [(MC, MR), (MC, STAR), (STAR, MR), (MR, MC), (MR, STAR), (STAR, MC), (VC, STAR), (VR, STAR),
         (STAR, VC), (STAR, VR), (STAR, STAR)]


class Layout(C.Structure):
    _fields_ = [("colDist", C.c_int), ("rowDist", C.c_int), ("colAlign", C.c_int), ("rowAlign", C.c_int)]


class PlanMsg(C.Structure):
    _fields_ = [("kind", C.c_int), ("peerRow", C.c_int), ("peerCol", C.c_int),
                ("nrows", C.c_int64), ("ncols", C.c_int64),
                ("s_off", C.c_int64), ("s_rs", C.c_int64), ("s_cs", C.c_int64),
                ("d_off", C.c_int64), ("d_rs", C.c_int64), ("d_cs", C.c_int64)]


def redist_plan(r, c, row, col, h, w, A, ldA, B, ldB, transpose):
    L = lib()
    out = (PlanMsg * (2 * r * c))()
    n = C.c_int()
    rc = L.elb200_redist_plan(r, c, row, col, C.c_int64(h), C.c_int64(w), A, C.c_int64(ldA), B, C.c_int64(ldB),
                              int(transpose), out, C.byref(n))
    assert rc == 0
    return [out[i] for i in range(n.value)]


def contract_plan(r, c, row, col, h, w, A, ldA, Bview):
    L = lib()
    packs = (PlanMsg * max(r * c, 1))()
    kind, T, chunk, n = C.c_int(), Layout(), C.c_int64(), C.c_int()
    rc = L.elb200_contract_plan(r, c, row, col, C.c_int64(h), C.c_int64(w), A, C.c_int64(ldA), Bview,
                                C.byref(kind), C.byref(T), C.byref(chunk), packs, C.byref(n))
    assert rc == 0
    return kind.value, T, chunk.value, [packs[i] for i in range(n.value)]


def gather_lattice(flat, m, src=True):
    """elements of message m from a flat (column-major, padded) local buffer, shape (nrows, ncols)"""
    t = np.arange(m.nrows)[:, None]
    u = np.arange(m.ncols)[None, :]
    idx = (m.s_off + t * m.s_rs + u * m.s_cs) if src else (m.d_off + t * m.d_rs + u * m.d_cs)
    return flat[idx]


def scatter_lattice(flat, m, vals, accumulate=False):
    t = np.arange(m.nrows)[:, None]
    u = np.arange(m.ncols)[None, :]
    idx = m.d_off + t * m.d_rs + u * m.d_cs
    if accumulate:
        flat[idx] += vals
    else:
        flat[idx] = vals


def local_flat(local, ld):
    """column-major flat buffer with leading dimension ld holding `local`"""
    lh, lw = local.shape
    buf = np.full(ld * max(lw, 1), np.nan, dtype=local.dtype)
    for j in range(lw):
        buf[j * ld:j * ld + lh] = local[:, j]
    return buf


def flat_local(buf, lh, lw, ld):
    out = np.empty((lh, lw), dtype=buf.dtype)
    for j in range(lw):
        out[:, j] = buf[j * ld:j * ld + lh]
    return out
