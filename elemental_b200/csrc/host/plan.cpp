// See plan.hpp.  Pure host arithmetic: residue-class intersections (CRT) per dimension.
#include "plan.hpp"

#include <algorithm>

#include "elb200/level3.hpp"
#include "elb200_plan.h"

namespace El {
namespace plan {

bool PinsRow(Dist d) { return d == MC || d == VC || d == VR; }
bool PinsCol(Dist d) { return d == MR || d == VC || d == VR; }

namespace {
// g = a (mod M) and g = b (mod N): smallest solution x0 >= 0 and period L, or false.
bool Crt(Int a, Int M, Int b, Int N, Int& x0, Int& L) {
    Int g = M, h = N;
    while (h) { Int t = g % h; g = h; h = t; }
    if ((a - b) % g != 0) return false;
    L = M / g * N;
    x0 = a;
    while (x0 % N != b) x0 += M;  // at most N/g steps
    return true;
}

struct Lat1D {
    bool empty = true;
    Int count = 0, src0 = 0, srcStep = 0, dst0 = 0, dstStep = 0;
};

// indices in [0,n) owned by rank (si,sj) under (sd,sAlign) and by (di,dj) under (dd,dAlign)
Lat1D Intersect(Int n, const Grid& g, Dist sd, int sAlign, int si, int sj, Dist dd, int dAlign, int di, int dj) {
    Lat1D L;
    const Int S1 = DistStride(sd, g), S2 = DistStride(dd, g);
    const Int sh1 = Shift_(DistRankOf(sd, g, si, sj), sAlign, S1);
    const Int sh2 = Shift_(DistRankOf(dd, g, di, dj), dAlign, S2);
    Int x0, per;
    if (!Crt(sh1, S1, sh2, S2, x0, per)) return L;
    L.count = Length_(n, x0, per);
    if (L.count <= 0) return L;
    L.empty = false;
    L.src0 = (x0 - sh1) / S1; L.srcStep = per / S1;
    L.dst0 = (x0 - sh2) / S2; L.dstStep = per / S2;
    return L;
}
}  // namespace

Msg ComputeMsg(const Grid& g, Int height, Int width, const Layout& A, i64 ldA, int si, int sj, const Layout& B,
               i64 ldB, int di, int dj, bool transpose, bool chooseOwner) {
    Msg m;
    if (chooseOwner) {
        const bool pinsI = PinsRow(A.U) || PinsRow(A.V);
        const bool pinsJ = PinsCol(A.U) || PinsCol(A.V);
        if (!pinsI && si != di) return m;
        if (!pinsJ && sj != dj) return m;
    }
    Lat1D rows, cols;
    if (!transpose) {
        rows = Intersect(height, g, A.U, A.colAlign, si, sj, B.U, B.colAlign, di, dj);
        cols = Intersect(width, g, A.V, A.rowAlign, si, sj, B.V, B.rowAlign, di, dj);
    } else {
        // B(g,h) = A(h,g): destination rows run over A's columns and vice versa
        rows = Intersect(height, g, A.V, A.rowAlign, si, sj, B.U, B.colAlign, di, dj);
        cols = Intersect(width, g, A.U, A.colAlign, si, sj, B.V, B.rowAlign, di, dj);
    }
    if (rows.empty || cols.empty) return m;
    m.empty = false;
    m.nrows = rows.count; m.ncols = cols.count;
    if (!transpose) {
        m.s_off = rows.src0 + (i64)cols.src0 * ldA; m.s_rs = rows.srcStep; m.s_cs = (i64)cols.srcStep * ldA;
    } else {
        m.s_off = cols.src0 + (i64)rows.src0 * ldA; m.s_rs = (i64)rows.srcStep * ldA; m.s_cs = cols.srcStep;
    }
    m.d_off = rows.dst0 + (i64)cols.dst0 * ldB; m.d_rs = rows.dstStep; m.d_cs = (i64)cols.dstStep * ldB;
    return m;
}

RedistPlan BuildRedistPlan(const Grid& g, Int h, Int w, const Layout& A, i64 ldA, const Layout& B, i64 ldB,
                           bool transpose) {
    const int r = g.Height(), p = g.Size();
    const int mi = g.Row(), mj = g.Col();
    RedistPlan P;
    P.send.resize(p);
    P.recv.resize(p);
    for (int v = 0; v < p; ++v) {
        const int qi = v % r, qj = v / r;
        P.send[v] = ComputeMsg(g, h, w, A, ldA, mi, mj, B, ldB, qi, qj, transpose, true);
        if (!(qi == mi && qj == mj)) P.recv[v] = ComputeMsg(g, h, w, A, ldA, qi, qj, B, ldB, mi, mj, transpose, true);
    }
    return P;
}

ContractPlan BuildContractPlan(const Grid& g, Int h, Int w, const Layout& A, i64 ldA, const Layout& bView) {
    ContractPlan P;
    const Dist U = A.U, V = A.V;
    const bool pinsI = PinsRow(U) || PinsRow(V), pinsJ = PinsCol(U) || PinsCol(V);
    if (pinsI && pinsJ) return P;  // NOT_REPLICATED
    const int r = g.Height(), mi = g.Row(), mj = g.Col();
    bool freeCol = false, freeRow = false;
    int np;
    if (!pinsI && !pinsJ) {  // [*,*]: summed over every rank
        P.kind = OVER_VC; np = g.Size();
        P.T = (bView.U == MR && bView.V == MC) ? Layout{MR, MC, 0, 0} : Layout{MC, MR, 0, 0};
        freeCol = freeRow = true;
    } else if (pinsI) {  // replicated along the grid row: scatter over the row communicator
        P.kind = OVER_MR; np = g.Width();
        if (U == MC) { P.T = Layout{MC, MR, A.colAlign, 0}; freeRow = true; }   // [MC,*]
        else { P.T = Layout{MR, MC, 0, A.rowAlign}; freeCol = true; }             // [*,MC]
    } else {  // replicated along the grid column: scatter over the column communicator
        P.kind = OVER_MC; np = g.Height();
        if (V == MR) { P.T = Layout{MC, MR, 0, A.rowAlign}; freeCol = true; }   // [*,MR]
        else { P.T = Layout{MR, MC, A.colAlign, 0}; freeRow = true; }             // [MR,*]
    }
    // let the free alignment of T coincide with B's whenever B has T's distribution there
    if (freeCol && bView.U == P.T.U) P.T.colAlign = bView.colAlign;
    if (freeRow && bView.V == P.T.V) P.T.rowAlign = bView.rowAlign;
    const Int SU = DistStride(P.T.U, g), SV = DistStride(P.T.V, g);
    P.chunk = std::max<i64>((i64)MaxLength(h, SU) * (i64)MaxLength(w, SV), 1);
    P.packs.resize(np);
    for (int q = 0; q < np; ++q) {
        int qi, qj;
        if (P.kind == OVER_MR) { qi = mi; qj = q; }
        else if (P.kind == OVER_MC) { qi = q; qj = mj; }
        else { qi = q % r; qj = q / r; }
        const i64 rowsQ = Length_(h, Shift_(DistRankOf(P.T.U, g, qi, qj), P.T.colAlign, SU), SU);
        if (qi == mi && qj == mj) P.myRows = rowsQ;
        // my (replicated) copy of everything member q owns in T, laid out as q's local T
        P.packs[q] = ComputeMsg(g, h, w, A, ldA, mi, mj, P.T, std::max<i64>(rowsQ, 1), qi, qj, false, false);
        if (P.packs[q].empty || P.packs[q].count() < P.chunk) P.needZero = true;
    }
    return P;
}

}  // namespace plan
}  // namespace El

// ---------------------------------------------------------------------------
// C-ABI (include/elb200_plan.h)
// ---------------------------------------------------------------------------
namespace {
using namespace El;
plan::Layout L(const elb200_layout& l) {
    return plan::Layout{(Dist)l.colDist, (Dist)l.rowDist, l.colAlign, l.rowAlign};
}
void Fill(elb200_plan_msg& o, int kind, int pi, int pj, const plan::Msg& m) {
    o.kind = kind; o.peerRow = pi; o.peerCol = pj;
    o.nrows = m.empty ? 0 : m.nrows; o.ncols = m.empty ? 0 : m.ncols;
    o.s_off = m.s_off; o.s_rs = m.s_rs; o.s_cs = m.s_cs;
    o.d_off = m.d_off; o.d_rs = m.d_rs; o.d_cs = m.d_cs;
}
template <class F>
int Guard(F&& f) {
    try { f(); return 0; } catch (...) { return 1; }
}
}  // namespace

extern "C" {

int elb200_redist_plan(int r, int c, int myRow, int myCol, int64_t height, int64_t width, elb200_layout A,
                       int64_t ldA, elb200_layout B, int64_t ldB, int transpose, elb200_plan_msg* out, int* nout) {
    return Guard([&] {
        Grid g(r, c, myRow, myCol, Grid::PlanningOnly());
        plan::RedistPlan P = plan::BuildRedistPlan(g, (Int)height, (Int)width, L(A), ldA, L(B), ldB, transpose != 0);
        int n = 0;
        const int me = myRow + r * myCol;
        for (int v = 0; v < r * c; ++v) {
            if (!P.send[v].empty) Fill(out[n++], v == me ? 2 : 0, v % r, v / r, P.send[v]);
            if (!P.recv[v].empty) Fill(out[n++], 1, v % r, v / r, P.recv[v]);
        }
        *nout = n;
    });
}

int elb200_contract_plan(int r, int c, int myRow, int myCol, int64_t height, int64_t width, elb200_layout A,
                         int64_t ldA, elb200_layout bLayout, int* commKind, elb200_layout* T, int64_t* chunk,
                         elb200_plan_msg* packs, int* npacks) {
    return Guard([&] {
        Grid g(r, c, myRow, myCol, Grid::PlanningOnly());
        plan::ContractPlan P = plan::BuildContractPlan(g, (Int)height, (Int)width, L(A), ldA, L(bLayout));
        *commKind = (int)P.kind;
        T->colDist = P.T.U; T->rowDist = P.T.V; T->colAlign = P.T.colAlign; T->rowAlign = P.T.rowAlign;
        *chunk = P.chunk;
        *npacks = (int)P.packs.size();
        for (size_t q = 0; q < P.packs.size(); ++q) Fill(packs[q], (int)q, -1, -1, P.packs[q]);
    });
}

int64_t elb200_shift(int64_t rank, int64_t align, int64_t stride) { return Shift_((Int)rank, (Int)align, (Int)stride); }
int64_t elb200_length(int64_t n, int64_t shift, int64_t stride) { return Length_((Int)n, (Int)shift, (Int)stride); }
int elb200_dist_stride(int dist, int r, int c) {
    int out = -1;
    Guard([&] { Grid g(r, c, 0, 0, Grid::PlanningOnly()); out = DistStride((Dist)dist, g); });
    return out;
}
int elb200_dist_rank(int dist, int r, int c, int row, int col) {
    int out = -1;
    Guard([&] { Grid g(r, c, row, col, Grid::PlanningOnly()); out = DistRank((Dist)dist, g); });
    return out;
}
int elb200_gemm_default_algorithm(int64_t m, int64_t n, int64_t k) {
    return (int)GemmDefaultAlgorithm((Int)m, (Int)n, (Int)k);
}

}  // extern "C"

extern "C" {
int elb200_perm_compose(int64_t size, int64_t nswaps, const int64_t* origins, const int64_t* dests, int64_t* pre,
                        int64_t* img) {
    if (size < 0 || nswaps < 0) return 1;
    for (int64_t i = 0; i < size; ++i) pre[i] = i;
    for (int64_t j = 0; j < nswaps; ++j) {
        const int64_t o = origins[j], d = dests[j];
        if (o < 0 || o >= size || d < 0 || d >= size) return 1;
        const int64_t t = pre[o];   // the rows trade places, so do their labels
        pre[o] = pre[d];
        pre[d] = t;
    }
    if (img)
        for (int64_t i = 0; i < size; ++i) img[pre[i]] = i;
    return 0;
}
int elb200_perm_parity(int64_t size, const int64_t* pre) {
    // a permutation with c cycles is a product of size - c transpositions
    std::vector<char> seen((size_t)(size > 0 ? size : 0), 0);
    int64_t cycles = 0;
    for (int64_t i = 0; i < size; ++i) {
        if (seen[(size_t)i]) continue;
        ++cycles;
        for (int64_t j = i; !seen[(size_t)j]; j = pre[j]) seen[(size_t)j] = 1;
    }
    return (int)((size - cycles) & 1);
}
}

