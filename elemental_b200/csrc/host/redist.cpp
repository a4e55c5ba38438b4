// The redistribution engine: B (+)= alpha op(A) between ANY two element-cyclic
// distributions of the same Grid, and the sum-scatter (Contract) family.
//
// Replaces, with one code path, the reference's pairwise primitives
//   copy::{RowAllGather, ColAllGather, AllGather, Filter, ColFilter, RowFilter,
//          PartialColFilter, PartialColAllGather, PartialRowAllGather,
//          ColAllToAllDemote/Promote, RowAllToAllDemote/Promote,
//          Colwise/RowwiseVectorExchange, TransposeDist, Translate}
//   (include/El/blas_like/level1/Copy/*.hpp, dispatched from
//    src/core/DistMatrix/Element/*.cpp operator=), transpose::* (level1/Transpose/*.hpp)
//   and axpy_contract::{RowScatter, ColScatter, Scatter} (level1/AxpyContract.hpp:132-432).
//
// Idea: under an element-cyclic distribution the set of global rows a rank owns is a
// residue class; the rows rank s owns in A AND rank d needs in B is the intersection of
// two residue classes, i.e. (by CRT) empty or again a residue class.  So every
// (source rank, destination rank) message is a 2-D strided lattice of the local
// matrices.  One fused kernel launch packs all outgoing lattices (transposing on the
// fly), one grouped NCCL send/recv over NVLink moves them, one launch unpacks
// (applying conj / alpha / accumulate).  Contiguous lattices are sent / received in
// place with no staging copy; identical outgoing lattices (the AllGather pattern) are
// packed once.  When the source is replicated, the destination pulls from the replica
// in its own grid row / column, which makes every "filter" redistribution purely local.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <map>
#include <tuple>

#include "dev.hpp"
#include "plan.hpp"

namespace El {

RedistStats& GetRedistStats() {
    static RedistStats s;
    return s;
}

namespace {

using plan::i64;
using plan::Layout;
using plan::Msg;
using plan::PinsRow;
using plan::PinsCol;
template <typename T>
Layout LayoutOf(const AbstractDistMatrix<T>& A) { return Layout{A.ColDist(), A.RowDist(), A.ColAlign(), A.RowAlign()}; }

bool Contig(i64 rs, i64 cs, i64 nrows, i64 ncols) { return rs == 1 && (ncols == 1 || cs == nrows); }

template <typename T>
void LaunchLattices(std::vector<elb200_lattice>& v, bool conj, const T* alpha, bool acc) {
    if (v.empty()) return;
    dev::D<T> a = dev::val<T>(alpha ? *alpha : T(1));
    dev::c_check(elb200_lattice_copy(dev::Code<T>(), v.data(), (int)v.size(), conj ? 1 : 0, alpha ? &a : nullptr,
                                     acc ? 1 : 0, (elb200_stream_t)dev::stream()),
                 "elb200_lattice_copy");
    GetRedistStats().packLaunches++;
}

// Adopt the source's alignment in every unconstrained dimension of B whose distribution is
// compatible with it (what each copy::* primitive does through AlignColsAndResize etc.).
template <typename T>
void AdoptAlignments(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B, bool transpose) {
    if (B.Viewing()) return;
    const Dist aU = transpose ? A.RowDist() : A.ColDist(), aV = transpose ? A.ColDist() : A.RowDist();
    const int aCA = transpose ? A.RowAlign() : A.ColAlign(), aRA = transpose ? A.ColAlign() : A.RowAlign();
    int ca = B.ColAlign(), ra = B.RowAlign();
    auto pick = [&](Dist mine, int stride, int cur) {
        if (mine == STAR) return 0;
        if (aU == mine || aU == PartialDist(mine)) return aCA % stride;
        if (aV == mine || aV == PartialDist(mine)) return aRA % stride;
        if (aU == PartialUnionDist(mine)) return aCA % stride;
        if (aV == PartialUnionDist(mine)) return aRA % stride;
        return cur;
    };
    if (!B.ColConstrained()) ca = pick(B.ColDist(), B.ColStride(), ca);
    if (!B.RowConstrained()) ra = pick(B.RowDist(), B.RowStride(), ra);
    if (ca != B.ColAlign()) B.AlignCols(ca, false);
    if (ra != B.RowAlign()) B.AlignRows(ra, false);
}

ncclDataType_t NcclReal(size_t realBytes) { return realBytes == 4 ? ncclFloat : ncclDouble; }

// B (+)= alpha * op(A)
template <typename T>
void Redistribute(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B, bool transpose, bool conj, T alpha,
                  bool accumulate) {
    const Grid& g = A.Grid();
    if (&B.Grid() != &g) LogicError("Redistribution requires both matrices on the same Grid");
    const Int h = transpose ? A.Width() : A.Height(), w = transpose ? A.Height() : A.Width();
    if (accumulate) {
        if (B.Height() != h || B.Width() != w) LogicError("Nonconformal Axpy");
    } else {
        AdoptAlignments(A, B, transpose);
        B.Resize(h, w);
    }
    RedistStats& st = GetRedistStats();
    st.copies++;
    if (h == 0 || w == 0) return;
    const bool doConj = conj && IsComplex<T>::value;
    const bool plain = !doConj && alpha == T(1) && !accumulate;
    const Layout la = LayoutOf(A), lb = LayoutOf(B);
    const i64 ldA = A.LDim(), ldB = B.LDim();
    const int r = g.Height(), c = g.Width(), p = r * c;
    const int mi = g.Row(), mj = g.Col();
    cudaStream_t s = dev::stream();
    const T* Abuf = A.LockedBuffer();
    T* Bbuf = B.Buffer();

    std::vector<elb200_lattice> packs, unpacks;
    struct Wire { int peer; int v; const void* sendPtr; void* recvPtr; size_t sendBytes, recvBytes; };
    std::vector<Wire> wires;

    // ---- peer-memory path (Grid::P2PState): one channel per stream that issues redistributions ----
    Grid::P2PState& pp = g.P2P();
    int ch = -1;
    if (pp.on && p > 1) {
        if (*pp.error) RuntimeError("peer-memory redistribution: a flag wait timed out (ranks out of step?)");
        ch = (s == elb200::aux_stream(0)) ? 1 : 0;
    }
    const unsigned ep = ch >= 0 ? ++pp.epoch[ch] : 0u;
    const int meW = g.WorldRank();
    std::vector<int> p2pDest, p2pSrc;  // world ranks this epoch pushes to / is pushed from
    struct Push { const void* src; void* dst; size_t bytes; };
    std::vector<Push> pushes;          // contiguous pieces: copy engine, no SM

    // ---- plan ----
    plan::RedistPlan P = plan::BuildRedistPlan(g, h, w, la, ldA, lb, ldB, transpose);
    std::vector<Msg>& sendMsg = P.send;
    std::vector<Msg>& recvMsg = P.recv;
    i64 packElems = 0, recvElems = 0;
    std::map<std::tuple<i64, i64, i64, i64, i64>, i64> packOffset;  // identical lattices share a slot
    std::vector<i64> sendOff(p, -1), recvOff(p, -1);
    for (int v = 0; v < p; ++v) {
        const int qi = v % r, qj = v / r;
        if (qi == mi && qj == mj) continue;
        const Msg& sm = sendMsg[v];
        const bool sendP2P = ch >= 0 && !sm.empty && sizeof(T) * (size_t)sm.count() <= pp.regionBytes;
        if (!sm.empty && !sendP2P && !Contig(sm.s_rs, sm.s_cs, sm.nrows, sm.ncols)) {
            auto key = std::make_tuple(sm.s_off, sm.s_rs, sm.s_cs, sm.nrows, sm.ncols);
            auto it = packOffset.find(key);
            if (it == packOffset.end()) {
                packOffset[key] = packElems;
                sendOff[v] = packElems;
                packElems += sm.count();
            } else {
                sendOff[v] = it->second;
            }
        }
        const Msg& rm = recvMsg[v];
        const bool recvP2P = ch >= 0 && !rm.empty && sizeof(T) * (size_t)rm.count() <= pp.regionBytes;
        if (!rm.empty && !recvP2P && !(plain && Contig(rm.d_rs, rm.d_cs, rm.nrows, rm.ncols))) {
            recvOff[v] = recvElems;
            recvElems += rm.count();
        }
    }
    T* packBuf = packElems ? (T*)elb200::scratch_alloc(sizeof(T) * (size_t)packElems, s) : nullptr;
    T* recvBuf = recvElems ? (T*)elb200::scratch_alloc(sizeof(T) * (size_t)recvElems, s) : nullptr;

    // ---- pack (pure strided gather, transposing if needed) ----
    {
        std::map<i64, bool> done;
        for (int v = 0; v < p; ++v) {
            if (sendOff[v] < 0 || done.count(sendOff[v])) continue;
            done[sendOff[v]] = true;
            const Msg& m = sendMsg[v];
            elb200_lattice d;
            d.src = Abuf; d.dst = packBuf + sendOff[v];
            d.nrows = m.nrows; d.ncols = m.ncols;
            d.s_off = m.s_off; d.s_rs = m.s_rs; d.s_cs = m.s_cs;
            d.d_off = 0; d.d_rs = 1; d.d_cs = m.nrows;
            packs.push_back(d);
        }
        LaunchLattices<T>(packs, false, nullptr, false);
        packs.clear();
    }
    // ---- wire ----
    for (int v = 0; v < p; ++v) {
        const int qi = v % r, qj = v / r;
        if (qi == mi && qj == mj) continue;
        const Msg &sm = sendMsg[v], &rm = recvMsg[v];
        if (sm.empty && rm.empty) continue;
        Wire wv;
        wv.peer = g.WorldRankOf(qi, qj);
        wv.v = v;
        wv.sendPtr = nullptr; wv.recvPtr = nullptr; wv.sendBytes = wv.recvBytes = 0;
        const bool sendP2P = ch >= 0 && !sm.empty && sizeof(T) * (size_t)sm.count() <= pp.regionBytes;
        const bool recvP2P = ch >= 0 && !rm.empty && sizeof(T) * (size_t)rm.count() <= pp.regionBytes;
        if (sendP2P) {
            // push this piece into the destination's window: region (channel, epoch parity, my rank)
            T* dstRemote = (T*)pp.Region(wv.peer, ch, ep, meW);
            if (Contig(sm.s_rs, sm.s_cs, sm.nrows, sm.ncols)) {
                pushes.push_back(Push{Abuf + sm.s_off, dstRemote, sizeof(T) * (size_t)sm.count()});
                st.zeroCopySends++;
            } else {
                elb200_lattice d;
                d.src = Abuf; d.dst = dstRemote;
                d.nrows = sm.nrows; d.ncols = sm.ncols;
                d.s_off = sm.s_off; d.s_rs = sm.s_rs; d.s_cs = sm.s_cs;
                d.d_off = 0; d.d_rs = 1; d.d_cs = sm.nrows;
                packs.push_back(d);
            }
            p2pDest.push_back(wv.peer);
            st.p2pPushes++;
            st.messages++;
            st.bytesSent += sizeof(T) * (size_t)sm.count();
        }
        if (recvP2P) {
            elb200_lattice d;
            d.src = (const T*)pp.Region(meW, ch, ep, wv.peer); d.dst = Bbuf;
            d.nrows = rm.nrows; d.ncols = rm.ncols;
            d.s_off = 0; d.s_rs = 1; d.s_cs = rm.nrows;
            d.d_off = rm.d_off; d.d_rs = rm.d_rs; d.d_cs = rm.d_cs;
            unpacks.push_back(d);
            p2pSrc.push_back(wv.peer);
        }
        if ((sm.empty || sendP2P) && (rm.empty || recvP2P)) continue;
        if (!sm.empty && !sendP2P) {
            wv.sendPtr = sendOff[v] >= 0 ? (const void*)(packBuf + sendOff[v]) : (const void*)(Abuf + sm.s_off);
            wv.sendBytes = sizeof(T) * (size_t)sm.count();
            if (sendOff[v] < 0) st.zeroCopySends++;
        }
        if (!rm.empty && !recvP2P) {
            wv.recvPtr = recvOff[v] >= 0 ? (void*)(recvBuf + recvOff[v]) : (void*)(Bbuf + rm.d_off);
            wv.recvBytes = sizeof(T) * (size_t)rm.count();
        }
        wires.push_back(wv);
    }
    // remote pack lattices were appended after the local ones were launched: launch them, then the
    // copy-engine pushes, then the flag exchange ("my pieces have landed" / wait for my sources)
    if (ch >= 0) {
        LaunchLattices<T>(packs, false, nullptr, false);
        for (const Push& pu : pushes) ELB_CUDA(cudaMemcpyAsync(pu.dst, pu.src, pu.bytes, cudaMemcpyDefault, s));
    }
    // AllGather pattern inside ONE communicator (the reference's RowAllGather.hpp:65-66, ColAllGather.hpp:72-73,
    // PartialColAllGather.hpp:80-82): this process sends the same piece to, and receives a piece of the same size
    // from, every other member of its grid row, grid column or the whole grid.  One ncclAllGather into staging in
    // communicator-rank order replaces the grouped send/recv; the unpack reads the staging directly.  (Only pieces
    // that do not ride the peer-memory path reach this point.)
    std::vector<const T*> recvSrc(p, nullptr);
    char* agStage = nullptr;
    static const bool agOn = [] { const char* e = std::getenv("ELB200_NCCL_ALLGATHER"); return !(e && std::atoi(e) == 0); }();
    if (agOn && !wires.empty() && wires[0].sendBytes > 0) {
        const void* sp = wires[0].sendPtr;
        const size_t bytes = wires[0].sendBytes;
        bool same = true;
        for (const Wire& wv : wires) same = same && wv.sendPtr == sp && wv.sendBytes == bytes && wv.recvBytes == bytes;
        const Comm* comms[3] = {&g.MRComm(), &g.MCComm(), &g.VCComm()};
        for (int ci = 0; same && ci < 3 && !agStage; ++ci) {
            const Comm& cm = *comms[ci];
            if (!cm.nccl || cm.size != (int)wires.size() + 1 || (int)cm.toWorld.size() != cm.size) continue;
            std::vector<int> rankOf(wires.size(), -1);
            bool match = cm.toWorld[cm.rank] == meW;
            for (size_t q = 0; match && q < wires.size(); ++q) {
                for (int t = 0; t < cm.size; ++t)
                    if (cm.toWorld[t] == wires[q].peer) rankOf[q] = t;
                match = rankOf[q] >= 0;
            }
            if (!match) continue;
            agStage = (char*)elb200::scratch_alloc(bytes * (size_t)cm.size, s);
            ELB_NCCL(ncclAllGather(sp, agStage, bytes, ncclInt8, (ncclComm_t)cm.nccl, s));
            st.allGathers++;
            st.messages += wires.size();
            st.bytesSent += bytes * wires.size();
            for (size_t q = 0; q < wires.size(); ++q) {
                const char* slot = agStage + bytes * (size_t)rankOf[q];
                if (recvOff[wires[q].v] >= 0) recvSrc[wires[q].v] = (const T*)slot;        // staged: unpack from here
                else ELB_CUDA(cudaMemcpyAsync(wires[q].recvPtr, slot, bytes, cudaMemcpyDeviceToDevice, s));   // in place
            }
            wires.clear();
        }
    }
    if (!wires.empty()) {
        if (!g.WorldNccl()) RuntimeError("Multi-rank redistribution without an NCCL communicator");
        ELB_NCCL(ncclGroupStart());
        for (const Wire& wv : wires) {
            if (wv.recvBytes) ELB_NCCL(ncclRecv(wv.recvPtr, wv.recvBytes, ncclInt8, wv.peer, (ncclComm_t)g.WorldNccl(), s));
            if (wv.sendBytes) {
                ELB_NCCL(ncclSend(wv.sendPtr, wv.sendBytes, ncclInt8, wv.peer, (ncclComm_t)g.WorldNccl(), s));
                st.messages++;
                st.bytesSent += wv.sendBytes;
            }
        }
        ELB_NCCL(ncclGroupEnd());
    }
    if (ch >= 0) {
        elb200::P2PFlagOps ops;
        ops.nsignal = ops.nwait = 0;
        ops.epoch = ep; ops.wait_value = ep; ops.error = pp.error;
        for (int w : p2pDest) ops.signal[ops.nsignal++] = pp.Ready(w, ch, meW);
        for (int w : p2pSrc) ops.wait[ops.nwait++] = pp.Ready(meW, ch, w);
        elb200::p2p_flags(ops, s);
    }
    // ---- unpack (+ the local part), applying conj / alpha / accumulate ----
    {
        const int me = mi + r * mj;
        const Msg& m = sendMsg[me];
        bool selfDone = false;
        if (!m.empty && plain && m.s_rs == 1 && m.d_rs == 1 && !transpose) {
            // plain local strided-column copy: let the copy engine do it
            ELB_CUDA(cudaMemcpy2DAsync(Bbuf + m.d_off, sizeof(T) * (size_t)m.d_cs, Abuf + m.s_off,
                                       sizeof(T) * (size_t)m.s_cs, sizeof(T) * (size_t)m.nrows, (size_t)m.ncols,
                                       cudaMemcpyDeviceToDevice, s));
            selfDone = true;
        }
        if (!m.empty && !selfDone) {
            elb200_lattice d;
            d.src = Abuf; d.dst = Bbuf;
            d.nrows = m.nrows; d.ncols = m.ncols;
            d.s_off = m.s_off; d.s_rs = m.s_rs; d.s_cs = m.s_cs;
            d.d_off = m.d_off; d.d_rs = m.d_rs; d.d_cs = m.d_cs;
            unpacks.push_back(d);
        }
        for (int v = 0; v < p; ++v) {
            if (recvOff[v] < 0) continue;
            const Msg& rm = recvMsg[v];
            elb200_lattice d;
            d.src = recvSrc[v] ? recvSrc[v] : recvBuf + recvOff[v]; d.dst = Bbuf;
            d.nrows = rm.nrows; d.ncols = rm.ncols;
            d.s_off = 0; d.s_rs = 1; d.s_cs = rm.nrows;
            d.d_off = rm.d_off; d.d_rs = rm.d_rs; d.d_cs = rm.d_cs;
            unpacks.push_back(d);
        }
        LaunchLattices<T>(unpacks, doConj, alpha == T(1) ? nullptr : &alpha, accumulate);
    }
    if (ch >= 0) {
        // "I have consumed epoch ep" to every peer; the window half of epoch ep + 1 is free once every
        // peer has consumed epoch ep - 1
        elb200::P2PFlagOps ops;
        ops.nsignal = ops.nwait = 0;
        ops.epoch = ep; ops.wait_value = ep - 1; ops.error = pp.error;
        for (int w = 0; w < p; ++w) {
            if (w == meW) continue;
            ops.signal[ops.nsignal++] = pp.Ack(w, ch, meW);
            if (ep >= 2) ops.wait[ops.nwait++] = pp.Ack(meW, ch, w);
        }
        elb200::p2p_flags(ops, s);
    }
    if (agStage) elb200::scratch_free(agStage, s);
    elb200::scratch_free(packBuf, s);
    elb200::scratch_free(recvBuf, s);
}

// Sum over the replicas of a partially replicated A and scatter: B += alpha op(sum A).
// One ncclReduceScatter inside the communicator A is replicated over (row comm for [MC,*] /
// [*,MC], column comm for [*,MR] / [MR,*], all ranks for [*,*]) produces the summed matrix
// in a non-replicated distribution T whose free alignment is chosen to coincide with B
// whenever possible, so the trailing accumulate into B is a purely local fused kernel.
template <typename T>
void SumScatter(T alpha, const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B, bool transposeOut, bool conj) {
    const Grid& g = A.Grid();
    if (&B.Grid() != &g) LogicError("AxpyContract requires both matrices on the same Grid");
    const Int h = A.Height(), w = A.Width();
    if ((transposeOut ? B.Width() : B.Height()) != h || (transposeOut ? B.Height() : B.Width()) != w)
        LogicError("Nonconformal AxpyContract");
    const Dist U = A.ColDist(), V = A.RowDist();
    const bool pinsI = PinsRow(U) || PinsRow(V), pinsJ = PinsCol(U) || PinsCol(V);
    if (pinsI && pinsJ) {  // nothing is replicated: plain axpy
        Redistribute(A, B, transposeOut, conj, alpha, true);
        return;
    }
    if (h == 0 || w == 0) return;
    // B's distribution as seen from T's index space
    const Layout bView = transposeOut ? Layout{B.RowDist(), B.ColDist(), B.RowAlign(), B.ColAlign()}
                                      : Layout{B.ColDist(), B.RowDist(), B.ColAlign(), B.RowAlign()};
    plan::ContractPlan P = plan::BuildContractPlan(g, h, w, LayoutOf(A), A.LDim(), bView);
    const Comm* comm = P.kind == plan::OVER_MR ? &g.MRComm() : (P.kind == plan::OVER_MC ? &g.MCComm() : &g.VCComm());
    const Layout lt = P.T;
    const int np = comm->size;
    const i64 chunk = P.chunk, myRows = P.myRows;
    const bool needZero = P.needZero;
    cudaStream_t s = dev::stream();
    T* sendBuf = (T*)elb200::scratch_alloc(sizeof(T) * (size_t)(chunk * np), s);
    T* recvBuf = np > 1 ? (T*)elb200::scratch_alloc(sizeof(T) * (size_t)chunk, s) : sendBuf;
    std::vector<elb200_lattice> packs;
    for (int q = 0; q < np; ++q) {
        const Msg& m = P.packs[q];
        if (m.empty) continue;
        elb200_lattice d;
        d.src = A.LockedBuffer(); d.dst = sendBuf + (i64)q * chunk;
        d.nrows = m.nrows; d.ncols = m.ncols;
        d.s_off = m.s_off; d.s_rs = m.s_rs; d.s_cs = m.s_cs;
        d.d_off = m.d_off; d.d_rs = m.d_rs; d.d_cs = m.d_cs;
        packs.push_back(d);
    }
    if (needZero) ELB_CUDA(cudaMemsetAsync(sendBuf, 0, sizeof(T) * (size_t)(chunk * np), s));
    LaunchLattices<T>(packs, false, nullptr, false);
    if (np > 1) {
        const size_t reals = (size_t)chunk * (IsComplex<T>::value ? 2 : 1);
        ELB_NCCL(ncclReduceScatter(sendBuf, recvBuf, reals, NcclReal(sizeof(Base<T>)), ncclSum, (ncclComm_t)comm->nccl, s));
        GetRedistStats().reduceScatters++;
        GetRedistStats().bytesSent += sizeof(T) * (size_t)chunk * (size_t)(np - 1);
    }
    // recvBuf holds my local piece of T; accumulate it into B (local when T coincides with B)
    AbstractDistMatrix<T> Tm(g, lt.U, lt.V);
    Tm.Attach(h, w, g, lt.colAlign, lt.rowAlign, recvBuf, (Int)std::max<i64>(myRows, 1));
    Redistribute(static_cast<const AbstractDistMatrix<T>&>(Tm), B, transposeOut, conj, alpha, true);
    if (np > 1) elb200::scratch_free(recvBuf, s);
    elb200::scratch_free(sendBuf, s);
}

}  // namespace

// ---------------------------------------------------------------------------
// public entry points
// ---------------------------------------------------------------------------
template <typename T>
void Copy(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B) { Redistribute(A, B, false, false, T(1), false); }
template <typename T>
void Transpose(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B, bool conjugate) {
    Redistribute(A, B, true, conjugate, T(1), false);
}
template <typename T>
void Adjoint(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B) { Redistribute(A, B, true, true, T(1), false); }
template <typename T>
void Axpy(T alpha, const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B) { Redistribute(A, B, false, false, alpha, true); }
template <typename T>
void AxpyContract(T alpha, const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B) { SumScatter(alpha, A, B, false, false); }
template <typename T>
void TransposeAxpyContract(T alpha, const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B, bool conjugate) {
    SumScatter(alpha, A, B, true, conjugate);
}
template <typename T>
void Contract(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B) {
    AdoptAlignments(A, B, false);
    B.Resize(A.Height(), A.Width());
    Zero(B);
    SumScatter(T(1), A, B, false, false);
}

template <typename T>
void Scale(T alpha, Matrix<T>& A) {
    if (alpha == T(1) || A.Height() == 0 || A.Width() == 0) return;
    if (alpha == T(0)) { Zero(A); return; }
    dev::D<T> a = dev::val<T>(alpha);
    elb200::lattice_copy_device<dev::D<T>>(dev::ptr(A.LockedBuffer()), dev::ptr(A.Buffer()), A.Height(), A.Width(), 0, 1,
                                           A.LDim(), 0, 1, A.LDim(), false, &a, false, dev::stream());
}
template <typename T>
void Scale(T alpha, AbstractDistMatrix<T>& A) { Scale(alpha, A.Matrix()); }
template <typename T>
void Zero(Matrix<T>& A) {
    if (A.Height() == 0 || A.Width() == 0) return;
    ELB_CUDA(cudaMemset2DAsync(A.Buffer(), sizeof(T) * (size_t)A.LDim(), 0, sizeof(T) * (size_t)A.Height(),
                               (size_t)A.Width(), dev::stream()));
}
template <typename T>
void Zero(AbstractDistMatrix<T>& A) { Zero(A.Matrix()); }
template <typename T>
void Zeros(AbstractDistMatrix<T>& A, Int m, Int n) { A.Resize(m, n); Zero(A); }
template <typename T>
void Conjugate(AbstractDistMatrix<T>& A) {
    if (!IsComplex<T>::value || A.LocalHeight() == 0 || A.LocalWidth() == 0) return;
    elb200::lattice_copy_device<dev::D<T>>(dev::ptr(A.LockedBuffer()), dev::ptr(A.Buffer()), A.LocalHeight(),
                                           A.LocalWidth(), 0, 1, A.LDim(), 0, 1, A.LDim(), true, nullptr, false,
                                           dev::stream());
}
template <typename T>
void ScaleTrapezoid(T alpha, UpperOrLower uplo, AbstractDistMatrix<T>& A, Int offset) {
    if (alpha == T(1)) return;
    dev::D<T> a = dev::val<T>(alpha);
    dev::c_check(elb200_scale_trapezoid(dev::Code<T>(), &a, UpperOrLowerToChar(uplo), A.LocalHeight(), A.LocalWidth(),
                                        A.Buffer(), A.LDim(), A.ColShift(), A.ColStride(), A.RowShift(), A.RowStride(),
                                        offset, (elb200_stream_t)dev::stream()),
                 "elb200_scale_trapezoid");
}
// Y_trap += alpha X_trap.  Equal distributions and alignments: LocalAxpyTrapezoid (AxpyTrapezoid.hpp:73-125), one
// kernel on the local staircases; otherwise X is first brought into Y's distribution (:128-160).
template <typename T>
void LocalAxpyTrapezoid(UpperOrLower uplo, T alpha, const AbstractDistMatrix<T>& X, AbstractDistMatrix<T>& Y, Int offset) {
    if (X.Height() != Y.Height() || X.Width() != Y.Width()) LogicError("Nonconformal AxpyTrapezoid");
    if (X.ColDist() != Y.ColDist() || X.RowDist() != Y.RowDist() || X.ColAlign() != Y.ColAlign() ||
        X.RowAlign() != Y.RowAlign())
        LogicError("LocalAxpyTrapezoid needs identically distributed and aligned operands");
    dev::D<T> a = dev::val<T>(alpha);
    dev::c_check(elb200_axpy_trapezoid(dev::Code<T>(), &a, UpperOrLowerToChar(uplo), Y.LocalHeight(), Y.LocalWidth(),
                                       X.LockedBuffer(), X.LDim(), Y.Buffer(), Y.LDim(), Y.ColShift(), Y.ColStride(),
                                       Y.RowShift(), Y.RowStride(), offset, (elb200_stream_t)dev::stream()),
                 "elb200_axpy_trapezoid");
}
template <typename T>
void AxpyTrapezoid(UpperOrLower uplo, T alpha, const AbstractDistMatrix<T>& X, AbstractDistMatrix<T>& Y, Int offset) {
    if (X.ColDist() == Y.ColDist() && X.RowDist() == Y.RowDist() && X.ColAlign() == Y.ColAlign() &&
        X.RowAlign() == Y.RowAlign()) {
        LocalAxpyTrapezoid(uplo, alpha, X, Y, offset);
        return;
    }
    AbstractDistMatrix<T> XCopy(Y.Grid(), Y.ColDist(), Y.RowDist());
    XCopy.AlignWith(Y);
    Copy(X, XCopy);
    LocalAxpyTrapezoid(uplo, alpha, static_cast<const AbstractDistMatrix<T>&>(XCopy), Y, offset);
}
template <typename T>
void MakeTrapezoidal(UpperOrLower uplo, AbstractDistMatrix<T>& A, Int offset) {
    dev::c_check(elb200_make_trapezoidal(dev::Code<T>(), UpperOrLowerToChar(uplo), A.LocalHeight(), A.LocalWidth(),
                                         A.Buffer(), A.LDim(), A.ColShift(), A.ColStride(), A.RowShift(), A.RowStride(),
                                         offset, (elb200_stream_t)dev::stream()),
                 "elb200_make_trapezoidal");
}
template <typename T>
void HashFill(AbstractDistMatrix<T>& A, int kind, uint64_t seed, double diag) {
    dev::c_check(elb200_fill_hash(dev::Code<T>(), kind, A.LocalHeight(), A.LocalWidth(), A.Buffer(), A.LDim(),
                                  A.ColShift(), A.ColStride(), A.RowShift(), A.RowStride(), seed, diag,
                                  (elb200_stream_t)dev::stream()),
                 "elb200_fill_hash");
}

namespace {
// the communicator over which the ranks hold DISTINCT pieces of A (null: every rank holds all of it)
template <typename T>
const Comm* OwnersComm(const AbstractDistMatrix<T>& A) {
    const Grid& g = A.Grid();
    const bool pinsI = PinsRow(A.ColDist()) || PinsRow(A.RowDist());
    const bool pinsJ = PinsCol(A.ColDist()) || PinsCol(A.RowDist());
    const Comm* comm = nullptr;
    if (pinsI && pinsJ) comm = &g.VCComm();
    else if (pinsI) comm = &g.MCComm();
    else if (pinsJ) comm = &g.MRComm();
    return (comm && comm->size > 1) ? comm : nullptr;
}
double ReadDeviceDouble(const double* d) {
    double out = 0.0;
    ELB_CUDA(cudaMemcpyAsync(&out, d, sizeof(double), cudaMemcpyDeviceToHost, dev::stream()));
    ELB_CUDA(cudaStreamSynchronize(dev::stream()));
    return out;
}
// max |a_ij| over the whole matrix into d[0].  The reduction across ranks runs on the BIT PATTERN
// (non-negative doubles order like unsigned integers, +inf and the canonical NaN on top), so a NaN
// on any rank reaches every rank whatever NCCL's floating-point max does with NaNs.
template <typename T>
void DeviceMaxAbs(const AbstractDistMatrix<T>& A, double* d) {
    ELB_CUDA(cudaMemsetAsync(d, 0, sizeof(double), dev::stream()));
    dev::c_check(elb200_maxabs(dev::Code<T>(), A.LocalHeight(), A.LocalWidth(), A.LockedBuffer(), A.LDim(), d,
                               (elb200_stream_t)dev::stream()),
                 "elb200_maxabs");
    if (const Comm* comm = OwnersComm(A))
        ELB_NCCL(ncclAllReduce(d, d, 1, ncclUint64, ncclMax, (ncclComm_t)comm->nccl, dev::stream()));
}
}  // namespace

// Two passes, both on the device: the max-abs, then the sum of squares of the entries divided by it
// (src/lapack_like/props/Norm/Frobenius.cpp keeps the same (scale, scaledSquare) pair in one host
// pass): no overflow for |a| > 1e154, no underflow for tiny entries, NaN / inf propagate.
template <typename T>
Base<T> FrobeniusNorm(const AbstractDistMatrix<T>& A) {
    double* d = (double*)elb200::scratch_alloc(2 * sizeof(double), dev::stream());
    ELB_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(double), dev::stream()));
    DeviceMaxAbs(A, d);
    dev::c_check(elb200_sumsq_scaled(dev::Code<T>(), A.LocalHeight(), A.LocalWidth(), A.LockedBuffer(), A.LDim(), d,
                                     d + 1, (elb200_stream_t)dev::stream()),
                 "elb200_sumsq_scaled");
    if (const Comm* comm = OwnersComm(A))
        ELB_NCCL(ncclAllReduce(d + 1, d + 1, 1, ncclDouble, ncclSum, (ncclComm_t)comm->nccl, dev::stream()));
    double h[2] = {0.0, 0.0};
    ELB_CUDA(cudaMemcpyAsync(h, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, dev::stream()));
    ELB_CUDA(cudaStreamSynchronize(dev::stream()));
    elb200::scratch_free(d, dev::stream());
    const double scale = h[0];
    if (!(scale > 0.0) || !(scale < HUGE_VAL)) return (Base<T>)scale;  // 0, inf or NaN
    return (Base<T>)(scale * std::sqrt(h[1]));
}
template <typename T>
Base<T> MaxNorm(const AbstractDistMatrix<T>& A) {
    double* d = (double*)elb200::scratch_alloc(sizeof(double), dev::stream());
    DeviceMaxAbs(A, d);
    const double s = ReadDeviceDouble(d);
    elb200::scratch_free(d, dev::stream());
    return (Base<T>)s;
}

#define ELB_INST(T)                                                                                   \
    template void Copy(const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&);                         \
    template void Transpose(const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&, bool);              \
    template void Adjoint(const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&);                      \
    template void Axpy(T, const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&);                      \
    template void AxpyContract(T, const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&);              \
    template void TransposeAxpyContract(T, const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&, bool); \
    template void Contract(const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&);                     \
    template void Scale(T, Matrix<T>&);                                                               \
    template void Scale(T, AbstractDistMatrix<T>&);                                                   \
    template void Zero(Matrix<T>&);                                                                   \
    template void Zero(AbstractDistMatrix<T>&);                                                       \
    template void Zeros(AbstractDistMatrix<T>&, Int, Int);                                            \
    template void Conjugate(AbstractDistMatrix<T>&);                                                  \
    template void ScaleTrapezoid(T, UpperOrLower, AbstractDistMatrix<T>&, Int);                       \
    template void MakeTrapezoidal(UpperOrLower, AbstractDistMatrix<T>&, Int);                         \
    template void AxpyTrapezoid(UpperOrLower, T, const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&, Int);      \
    template void LocalAxpyTrapezoid(UpperOrLower, T, const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&, Int); \
    template void HashFill(AbstractDistMatrix<T>&, int, uint64_t, double);                            \
    template Base<T> FrobeniusNorm(const AbstractDistMatrix<T>&);                                     \
    template Base<T> MaxNorm(const AbstractDistMatrix<T>&);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
