// Distributed level-3: Gemm (SUMMA stationary-A / -B / -C / Dot for all orientations),
// Trrk, Herk/Syrk and Trsm (all eight side/uplo/orientation families) on
// DistMatrix<T,MC,MR> over a Grid.
//
// Each routine keeps the reference's panel loop at Blocksize() granularity and its
// algorithm selection, but the panels are formed by the generic redistribution engine
// (redist.cpp) directly in the layout the GEMM kernel wants, which removes the
// reference's intermediate hops ([VR,*] demotes, explicit transposes, the 3-collective
// [MR,MC]->[MC,MR] epilogue): every variant reduces to "redistribute panel(s), one local
// tensor-pipe GEMM (always N,N or a masked GEMM), optionally one reduce-scatter".
//   reference: src/blas_like/level3/Gemm.cpp:90-133, Gemm/{NN,NT,TN,TT}.hpp,
//              Trrk.cpp:100-116 + Trrk/*.hpp, Syrk.cpp:70-86 + Syrk/*.hpp, Herk.cpp,
//              Trsm.cpp:67-375 + Trsm/{LLN,LLT,LUN,LUT,RLN,RLT,RUN,RUT}.hpp
#include <algorithm>
#include <cstdlib>
#include <memory>

#include "dev.hpp"
#include "elb200/level3.hpp"

namespace El {

namespace {

template <typename T>
AbstractDistMatrix<T> LockedView(const AbstractDistMatrix<T>& A, Int i, Int j, Int h, Int w) {
    AbstractDistMatrix<T> V(A.Grid(), A.ColDist(), A.RowDist());
    V.LockedViewOf(A, i, j, h, w);
    return V;
}
template <typename T>
AbstractDistMatrix<T> View(AbstractDistMatrix<T>& A, Int i, Int j, Int h, Int w) {
    AbstractDistMatrix<T> V(A.Grid(), A.ColDist(), A.RowDist());
    V.ViewOf(A, i, j, h, w);
    return V;
}

// Read proxy: the matrix itself when already [MC,MR], else a redistributed copy
// (DistMatrixReadProxy<T,T,MC,MR>, include/El/core/Proxy.hpp:234-253)
template <typename T>
struct ReadProxy {
    const AbstractDistMatrix<T>* ptr;
    std::unique_ptr<AbstractDistMatrix<T>> copy;
    explicit ReadProxy(const AbstractDistMatrix<T>& A, Dist U = MC, Dist V = MR) {
        if (A.ColDist() == U && A.RowDist() == V) ptr = &A;
        else {
            copy.reset(new AbstractDistMatrix<T>(A.Grid(), U, V));
            Copy(A, *copy);
            ptr = copy.get();
        }
    }
    const AbstractDistMatrix<T>& Get() const { return *ptr; }
};
template <typename T>
struct ReadWriteProxy {
    AbstractDistMatrix<T>* orig;
    AbstractDistMatrix<T>* ptr;
    std::unique_ptr<AbstractDistMatrix<T>> copy;
    explicit ReadWriteProxy(AbstractDistMatrix<T>& A) : orig(&A) {
        if (A.ColDist() == MC && A.RowDist() == MR) ptr = &A;
        else {
            copy.reset(new AbstractDistMatrix<T>(A.Grid(), MC, MR));
            Copy(A, *copy);
            ptr = copy.get();
        }
    }
    AbstractDistMatrix<T>& Get() { return *ptr; }
    void Commit() { if (copy) Copy(*copy, *orig); }
};

template <typename T>
void LocalGemmRaw(char ta, char tb, T alpha, const Matrix<T>& A, const Matrix<T>& B, T beta, Matrix<T>& C) {
    const Int m = C.Height(), n = C.Width();
    const Int k = (ta == 'N') ? A.Width() : A.Height();
    const Int am = (ta == 'N') ? A.Height() : A.Width();
    const Int bk = (tb == 'N') ? B.Height() : B.Width();
    const Int bn = (tb == 'N') ? B.Width() : B.Height();
    if (am != m || bn != n || bk != k) LogicError("Nonconformal local Gemm");
    if (m == 0 || n == 0) return;
    if (k == 0) { Scale(beta, C); return; }  // Gemm.cpp:68-71
    elb200::gemm_device<dev::D<T>>(0, ta, tb, m, n, k, dev::val<T>(alpha), dev::ptr(A.LockedBuffer()), A.LDim(),
                                   dev::ptr(B.LockedBuffer()), B.LDim(), dev::val<T>(beta), dev::ptr(C.Buffer()),
                                   C.LDim(), 0, 1, 0, 1, dev::stream());
}

template <typename T>
void AssertSameGrid(const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B) {
    if (&A.Grid() != &B.Grid()) LogicError("Grids did not match");
}

}  // namespace

GemmAlgorithm GemmDefaultAlgorithm(Int m, Int n, Int k) {
    const double weightTowardsC = 2., weightAwayFromDot = 10.;
    if (weightAwayFromDot * m <= k && weightAwayFromDot * n <= k) return GEMM_SUMMA_DOT;
    if (m <= n && weightTowardsC * m <= k) return GEMM_SUMMA_B;
    if (n <= m && weightTowardsC * n <= k) return GEMM_SUMMA_A;
    return GEMM_SUMMA_C;
}

// ---------------------------------------------------------------------------
// Gemm
// ---------------------------------------------------------------------------
template <typename T>
void Gemm(Orientation oA, Orientation oB, T alpha, const Matrix<T>& A, const Matrix<T>& B, T beta, Matrix<T>& C) {
    LocalGemmRaw(OrientationToChar(oA), OrientationToChar(oB), alpha, A, B, beta, C);
}
template <typename T>
void Gemm(Orientation oA, Orientation oB, T alpha, const Matrix<T>& A, const Matrix<T>& B, Matrix<T>& C) {
    const Int m = (oA == NORMAL) ? A.Height() : A.Width();
    const Int n = (oB == NORMAL) ? B.Width() : B.Height();
    C.Resize(m, n);
    Zero(C);
    Gemm(oA, oB, alpha, A, B, T(0), C);
}
template <typename T>
void LocalGemm(Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& A,
               const AbstractDistMatrix<T>& B, T beta, AbstractDistMatrix<T>& C) {
    LocalGemmRaw(OrientationToChar(oA), OrientationToChar(oB), alpha, A.LockedMatrix(), B.LockedMatrix(), beta,
                 C.Matrix());
}
template <typename T>
void LocalGemm(Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& A,
               const AbstractDistMatrix<T>& B, AbstractDistMatrix<T>& C) {
    // Gemm.cpp:247-259: Resize, Zero, multiply -- the caller has aligned C (the beta form checks conformity)
    const Int m = (oA == NORMAL) ? A.Height() : A.Width();
    const Int n = (oB == NORMAL) ? B.Width() : B.Height();
    C.Resize(m, n);
    Zero(C);
    LocalGemm(oA, oB, alpha, A, B, T(0), C);
}

namespace {

// op(X) restricted to rows [i0,i0+h) x cols [j0,j0+w) of op(X), redistributed into P
// (P's distribution / alignment already set): a Copy of the corresponding block of X
// when o == NORMAL, a (conjugate-)transposing redistribution otherwise.
template <typename T>
void FormPanel(Orientation o, const AbstractDistMatrix<T>& X, Int i0, Int j0, Int h, Int w, AbstractDistMatrix<T>& P) {
    if (o == NORMAL) {
        auto V = LockedView(X, i0, j0, h, w);
        Copy(static_cast<const AbstractDistMatrix<T>&>(V), P);
    } else {
        auto V = LockedView(X, j0, i0, w, h);
        Transpose(static_cast<const AbstractDistMatrix<T>&>(V), P, o == ADJOINT);
    }
}

// Stationary C: rank-nb updates (Gemm/NN.hpp:179-218 and the NT/TN/TT mirrors).
// On a multi-process grid the panels of step k+1 are gathered on the panel stream (NCCL over
// NVLink) while the tensor-pipe update of step k runs on the main stream: double-buffered
// panels, two events per slot.  The persistent GEMM leaves PanelSms() SMs to the gather.
template <typename T>
void SummaC(Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
            AbstractDistMatrix<T>& C) {
    const Grid& g = C.Grid();
    const Int m = C.Height(), n = C.Width();
    const Int sumDim = (oA == NORMAL) ? A.Width() : A.Height();
    const Int bsize = Blocksize();
    const Int steps = (sumDim + bsize - 1) / bsize;
    if (g.Size() == 1 || steps < 2 || !dev::OverlapEnabled()) {
        AbstractDistMatrix<T> A1(g, MC, STAR), B1(g, STAR, MR);
        A1.AlignWith(C);
        B1.AlignWith(C);
        for (Int k = 0; k < sumDim; k += bsize) {
            const Int nb = std::min(bsize, sumDim - k);
            FormPanel(oA, A, 0, k, m, nb, A1);  // op(A)(:, k:k+nb) -> [MC,*]
            FormPanel(oB, B, k, 0, nb, n, B1);  // op(B)(k:k+nb, :) -> [*,MR]
            LocalGemm(NORMAL, NORMAL, alpha, A1, B1, T(1), C);
        }
        return;
    }
    cudaStream_t mainS = dev::stream(), panelS = elb200::aux_stream(0);
    // SMs left to the gather kernels of the panel stream: 8 for NCCL's send / recv kernels, none with the
    // peer-memory path (the wire step runs on the copy engines; the short unpack kernels of the high-priority panel
    // stream slip in between two updates).  Measured at N = 2: 63.9 -> 67.3 TFLOP/s.  ELB200_SUMMA_PANEL_SMS overrides.
    int reserve = g.P2P().on ? 0 : dev::PanelSms(8);
    if (const char* e = std::getenv("ELB200_SUMMA_PANEL_SMS")) reserve = std::atoi(e);
    const int gemmSms = reserve > 0 ? std::max(1, elb200::sm_count() - reserve) : 0;
    // Panels are gathered SUMMA_BATCH (default 4) k-steps at a time -- one redistribution of a (batch x Blocksize())
    // wide panel instead of `batch` narrow ones -- and consumed by `batch` local rank-Blocksize() updates on views
    // of it: every entry of C still receives the same updates in the same order (bit-identical), but the latency
    // chain of a redistribution (pushes, flags, unpack) is paid once per batch.  That chain is what bounds SUMMA-C
    // once the local update of one step is short (8 GPUs; the column bands of GemmHost).
    Int batch = 4;
    if (const char* e = std::getenv("ELB200_SUMMA_BATCH")) batch = std::max(1, std::atoi(e));
    const Int wide = batch * bsize;
    AbstractDistMatrix<T> A1[2] = {AbstractDistMatrix<T>(g, MC, STAR), AbstractDistMatrix<T>(g, MC, STAR)};
    AbstractDistMatrix<T> B1[2] = {AbstractDistMatrix<T>(g, STAR, MR), AbstractDistMatrix<T>(g, STAR, MR)};
    dev::Event ready[2], freed[2], fork;
    for (int s = 0; s < 2; ++s) { A1[s].AlignWith(C); B1[s].AlignWith(C); }
    fork.Record(mainS);   // A, B (and the scaled C) are final on the main stream from here on
    fork.Wait(panelS);
    Int it = 0;
    for (Int k = 0; k < sumDim; k += wide, ++it) {
        const Int kw = std::min(wide, sumDim - k);
        const int slot = (int)(it & 1);
        {
            dev::StreamScope onPanel(panelS);
            if (it >= 2) freed[slot].Wait(panelS);
            FormPanel(oA, A, 0, k, m, kw, A1[slot]);
            FormPanel(oB, B, k, 0, kw, n, B1[slot]);
            ready[slot].Record(panelS);
        }
        ready[slot].Wait(mainS);
        {
            dev::SmLimitScope lim(gemmSms);
            for (Int j = 0; j < kw; j += bsize) {
                const Int nb = std::min(bsize, kw - j);
                auto Av = LockedView(static_cast<const AbstractDistMatrix<T>&>(A1[slot]), 0, j, m, nb);
                auto Bv = LockedView(static_cast<const AbstractDistMatrix<T>&>(B1[slot]), j, 0, nb, n);
                LocalGemm(NORMAL, NORMAL, alpha, Av, Bv, T(1), C);
            }
        }
        freed[slot].Record(mainS);
    }
    // the panels are released in main-stream order (their last reader)
}

// Stationary A: panels of op(B), skinny local product, sum-scatter into C
// (Gemm/NN.hpp:93-134, NT.hpp:15-55, TN.hpp:15-57, TT.hpp:15-59)
template <typename T>
void SummaA(Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
            AbstractDistMatrix<T>& C) {
    const Grid& g = C.Grid();
    const Int n = C.Width(), m = C.Height();
    const Int sumDim = (oA == NORMAL) ? A.Width() : A.Height();
    const Int bsize = Blocksize();
    // op(B) panel must be distributed like the summation index of A's local matrix
    AbstractDistMatrix<T> B1(g, oA == NORMAL ? MR : MC, STAR);
    AbstractDistMatrix<T> D1(g, oA == NORMAL ? MC : MR, STAR);
    if (oA == NORMAL) { B1.AlignCols(A.RowAlign()); D1.AlignCols(A.ColAlign()); }
    else { B1.AlignCols(A.ColAlign()); D1.AlignCols(A.RowAlign()); }
    // The columns of D1 = op(A) op(B)(:, panel) are independent products, so widening the panel changes no entry of
    // the result; four Blocksize() panels per step give the local GEMM 4x the columns (a 128-column product fills
    // 1.7 waves of the persistent kernel; measured 15 -> see profiles) and quarter the number of sum-scatters.
    const Int step = 4 * bsize;
    for (Int k = 0; k < n; k += step) {
        const Int nb = std::min(step, n - k);
        FormPanel(oB, B, 0, k, sumDim, nb, B1);  // op(B)(:, k:k+nb)
        D1.Resize(m, nb);
        LocalGemm(oA, NORMAL, alpha, A, B1, T(0), D1);
        auto C1 = View(C, 0, k, m, nb);
        AxpyContract(T(1), static_cast<const AbstractDistMatrix<T>&>(D1), C1);
    }
}

// Stationary B (Gemm/NN.hpp:138-176, NT.hpp:59-101, TN.hpp:61-100, TT.hpp:63-110)
template <typename T>
void SummaB(Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
            AbstractDistMatrix<T>& C) {
    const Grid& g = C.Grid();
    const Int n = C.Width(), m = C.Height();
    const Int sumDim = (oA == NORMAL) ? A.Width() : A.Height();
    const Int bsize = Blocksize();
    AbstractDistMatrix<T> A1(g, STAR, oB == NORMAL ? MC : MR);
    AbstractDistMatrix<T> D1(g, STAR, oB == NORMAL ? MR : MC);
    if (oB == NORMAL) { A1.AlignRows(B.ColAlign()); D1.AlignRows(B.RowAlign()); }
    else { A1.AlignRows(B.RowAlign()); D1.AlignRows(B.ColAlign()); }
    const Int step = 4 * bsize;   // rows of D1 are independent products: see SummaA
    for (Int k = 0; k < m; k += step) {
        const Int nb = std::min(step, m - k);
        FormPanel(oA, A, k, 0, nb, sumDim, A1);  // op(A)(k:k+nb, :)
        D1.Resize(nb, n);
        LocalGemm(NORMAL, oB, alpha, A1, B, T(0), D1);
        auto C1 = View(C, k, 0, nb, n);
        AxpyContract(T(1), static_cast<const AbstractDistMatrix<T>&>(D1), C1);
    }
}

// Dot: 1-D distribute the summation index over all p ranks once, then one local product
// and one world reduce-scatter per blockSize x blockSize block of C (Gemm/NN.hpp:226-270)
template <typename T>
void SummaDot(Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
              AbstractDistMatrix<T>& C) {
    const Int blockSize = GemmDotBlocksize(sizeof(T));
    const Grid& g = C.Grid();
    const Int m = C.Height(), n = C.Width();
    const Int sumDim = (oA == NORMAL) ? A.Width() : A.Height();
    AbstractDistMatrix<T> AV(g, STAR, VC), BV(g, VC, STAR);
    FormPanel(oA, A, 0, 0, m, sumDim, AV);  // op(A)[*,VC]
    BV.AlignCols(AV.RowAlign());
    FormPanel(oB, B, 0, 0, sumDim, n, BV);  // op(B)[VC,*]
    AbstractDistMatrix<T> C11(g, STAR, STAR);
    for (Int ko = 0; ko < m; ko += blockSize) {
        const Int nbo = std::min(blockSize, m - ko);
        auto A1 = LockedView(static_cast<const AbstractDistMatrix<T>&>(AV), ko, 0, nbo, sumDim);
        for (Int ki = 0; ki < n; ki += blockSize) {
            const Int nbi = std::min(blockSize, n - ki);
            auto B1 = LockedView(static_cast<const AbstractDistMatrix<T>&>(BV), 0, ki, sumDim, nbi);
            C11.Resize(nbo, nbi);
            LocalGemm(NORMAL, NORMAL, alpha, A1, B1, T(0), C11);
            auto Cb = View(C, ko, ki, nbo, nbi);
            AxpyContract(T(1), static_cast<const AbstractDistMatrix<T>&>(C11), Cb);
        }
    }
}

// Cannon's algorithm on a square grid (Gemm/NN.hpp:15-89): every process keeps one package of A and one of B;
// after an initial skew the package of A moves one process to the left and the package of B one process up per
// step, and the local product of what a process holds is accumulated into its block of C.  sqrt(p) steps, each
// moving the WHOLE local matrices once -- the least data of any variant.  The ring shifts are ncclSend / ncclRecv
// pairs inside the row and the column communicator; packages are double-buffered, so the shift for step q + 1 runs
// on the panel stream underneath the product of step q.
template <typename T>
void CannonNN(T alpha, const AbstractDistMatrix<T>& APre, const AbstractDistMatrix<T>& BPre, AbstractDistMatrix<T>& C) {
    const Grid& g = C.Grid();
    if (g.Height() != g.Width()) LogicError("Process grid must be square for Cannon's");
    const int ps = g.Height();
    if (APre.Width() % ps != 0) LogicError("For now, width(A) must be integer multiple of sqrt(p)");
    // A and B in [MC,MR], A's rows aligned with C's, B's columns aligned with C's (NN.hpp:29-36)
    AbstractDistMatrix<T> Ac(g, MC, MR), Bc(g, MC, MR);
    const AbstractDistMatrix<T>* A = &APre;
    const AbstractDistMatrix<T>* B = &BPre;
    if (APre.ColAlign() != C.ColAlign()) { Ac.AlignCols(C.ColAlign()); Copy(APre, Ac); A = &Ac; }
    if (BPre.RowAlign() != C.RowAlign()) { Bc.AlignRows(C.RowAlign()); Copy(BPre, Bc); B = &Bc; }
    const int row = g.Row(), col = g.Col();
    const Int lhA = A->LocalHeight(), lwA = A->LocalWidth(), lhB = B->LocalHeight(), lwB = B->LocalWidth();
    Matrix<T> pkgA[2] = {Matrix<T>(lhA, lwA), Matrix<T>(lhA, lwA)};
    Matrix<T> pkgB[2] = {Matrix<T>(lhB, lwB), Matrix<T>(lhB, lwB)};
    Copy(A->LockedMatrix(), pkgA[0]);
    Copy(B->LockedMatrix(), pkgB[0]);
    cudaStream_t mainS = dev::stream();
    auto shift = [&](Matrix<T>& from, Matrix<T>& to, const Comm& comm, int sendTo, int recvFrom, cudaStream_t s) {
        // packages of one ring all have the same local shape (width(A) is a multiple of sqrt(p); the other
        // dimension is shared by the whole grid row / column)
        const size_t bytes = sizeof(T) * size_t(from.LDim()) * size_t(from.Width());
        if (comm.size == 1 || (sendTo == comm.rank && recvFrom == comm.rank)) {
            ELB_CUDA(cudaMemcpyAsync(to.Buffer(), from.LockedBuffer(), bytes, cudaMemcpyDeviceToDevice, s));
            return;
        }
        ELB_NCCL(ncclGroupStart());
        ELB_NCCL(ncclSend(from.LockedBuffer(), bytes, ncclInt8, sendTo, (ncclComm_t)comm.nccl, s));
        ELB_NCCL(ncclRecv(to.Buffer(), bytes, ncclInt8, recvFrom, (ncclComm_t)comm.nccl, s));
        ELB_NCCL(ncclGroupEnd());
    };
    auto mod = [&](int x) { return ((x % ps) + ps) % ps; };
    // initial skew (NN.hpp:55-64): afterwards the column residue of A's package equals the row residue of B's
    const int rowShiftA = A->RowShift(), colShiftB = B->ColShift();
    int cur = 0;
    if (ps > 1) {
        shift(pkgA[0], pkgA[1], g.MRComm(), mod(col - colShiftB), mod(col + colShiftB), mainS);
        shift(pkgB[0], pkgB[1], g.MCComm(), mod(row - rowShiftA), mod(row + rowShiftA), mainS);
        cur = 1;
    }
    const bool overlap = ps > 1 && dev::OverlapEnabled();
    cudaStream_t panelS = overlap ? elb200::aux_stream(0) : mainS;
    dev::Event ready, consumed[2], fork;
    for (int q = 0; q < ps; ++q) {
        const int nxt = cur ^ 1;
        if (q != ps - 1) {
            // the packages of step q are final on the main stream; the next ones may be overwritten once the product
            // that last read them (step q - 1) is done
            fork.Record(mainS);
            fork.Wait(panelS);
            if (q >= 1) consumed[nxt].Wait(panelS);
            shift(pkgA[cur], pkgA[nxt], g.MRComm(), mod(col - 1), mod(col + 1), panelS);
            shift(pkgB[cur], pkgB[nxt], g.MCComm(), mod(row - 1), mod(row + 1), panelS);
            ready.Record(panelS);
        }
        {
            dev::SmLimitScope lim(q != ps - 1 && overlap ? std::max(1, elb200::sm_count() - dev::PanelSms(8)) : 0);
            LocalGemmRaw('N', 'N', alpha, pkgA[cur], pkgB[cur], T(1), C.Matrix());
        }
        consumed[cur].Record(mainS);
        if (q != ps - 1) ready.Wait(mainS);
        cur = nxt;
    }
}

}  // namespace

template <typename T>
void Gemm(Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& APre, const AbstractDistMatrix<T>& BPre,
          T beta, AbstractDistMatrix<T>& CPre, GemmAlgorithm alg) {
    AssertSameGrid(APre, CPre);
    AssertSameGrid(BPre, CPre);
    const Int m = CPre.Height(), n = CPre.Width();
    const Int am = (oA == NORMAL) ? APre.Height() : APre.Width();
    const Int ak = (oA == NORMAL) ? APre.Width() : APre.Height();
    const Int bk = (oB == NORMAL) ? BPre.Height() : BPre.Width();
    const Int bn = (oB == NORMAL) ? BPre.Width() : BPre.Height();
    if (am != m || bn != n || ak != bk) LogicError("Nonconformal matrices in Gemm");
    Scale(beta, CPre);  // Gemm.cpp:98
    if (m == 0 || n == 0 || ak == 0) return;
    ReadProxy<T> AP(APre), BP(BPre);
    ReadWriteProxy<T> CP(CPre);
    const auto& A = AP.Get();
    const auto& B = BP.Get();
    auto& C = CP.Get();
    if (alg == GEMM_DEFAULT) alg = GemmDefaultAlgorithm(m, n, ak);
    switch (alg) {
        case GEMM_SUMMA_A: SummaA(oA, oB, alpha, A, B, C); break;
        case GEMM_SUMMA_B: SummaB(oA, oB, alpha, A, B, C); break;
        case GEMM_SUMMA_C: SummaC(oA, oB, alpha, A, B, C); break;
        case GEMM_SUMMA_DOT: SummaDot(oA, oB, alpha, A, B, C); break;
        case GEMM_CANNON:   // only the NN case has it (Gemm/NN.hpp:283, NT/TN/TT reject it)
            if (oA != NORMAL || oB != NORMAL) LogicError("Unsupported Gemm option");
            CannonNN(alpha, A, B, C);
            break;
        default: LogicError("Unsupported Gemm option");
    }
    CP.Commit();
}
template <typename T>
void Gemm(Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
          AbstractDistMatrix<T>& C, GemmAlgorithm alg) {
    const Int m = (oA == NORMAL) ? A.Height() : A.Width();
    const Int n = (oB == NORMAL) ? B.Width() : B.Height();
    C.Resize(m, n);
    Zero(C);
    Gemm(oA, oB, alpha, A, B, T(0), C, alg);
}

// ---------------------------------------------------------------------------
// Trrk / Syrk / Herk
// ---------------------------------------------------------------------------
template <typename T>
void Trrk(UpperOrLower uplo, Orientation oA, Orientation oB, T alpha, const Matrix<T>& A, const Matrix<T>& B, T beta,
          Matrix<T>& C) {
    const Int n = C.Height();
    const Int k = (oA == NORMAL) ? A.Width() : A.Height();
    if (C.Width() != n || ((oA == NORMAL) ? A.Height() : A.Width()) != n ||
        ((oB == NORMAL) ? B.Width() : B.Height()) != n || ((oB == NORMAL) ? B.Height() : B.Width()) != k)
        LogicError("Nonconformal Trrk");
    if (n == 0) return;
    elb200::gemm_device<dev::D<T>>(uplo == LOWER ? 1 : 2, OrientationToChar(oA), OrientationToChar(oB), n, n, k,
                                   dev::val<T>(alpha), dev::ptr(A.LockedBuffer()), A.LDim(),
                                   dev::ptr(B.LockedBuffer()), B.LDim(), dev::val<T>(beta), dev::ptr(C.Buffer()),
                                   C.LDim(), 0, 1, 0, 1, dev::stream());
}

template <typename T>
void LocalTrrk(UpperOrLower uplo, Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& A,
               const AbstractDistMatrix<T>& B, T beta, AbstractDistMatrix<T>& C) {
    const Matrix<T>& Al = A.LockedMatrix();
    const Matrix<T>& Bl = B.LockedMatrix();
    Matrix<T>& Cl = C.Matrix();
    const Int m = Cl.Height(), n = Cl.Width();
    const Int k = (oA == NORMAL) ? Al.Width() : Al.Height();
    if (((oA == NORMAL) ? Al.Height() : Al.Width()) != m || ((oB == NORMAL) ? Bl.Width() : Bl.Height()) != n ||
        ((oB == NORMAL) ? Bl.Height() : Bl.Width()) != k)
        LogicError("Nonconformal LocalTrrk");
    if (m == 0 || n == 0) return;
    // one masked GEMM over the global staircase (replaces the recursion of Trrk/Local.hpp:782-830)
    elb200::gemm_device<dev::D<T>>(uplo == LOWER ? 1 : 2, OrientationToChar(oA), OrientationToChar(oB), m, n, k,
                                   dev::val<T>(alpha), dev::ptr(Al.LockedBuffer()), Al.LDim(),
                                   dev::ptr(Bl.LockedBuffer()), Bl.LDim(), dev::val<T>(beta), dev::ptr(Cl.Buffer()),
                                   Cl.LDim(), C.ColShift(), C.ColStride(), C.RowShift(), C.RowStride(), dev::stream());
}

template <typename T>
void Trrk(UpperOrLower uplo, Orientation oA, Orientation oB, T alpha, const AbstractDistMatrix<T>& APre,
          const AbstractDistMatrix<T>& BPre, T beta, AbstractDistMatrix<T>& CPre) {
    AssertSameGrid(APre, CPre);
    AssertSameGrid(BPre, CPre);
    const Int n = CPre.Height();
    const Int k = (oA == NORMAL) ? APre.Width() : APre.Height();
    if (CPre.Width() != n || ((oA == NORMAL) ? APre.Height() : APre.Width()) != n ||
        ((oB == NORMAL) ? BPre.Width() : BPre.Height()) != n || ((oB == NORMAL) ? BPre.Height() : BPre.Width()) != k)
        LogicError("Nonconformal Trrk");
    ReadProxy<T> AP(APre), BP(BPre);
    ReadWriteProxy<T> CP(CPre);
    const auto& A = AP.Get();
    const auto& B = BP.Get();
    auto& C = CP.Get();
    const Grid& g = C.Grid();
    ScaleTrapezoid(beta, uplo, C);
    const Int bsize = Blocksize();
    AbstractDistMatrix<T> A1(g, MC, STAR), B1(g, STAR, MR);
    A1.AlignWith(C);
    B1.AlignWith(C);
    for (Int s = 0; s < k; s += bsize) {
        const Int nb = std::min(bsize, k - s);
        FormPanel(oA, A, 0, s, n, nb, A1);
        FormPanel(oB, B, s, 0, nb, n, B1);
        LocalTrrk(uplo, NORMAL, NORMAL, alpha, A1, B1, T(1), C);
    }
    CP.Commit();
}

template <typename T>
void Syrk(UpperOrLower uplo, Orientation o, T alpha, const Matrix<T>& A, T beta, Matrix<T>& C, bool conjugate) {
    const Orientation other = conjugate ? ADJOINT : TRANSPOSE;
    if (o == NORMAL) Trrk(uplo, NORMAL, other, alpha, A, A, beta, C);
    else Trrk(uplo, other, NORMAL, alpha, A, A, beta, C);
}
namespace {
// Syrk with a long summation index (k > 10 n: syrk::LN_Dot / LT_Dot / UN_Dot / UT_Dot, Syrk/LN.hpp:88-128 and
// siblings): distribute the summation index over all p processes once, then per block of C's triangle one local
// product of the process's k / p slice and one sum-scatter.  Diagonal blocks are masked products into a zeroed
// block, so the strictly-other triangle of C only ever receives zeros.  The block edge is GemmDotBlocksize (the
// reference hard-codes 2000, a CPU-cache size).
template <typename T>
void SyrkDot(UpperOrLower uplo, Orientation o, T alpha, const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& C,
             bool conjugate) {
    const Grid& g = C.Grid();
    const Int n = C.Height();
    const Int blockSize = GemmDotBlocksize(sizeof(T));
    const Orientation other = conjugate ? ADJOINT : TRANSPOSE;
    const bool normal = (o == NORMAL);
    AbstractDistMatrix<T> AV(g, normal ? STAR : VC, normal ? VC : STAR);
    Copy(A, AV);
    const Int k = normal ? AV.Width() : AV.Height();
    AbstractDistMatrix<T> Z(g, STAR, STAR);
    auto rows = [&](Int i0, Int h) {   // op(AV)(i0 : i0 + h, :) as a view of AV
        return normal ? LockedView(static_cast<const AbstractDistMatrix<T>&>(AV), i0, 0, h, k)
                      : LockedView(static_cast<const AbstractDistMatrix<T>&>(AV), 0, i0, k, h);
    };
    for (Int i0 = 0; i0 < n; i0 += blockSize) {
        const Int bi = std::min(blockSize, n - i0);
        auto Ai = rows(i0, bi);
        // diagonal block
        Zeros(Z, bi, bi);
        if (normal) Trrk(uplo, NORMAL, other, alpha, Ai.LockedMatrix(), Ai.LockedMatrix(), T(0), Z.Matrix());
        else Trrk(uplo, other, NORMAL, alpha, Ai.LockedMatrix(), Ai.LockedMatrix(), T(0), Z.Matrix());
        {
            auto Cii = View(C, i0, i0, bi, bi);
            AxpyContract(T(1), static_cast<const AbstractDistMatrix<T>&>(Z), Cii);
        }
        // the blocks of this block row (UPPER) / block column (LOWER) beyond the diagonal
        for (Int j0 = i0 + bi; j0 < n; j0 += blockSize) {
            const Int bj = std::min(blockSize, n - j0);
            auto Aj = rows(j0, bj);
            if (uplo == LOWER) {   // C(J, I) += alpha op(A)_J op(A)_I'
                Z.Resize(bj, bi);
                if (normal) LocalGemm(NORMAL, other, alpha, Aj, Ai, T(0), Z);
                else LocalGemm(other, NORMAL, alpha, Aj, Ai, T(0), Z);
                auto Cji = View(C, j0, i0, bj, bi);
                AxpyContract(T(1), static_cast<const AbstractDistMatrix<T>&>(Z), Cji);
            } else {               // C(I, J) += alpha op(A)_I op(A)_J'
                Z.Resize(bi, bj);
                if (normal) LocalGemm(NORMAL, other, alpha, Ai, Aj, T(0), Z);
                else LocalGemm(other, NORMAL, alpha, Ai, Aj, T(0), Z);
                auto Cij = View(C, i0, j0, bi, bj);
                AxpyContract(T(1), static_cast<const AbstractDistMatrix<T>&>(Z), Cij);
            }
        }
    }
}
}  // namespace

template <typename T>
void Syrk(UpperOrLower uplo, Orientation o, T alpha, const AbstractDistMatrix<T>& A, T beta, AbstractDistMatrix<T>& C,
          bool conjugate) {
    const Orientation other = conjugate ? ADJOINT : TRANSPOSE;
    const Int n = C.Height();
    const Int k = (o == NORMAL) ? A.Width() : A.Height();
    const double weightAwayFromDot = 10.;   // Syrk/LN.hpp:152-157
    if (C.Width() == n && ((o == NORMAL) ? A.Height() : A.Width()) == n && k > weightAwayFromDot * n && n > 0) {
        AssertSameGrid(A, C);
        ReadWriteProxy<T> CP(C);
        ScaleTrapezoid(beta, uplo, CP.Get());
        SyrkDot(uplo, o, alpha, A, CP.Get(), conjugate);
        CP.Commit();
        return;
    }
    if (o == NORMAL) Trrk(uplo, NORMAL, other, alpha, A, A, beta, C);
    else Trrk(uplo, other, NORMAL, alpha, A, A, beta, C);
}
template <typename T>
void Herk(UpperOrLower uplo, Orientation o, Base<T> alpha, const Matrix<T>& A, Base<T> beta, Matrix<T>& C) {
    Syrk(uplo, o, T(alpha), A, T(beta), C, true);
}
template <typename T>
void Herk(UpperOrLower uplo, Orientation o, Base<T> alpha, const AbstractDistMatrix<T>& A, Base<T> beta,
          AbstractDistMatrix<T>& C) {
    Syrk(uplo, o, T(alpha), A, T(beta), C, true);
}


// ---------------------------------------------------------------------------
// Syr2k / Her2k, Symm / Hemm, Trmm: compositions over Trrk / Gemm / the redistribution engine
// ---------------------------------------------------------------------------
namespace {
template <typename T> T ConjIf(T a, bool c) { return a; }
template <> Complex<float> ConjIf(Complex<float> a, bool c) { return c ? std::conj(a) : a; }
template <> Complex<double> ConjIf(Complex<double> a, bool c) { return c ? std::conj(a) : a; }
}  // namespace

// ---- Trr2k: E_tri := alpha op(A) op(B) + beta op(C) op(D) + gamma E_tri ----
// The reference's LocalTrr2k (src/blas_like/level3/Trr2k/Local.hpp, 276 lines) recurses like LocalTrrk with two
// products per leaf.  Here the two products are ONE masked tensor-pipe GEMM: with L = [alpha op(A) | beta op(C)]
// and R = [op(B) ; op(D)] stacked along the summation index, L R = alpha op(A) op(B) + beta op(C) op(D), so E's
// triangle is read and written once per panel step instead of twice.  The stacking is two strided device copies of
// panel-sized data (alpha / beta, transposition and conjugation folded in).
template <typename T>
void LocalTrr2k(UpperOrLower uplo, Orientation oA, Orientation oB, Orientation oC, Orientation oD, T alpha,
                const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B, T beta, const AbstractDistMatrix<T>& C,
                const AbstractDistMatrix<T>& D, T gamma, AbstractDistMatrix<T>& E) {
    const Matrix<T>&Al = A.LockedMatrix(), &Bl = B.LockedMatrix(), &Cl = C.LockedMatrix(), &Dl = D.LockedMatrix();
    Matrix<T>& El_ = E.Matrix();
    const Int m = El_.Height(), n = El_.Width();
    const Int kA = (oA == NORMAL) ? Al.Width() : Al.Height(), kC = (oC == NORMAL) ? Cl.Width() : Cl.Height();
    auto rowsOf = [](Orientation o, const Matrix<T>& M) { return o == NORMAL ? M.Height() : M.Width(); };
    auto colsOf = [](Orientation o, const Matrix<T>& M) { return o == NORMAL ? M.Width() : M.Height(); };
    if (rowsOf(oA, Al) != m || rowsOf(oC, Cl) != m || colsOf(oB, Bl) != n || colsOf(oD, Dl) != n ||
        rowsOf(oB, Bl) != kA || rowsOf(oD, Dl) != kC)
        LogicError("Nonconformal LocalTrr2k");
    ScaleTrapezoid(gamma, uplo, E);
    if (m == 0 || n == 0 || kA + kC == 0) return;
    const Int kk = kA + kC;
    Matrix<T> Lm(m, kk), Rm(kk, n);
    typedef dev::D<T> DT;
    auto place = [&](Orientation o, const Matrix<T>& src, Matrix<T>& dst, Int i0, Int j0, Int h, Int w, const T* scale) {
        // dst(i0 : i0 + h, j0 : j0 + w) = scale * op(src)
        if (h == 0 || w == 0) return;
        DT sc = dev::val<T>(scale ? *scale : T(1));
        const bool tr = (o != NORMAL);
        elb200::lattice_copy_device<DT>(dev::ptr(src.LockedBuffer()), dev::ptr(dst.Buffer()), h, w, 0, tr ? src.LDim() : 1,
                                        tr ? 1 : src.LDim(), i0 + (elb200::i64)j0 * dst.LDim(), 1, dst.LDim(),
                                        o == ADJOINT, scale ? &sc : nullptr, false, dev::stream());
    };
    place(oA, Al, Lm, 0, 0, m, kA, alpha == T(1) ? nullptr : &alpha);
    place(oC, Cl, Lm, 0, kA, m, kC, beta == T(1) ? nullptr : &beta);
    place(oB, Bl, Rm, 0, 0, kA, n, nullptr);
    place(oD, Dl, Rm, kA, 0, kC, n, nullptr);
    elb200::gemm_device<DT>(uplo == LOWER ? 1 : 2, 'N', 'N', m, n, kk, dev::val<T>(T(1)), dev::ptr(Lm.LockedBuffer()),
                            Lm.LDim(), dev::ptr(Rm.LockedBuffer()), Rm.LDim(), dev::val<T>(T(1)), dev::ptr(El_.Buffer()),
                            El_.LDim(), E.ColShift(), E.ColStride(), E.RowShift(), E.RowStride(), dev::stream());
}

// Trr2k.cpp:34-...: panel loop at Blocksize(); every orientation case reduces to "form the four panels in
// [MC,*] / [*,MR], one LocalTrr2k".
template <typename T>
void Trr2k(UpperOrLower uplo, Orientation oA, Orientation oB, Orientation oC, Orientation oD, T alpha,
           const AbstractDistMatrix<T>& APre, const AbstractDistMatrix<T>& BPre, T beta,
           const AbstractDistMatrix<T>& CPre, const AbstractDistMatrix<T>& DPre, T gamma, AbstractDistMatrix<T>& EPre) {
    AssertSameGrid(APre, EPre); AssertSameGrid(BPre, EPre); AssertSameGrid(CPre, EPre); AssertSameGrid(DPre, EPre);
    const Int n = EPre.Height();
    const Int k = (oA == NORMAL) ? APre.Width() : APre.Height();
    auto dimsOk = [&](Orientation o, const AbstractDistMatrix<T>& M, bool left) {
        const Int r = (o == NORMAL) ? M.Height() : M.Width(), c = (o == NORMAL) ? M.Width() : M.Height();
        return left ? (r == n && c == k) : (r == k && c == n);
    };
    if (EPre.Width() != n || !dimsOk(oA, APre, true) || !dimsOk(oC, CPre, true) || !dimsOk(oB, BPre, false) ||
        !dimsOk(oD, DPre, false))
        LogicError("Nonconformal Trr2k");
    ReadProxy<T> AP(APre), BP(BPre), CP(CPre), DP(DPre);
    ReadWriteProxy<T> EP(EPre);
    const auto &A = AP.Get(), &B = BP.Get(), &C = CP.Get(), &D = DP.Get();
    auto& E = EP.Get();
    const Grid& g = E.Grid();
    ScaleTrapezoid(gamma, uplo, E);
    const Int bsize = Blocksize();
    AbstractDistMatrix<T> A1(g, MC, STAR), C1(g, MC, STAR), B1(g, STAR, MR), D1(g, STAR, MR);
    A1.AlignWith(E); C1.AlignWith(E); B1.AlignWith(E); D1.AlignWith(E);
    for (Int s = 0; s < k; s += bsize) {
        const Int nb = std::min(bsize, k - s);
        FormPanel(oA, A, 0, s, n, nb, A1);
        FormPanel(oC, C, 0, s, n, nb, C1);
        FormPanel(oB, B, s, 0, nb, n, B1);
        FormPanel(oD, D, s, 0, nb, n, D1);
        LocalTrr2k(uplo, NORMAL, NORMAL, NORMAL, NORMAL, alpha, A1, B1, beta, C1, D1, T(1), E);
    }
    EP.Commit();
}

// C_tri := alpha op(A) op(B)' + alphaSec op(B) op(A)' + beta C_tri, alphaSec = conj(alpha) for Her2k
// (Syr2k/LN.hpp:14-63 and siblings): one Trr2k, i.e. one pass over the triangle of C per panel step.
template <typename T>
void Syr2k(UpperOrLower uplo, Orientation o, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
           T beta, AbstractDistMatrix<T>& C, bool conjugate) {
    const Int n = C.Height();
    const bool normal = (o == NORMAL);
    if (C.Width() != n || (normal ? A.Height() : A.Width()) != n || (normal ? B.Height() : B.Width()) != n ||
        (normal ? A.Width() : A.Height()) != (normal ? B.Width() : B.Height()))
        LogicError("Nonconformal Syr2k");
    const Orientation other = conjugate ? ADJOINT : TRANSPOSE;
    const T alphaSec = ConjIf(alpha, conjugate);
    if (normal) Trr2k(uplo, NORMAL, other, NORMAL, other, alpha, A, B, alphaSec, B, A, beta, C);
    else Trr2k(uplo, other, NORMAL, other, NORMAL, alpha, A, B, alphaSec, B, A, beta, C);
}
template <typename T>
void Her2k(UpperOrLower uplo, Orientation o, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
           Base<T> beta, AbstractDistMatrix<T>& C) {
    Syr2k(uplo, o, alpha, A, B, T(beta), C, true);
}

// C := alpha A B + beta C (LEFT) / alpha B A + beta C (RIGHT); only the uplo triangle of A is read.
// F = tri(A) + strict(tri(A))' is formed once on the device (one transposing redistribution), then
// the product is an ordinary SUMMA -- all of its flops on the tensor pipe, no triangular special cases.
template <typename T>
void Symm(LeftOrRight side, UpperOrLower uplo, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
          T beta, AbstractDistMatrix<T>& C, bool conjugate) {
    AssertSameGrid(A, C);
    AssertSameGrid(B, C);
    const Int m = C.Height(), n = C.Width(), ka = (side == LEFT) ? m : n;
    if (A.Height() != ka || A.Width() != ka || B.Height() != m || B.Width() != n) LogicError("Nonconformal Symm");
    const Grid& g = C.Grid();
    AbstractDistMatrix<T> F(g, MC, MR), Ft(g, MC, MR);
    Copy(A, F);
    MakeTrapezoidal(uplo, F, 0);
    Ft.AlignWith(F);
    Transpose(static_cast<const AbstractDistMatrix<T>&>(F), Ft, conjugate);
    // keep only the strict part of the mirrored triangle
    MakeTrapezoidal(uplo == LOWER ? UPPER : LOWER, Ft, uplo == LOWER ? 1 : -1);
    Axpy(T(1), static_cast<const AbstractDistMatrix<T>&>(Ft), F);
    Ft.Empty();
    if (side == LEFT) Gemm(NORMAL, NORMAL, alpha, static_cast<const AbstractDistMatrix<T>&>(F), B, beta, C, GEMM_DEFAULT);
    else Gemm(NORMAL, NORMAL, alpha, B, static_cast<const AbstractDistMatrix<T>&>(F), beta, C, GEMM_DEFAULT);
}
template <typename T>
void Hemm(LeftOrRight side, UpperOrLower uplo, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
          T beta, AbstractDistMatrix<T>& C) {
    Symm(side, uplo, alpha, A, B, beta, C, true);
}

// B := alpha op(tri(A)) B (LEFT) / alpha B op(tri(A)) (RIGHT).  UNIT: the stored diagonal is never read --
// the strict triangle multiplies and alpha B is added back.
template <typename T>
void Trmm(LeftOrRight side, UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, T alpha,
          const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B) {
    AssertSameGrid(A, B);
    const Int m = B.Height(), n = B.Width(), ka = (side == LEFT) ? m : n;
    if (A.Height() != ka || A.Width() != ka) LogicError("Nonconformal Trmm");
    const Grid& g = B.Grid();
    AbstractDistMatrix<T> Tm(g, MC, MR), B0(g, MC, MR);
    Copy(A, Tm);
    const bool unit = (diag == UNIT);
    MakeTrapezoidal(uplo, Tm, unit ? (uplo == LOWER ? -1 : 1) : 0);
    B0.AlignWith(B);
    Copy(static_cast<const AbstractDistMatrix<T>&>(B), B0);
    const T beta = unit ? alpha : T(0);   // unit diagonal: B := alpha (strict B0 + B0)
    if (unit) { /* B already holds B0; it is scaled by beta = alpha inside Gemm */ }
    if (side == LEFT)
        Gemm(o, NORMAL, alpha, static_cast<const AbstractDistMatrix<T>&>(Tm), static_cast<const AbstractDistMatrix<T>&>(B0),
             beta, B, GEMM_DEFAULT);
    else
        Gemm(NORMAL, o, alpha, static_cast<const AbstractDistMatrix<T>&>(B0), static_cast<const AbstractDistMatrix<T>&>(Tm),
             beta, B, GEMM_DEFAULT);
}

// ---- TwoSidedTrsm / TwoSidedTrmm (src/blas_like/level3/TwoSidedTrsm.cpp, TwoSidedTrmm.cpp, */LVar4.hpp, */UVar4.hpp) ----
//   TwoSidedTrsm: A := inv(L) A inv(L)^H (LOWER) / inv(U)^H A inv(U) (UPPER);  TwoSidedTrmm: A := L^H A L / U A U^H,
// A Hermitian with only its `uplo` triangle stored and updated.  The reference's variant-4 loops exploit the
// symmetry (n^3 flops) with eight distributed temporaries per step; here the Hermitian matrix is completed once
// (one transposing redistribution, as Hemm does) and the two one-sided operations run on the tensor pipe through
// the Trsm / Trmm of this file -- 2 n^3 flops, every one of them in a large GEMM -- and the `uplo` triangle of the
// result is written back; the other triangle of A is left untouched, as in the reference.
namespace {
template <typename F>
void TwoSided(bool solve, UpperOrLower uplo, UnitOrNonUnit diag, AbstractDistMatrix<F>& A, const AbstractDistMatrix<F>& B) {
    AssertSameGrid(A, B);
    const Int n = A.Height();
    if (A.Width() != n || B.Height() != n || B.Width() != n) LogicError("Nonconformal two-sided triangular operation");
    if (n == 0) return;
    const Grid& g = A.Grid();
    AbstractDistMatrix<F> H(g, MC, MR), Ht(g, MC, MR);
    Copy(static_cast<const AbstractDistMatrix<F>&>(A), H);
    MakeTrapezoidal(uplo, H, 0);
    Ht.AlignWith(H);
    Transpose(static_cast<const AbstractDistMatrix<F>&>(H), Ht, true);
    MakeTrapezoidal(uplo == LOWER ? UPPER : LOWER, Ht, uplo == LOWER ? 1 : -1);
    Axpy(F(1), static_cast<const AbstractDistMatrix<F>&>(Ht), H);
    Ht.Empty();
    if (solve) {
        if (uplo == LOWER) { Trsm(LEFT, LOWER, NORMAL, diag, F(1), B, H); Trsm(RIGHT, LOWER, ADJOINT, diag, F(1), B, H); }
        else { Trsm(LEFT, UPPER, ADJOINT, diag, F(1), B, H); Trsm(RIGHT, UPPER, NORMAL, diag, F(1), B, H); }
    } else {
        if (uplo == LOWER) { Trmm(LEFT, LOWER, ADJOINT, diag, F(1), B, H); Trmm(RIGHT, LOWER, NORMAL, diag, F(1), B, H); }
        else { Trmm(LEFT, UPPER, NORMAL, diag, F(1), B, H); Trmm(RIGHT, UPPER, ADJOINT, diag, F(1), B, H); }
    }
    ScaleTrapezoid(F(0), uplo, A);
    AxpyTrapezoid(uplo, F(1), static_cast<const AbstractDistMatrix<F>&>(H), A);
}
}  // namespace
template <typename F>
void TwoSidedTrsm(UpperOrLower uplo, UnitOrNonUnit diag, AbstractDistMatrix<F>& A, const AbstractDistMatrix<F>& B) {
    TwoSided(true, uplo, diag, A, B);
}
template <typename F>
void TwoSidedTrmm(UpperOrLower uplo, UnitOrNonUnit diag, AbstractDistMatrix<F>& A, const AbstractDistMatrix<F>& B) {
    TwoSided(false, uplo, diag, A, B);
}

// ---------------------------------------------------------------------------
// Trsm
// ---------------------------------------------------------------------------
namespace {
// the zero-diagonal scan of Trsm.cpp:54-60 on the device: flag <- offset + j + 1 for the first A(j,j) == 0
template <typename F>
void ScanDiagonal(const Matrix<F>& A, int* flagDev, Int offset) {
    dev::c_check(elb200_diag_zero_check(dev::Code<F>(), A.Height(), A.LockedBuffer(), A.LDim(), offset, flagDev,
                                        (elb200_stream_t)dev::stream()),
                 "elb200_diag_zero_check");
}
template <typename F>
void LocalTrsmRaw(LeftOrRight side, UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, F alpha, const Matrix<F>& A,
                  Matrix<F>& B) {
    const Int na = (side == LEFT) ? B.Height() : B.Width();
    if (A.Height() != A.Width()) LogicError("Triangular matrix must be square");
    if (A.Height() != na) LogicError("Nonconformal Trsm");
    elb200::trsm_device<dev::D<F>>(LeftOrRightToChar(side), UpperOrLowerToChar(uplo), OrientationToChar(o),
                                   UnitOrNonUnitToChar(diag), B.Height(), B.Width(), dev::val<F>(alpha),
                                   dev::ptr(A.LockedBuffer()), A.LDim(), dev::ptr(B.Buffer()), B.LDim(), dev::stream());
}
}  // namespace

template <typename F>
void Trsm(LeftOrRight side, UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, F alpha, const Matrix<F>& A,
          Matrix<F>& B, bool checkIfSingular) {
    if (checkIfSingular && diag != UNIT) {
        // Trsm.cpp:54-60 throws before the solve
        dev::DeviceFlag flag;
        ScanDiagonal(A, flag.dev_, 0);
        if (flag.Read() != 0) throw SingularMatrixException();
    }
    LocalTrsmRaw(side, uplo, o, diag, alpha, A, B);
}
template <typename F>
void LocalTrsm(LeftOrRight side, UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, F alpha,
               const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& X, bool checkIfSingular) {
    if (A.ColDist() != STAR || A.RowDist() != STAR) LogicError("LocalTrsm needs a [*,*] triangular matrix");
    if (side == LEFT && X.ColDist() != STAR) LogicError("Dist of RHS must conform with that of triangle");
    if (side == RIGHT && X.RowDist() != STAR) LogicError("Dist of RHS must conform with that of triangle");
    Trsm(side, uplo, o, diag, alpha, A.LockedMatrix(), X.Matrix(), checkIfSingular);
}

namespace {

// One block step of a distributed triangular solve needs L11 = tri(k:k+nb, k:k+nb) replicated and the panel of the
// triangle that couples block k to the blocks not yet solved.  Both depend only on the (read-only) triangle, so they
// are formed one step AHEAD on the panel stream (double-buffered), underneath the previous step's update.
template <typename F>
struct TrsmPanels {
    const AbstractDistMatrix<F>& L;
    const Grid& g;
    bool overlap;
    cudaStream_t mainS, panelS;
    AbstractDistMatrix<F> L11[2];
    AbstractDistMatrix<F> Lp[2];
    dev::Event ready[2], freed[2], fork;
    int issued = 0;
    TrsmPanels(const AbstractDistMatrix<F>& L_, Dist pU, Dist pV, bool wantOverlap)
        : L(L_), g(L_.Grid()), overlap(wantOverlap && dev::OverlapEnabled()),
          mainS(dev::stream()), panelS(overlap ? elb200::aux_stream(0) : dev::stream()),
          L11{AbstractDistMatrix<F>(g, STAR, STAR), AbstractDistMatrix<F>(g, STAR, STAR)},
          Lp{AbstractDistMatrix<F>(g, pU, pV), AbstractDistMatrix<F>(g, pU, pV)} {
        if (overlap) { fork.Record(mainS); fork.Wait(panelS); }
    }
    // enqueue the panels of block (k, nb); (pi, pj, ph, pw) = the coupling panel of the triangle, aligned with `like`
    void Issue(Int k, Int nb, Int pi, Int pj, Int ph, Int pw, const AbstractDistMatrix<F>& like) {
        const int s = issued & 1;
        dev::StreamScope onPanel(panelS);
        if (overlap && issued >= 2) freed[s].Wait(panelS);
        auto L11v = LockedView(L, k, k, nb, nb);
        Copy(static_cast<const AbstractDistMatrix<F>&>(L11v), L11[s]);
        if (ph > 0 && pw > 0) {
            Lp[s].AlignWith(like);
            auto Lv = LockedView(L, pi, pj, ph, pw);
            Copy(static_cast<const AbstractDistMatrix<F>&>(Lv), Lp[s]);
        }
        if (overlap) ready[s].Record(panelS);
        ++issued;
    }
    void Acquire(int s) { if (overlap) ready[s].Wait(mainS); }
    void Release(int s) { if (overlap) freed[s].Record(mainS); }
    ~TrsmPanels() {
        if (overlap) {
            // the panel buffers are released in main-stream order: make sure the panel stream is done with them
            dev::Event join;
            join.Record(panelS);
            join.Wait(mainS);
        }
    }
};

// LEFT solves, "Large" and "Medium" (Trsm/LLN.hpp:18-126, LLT.hpp:20-141 and the LUN / LUT mirrors): they differ
// only in how the solved block row travels.  Large: X1[*,VR] (solve spread over all p processes) then gathered to
// [*,MR]; Medium: X1^T[MR,*] (one transposing redistribution, solve on the right, no second gather).
template <typename F>
void TrsmLeft(UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, const AbstractDistMatrix<F>& L,
              AbstractDistMatrix<F>& X, int* singularFlag, bool medium) {
    const Grid& g = X.Grid();
    // The block loop runs on ELB200_TRSM_BLOCK_FACTOR (default 4) x Blocksize() rows at a time: the solve is a chain of
    // dependent steps (three redistributions, a small solve and a rank-nb product each), latency-bound long before it
    // is flop-bound -- at Blocksize() = 128 and 1024 right-hand sides on 2x4 GPUs a step costs ~0.4 ms for 0.1 ms of
    // arithmetic.  Four blocks per step pay that chain a quarter as often; the diagonal block grows to 512 x 512,
    // which the local kernel solves in its own 32-wide sweeps.  The result is the same up to the order of the sums.
    const Int factor = [] { const char* e = std::getenv("ELB200_TRSM_BLOCK_FACTOR"); const int v = e ? std::atoi(e) : 4; return v >= 1 ? v : 1; }();
    const Int mTri = L.Height(), bsize = Blocksize() * factor, nrhs = X.Width();
    const bool effLower = (uplo == LOWER) == (o == NORMAL);
    const bool forward = effLower;
    const Int nblk = (mTri + bsize - 1) / bsize;
    TrsmPanels<F> panels(L, o == NORMAL ? MC : STAR, o == NORMAL ? STAR : MC, g.Size() > 1 && nblk > 2);
    AbstractDistMatrix<F> X1_STAR_VR(g, STAR, VR), X1_STAR_MR(g, STAR, MR), X1T_MR_STAR(g, MR, STAR);
    auto block = [&](Int step, Int& k, Int& nb, Int& r0, Int& rl) {
        const Int kb = forward ? step : nblk - 1 - step;
        k = kb * bsize;
        nb = std::min(bsize, mTri - k);
        r0 = forward ? k + nb : 0;            // the not-yet-solved part
        rl = forward ? mTri - (k + nb) : k;
    };
    auto issue = [&](Int step) {
        Int k, nb, r0, rl;
        block(step, k, nb, r0, rl);
        auto X2 = LockedView(static_cast<const AbstractDistMatrix<F>&>(X), r0, 0, rl, nrhs);
        if (o == NORMAL) panels.Issue(k, nb, r0, k, rl, nb, X2);   // L(rest, blk)[MC,*]
        else panels.Issue(k, nb, k, r0, nb, rl, X2);               // L(blk, rest)[*,MC]
    };
    issue(0);
    for (Int step = 0; step < nblk; ++step) {
        Int k, nb, r0, rl;
        block(step, k, nb, r0, rl);
        const int s = (int)(step & 1);
        if (step + 1 < nblk) issue(step + 1);
        auto X1 = View(X, k, 0, nb, nrhs);
        panels.Acquire(s);
        if (singularFlag && diag != UNIT) ScanDiagonal(panels.L11[s].LockedMatrix(), singularFlag, k);
        const Orientation oX = medium ? (o == ADJOINT ? ADJOINT : TRANSPOSE) : NORMAL;
        if (!medium) {
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1), X1_STAR_VR);
            LocalTrsmRaw(LEFT, uplo, o, diag, F(1), panels.L11[s].LockedMatrix(), X1_STAR_VR.Matrix());
            if (rl > 0) {
                auto X2 = View(X, r0, 0, rl, nrhs);
                X1_STAR_MR.AlignWith(X2);
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_STAR_VR), X1_STAR_MR);
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_STAR_MR), X1);
            } else {
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_STAR_VR), X1);
            }
        } else {
            // X1^[T/H][MR,*] := X1^[T/H][MR,*] op'(L11)^-1, op' = transpose for NORMAL, identity otherwise
            X1T_MR_STAR.AlignWith(X);
            Transpose(static_cast<const AbstractDistMatrix<F>&>(X1), X1T_MR_STAR, o == ADJOINT);
            LocalTrsmRaw(RIGHT, uplo, o == NORMAL ? TRANSPOSE : NORMAL, diag, F(1), panels.L11[s].LockedMatrix(),
                         X1T_MR_STAR.Matrix());
            Transpose(static_cast<const AbstractDistMatrix<F>&>(X1T_MR_STAR), X1, o == ADJOINT);
        }
        if (rl > 0) {
            auto X2 = View(X, r0, 0, rl, nrhs);
            const AbstractDistMatrix<F>& Xrep = medium ? X1T_MR_STAR : X1_STAR_MR;
            LocalGemm(o == NORMAL ? NORMAL : o, oX, F(-1), panels.Lp[s], Xrep, F(1), X2);
        }
        panels.Release(s);
    }
}

// LEFT, "Small" (width(X) << p; LLN.hpp:129-175, LLT.hpp:144-252): triangle and right-hand sides both [VC,*], no
// communication in the updates.  NORMAL: solve a block, update the rest (right-looking).  (Conjugate-)transposed:
// the block first receives A(rest, blk)^[T/H] X(rest) -- partial sums on every process, summed over all of them --
// and is then solved (left-looking).
template <typename F>
void TrsmLeftSmall(UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, const AbstractDistMatrix<F>& APre,
                   AbstractDistMatrix<F>& XPre, int* singularFlag) {
    const Grid& g = XPre.Grid();
    AbstractDistMatrix<F> A(g, VC, STAR), X(g, VC, STAR);
    Copy(APre, A);
    X.AlignCols(A.ColAlign());
    Copy(static_cast<const AbstractDistMatrix<F>&>(XPre), X);
    const Int mTri = A.Height(), bsize = Blocksize(), nrhs = X.Width();
    const bool forward = (uplo == LOWER) == (o == NORMAL);
    const Int nblk = (mTri + bsize - 1) / bsize;
    AbstractDistMatrix<F> A11(g, STAR, STAR), X1s(g, STAR, STAR), Z1(g, STAR, STAR);
    for (Int step = 0; step < nblk; ++step) {
        const Int kb = forward ? step : nblk - 1 - step;
        const Int k = kb * bsize, nb = std::min(bsize, mTri - k);
        const Int r0 = forward ? k + nb : 0, rl = forward ? mTri - (k + nb) : k;       // not yet solved
        const Int d0 = forward ? 0 : k + nb, dl = forward ? k : mTri - (k + nb);       // already solved
        auto A11v = LockedView(static_cast<const AbstractDistMatrix<F>&>(A), k, k, nb, nb);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A11v), A11);
        if (singularFlag && diag != UNIT) ScanDiagonal(A11.LockedMatrix(), singularFlag, k);
        auto X1 = View(X, k, 0, nb, nrhs);
        if (o == NORMAL) {
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1), X1s);
            LocalTrsmRaw(LEFT, uplo, o, diag, F(1), A11.LockedMatrix(), X1s.Matrix());
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1s), X1);
            if (rl > 0) {
                auto A21 = LockedView(static_cast<const AbstractDistMatrix<F>&>(A), r0, k, rl, nb);
                auto X2 = View(X, r0, 0, rl, nrhs);
                LocalGemm(NORMAL, NORMAL, F(-1), A21, X1s, F(1), X2);
            }
        } else {
            if (dl > 0) {
                auto Ad1 = LockedView(static_cast<const AbstractDistMatrix<F>&>(A), d0, k, dl, nb);
                auto Xd = LockedView(static_cast<const AbstractDistMatrix<F>&>(X), d0, 0, dl, nrhs);
                Z1.Resize(nb, nrhs);
                LocalGemm(o, NORMAL, F(-1), Ad1, Xd, F(0), Z1);
                AxpyContract(F(1), static_cast<const AbstractDistMatrix<F>&>(Z1), X1);   // X1 += sum over processes
            }
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1), X1s);
            LocalTrsmRaw(LEFT, uplo, o, diag, F(1), A11.LockedMatrix(), X1s.Matrix());
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1s), X1);
            (void)rl; (void)r0;
        }
    }
    Copy(static_cast<const AbstractDistMatrix<F>&>(X), XPre);
}

// RIGHT solves (Trsm/RLN.hpp, RLT.hpp, RUN.hpp, RUT.hpp)
template <typename F>
void TrsmRight(UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, const AbstractDistMatrix<F>& L,
               AbstractDistMatrix<F>& X, int* singularFlag) {
    const Grid& g = X.Grid();
    const Int mTri = L.Height(), bsize = Blocksize(), xm = X.Height();
    const bool effLower = (uplo == LOWER) == (o == NORMAL);
    const bool forward = !effLower;
    const Int nblk = (mTri + bsize - 1) / bsize;
    TrsmPanels<F> panels(L, o == NORMAL ? STAR : MR, o == NORMAL ? MR : STAR, g.Size() > 1 && nblk > 2);
    AbstractDistMatrix<F> X1_VC_STAR(g, VC, STAR), X1_MC_STAR(g, MC, STAR);
    auto block = [&](Int step, Int& k, Int& nb, Int& r0, Int& rl) {
        const Int kb = forward ? step : nblk - 1 - step;
        k = kb * bsize;
        nb = std::min(bsize, mTri - k);
        r0 = forward ? k + nb : 0;
        rl = forward ? mTri - (k + nb) : k;
    };
    auto issue = [&](Int step) {
        Int k, nb, r0, rl;
        block(step, k, nb, r0, rl);
        auto X2 = LockedView(static_cast<const AbstractDistMatrix<F>&>(X), 0, r0, xm, rl);
        if (o == NORMAL) panels.Issue(k, nb, k, r0, nb, rl, X2);   // L(blk, rest)[*,MR]
        else panels.Issue(k, nb, r0, k, rl, nb, X2);               // L(rest, blk)[MR,*]
    };
    issue(0);
    for (Int step = 0; step < nblk; ++step) {
        Int k, nb, r0, rl;
        block(step, k, nb, r0, rl);
        const int s = (int)(step & 1);
        if (step + 1 < nblk) issue(step + 1);
        auto X1 = View(X, 0, k, xm, nb);
        panels.Acquire(s);
        if (singularFlag && diag != UNIT) ScanDiagonal(panels.L11[s].LockedMatrix(), singularFlag, k);
        Copy(static_cast<const AbstractDistMatrix<F>&>(X1), X1_VC_STAR);
        LocalTrsmRaw(RIGHT, uplo, o, diag, F(1), panels.L11[s].LockedMatrix(), X1_VC_STAR.Matrix());
        if (rl > 0) {
            auto X2 = View(X, 0, r0, xm, rl);
            X1_MC_STAR.AlignWith(X2);
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1_VC_STAR), X1_MC_STAR);
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1_MC_STAR), X1);
            LocalGemm(NORMAL, o == NORMAL ? NORMAL : o, F(-1), X1_MC_STAR, panels.Lp[s], F(1), X2);
        } else {
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1_VC_STAR), X1);
        }
        panels.Release(s);
    }
}

}  // namespace

// x := op(tri(A))^-1 x for a single column (src/blas_like/level2/Trsv.cpp:47-68, Trsv/{LN,LT,UN,UT}.hpp).  The
// reference accumulates the updates in a [MC,*] / [*,MC] vector and reduces them block by block; here the column
// runs through the Medium block substitution (its x1 is replicated as x1^T[MR,*], the update is a local
// matrix-vector product on the tensor pipe) -- the same arithmetic per block, one code path.
template <typename F>
void Trsv(UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, const AbstractDistMatrix<F>& APre,
          AbstractDistMatrix<F>& xPre) {
    AssertSameGrid(APre, xPre);
    if (APre.Height() != APre.Width()) LogicError("A must be square");
    if (xPre.Width() != 1 && xPre.Height() != 1) LogicError("x must be a vector");
    const bool column = xPre.Width() == 1;
    if ((column ? xPre.Height() : xPre.Width()) != APre.Height()) LogicError("x must conform with A");
    ReadProxy<F> AP(APre);
    if (column) {
        ReadWriteProxy<F> XP(xPre);
        TrsmLeft(uplo, o, diag, AP.Get(), XP.Get(), nullptr, true);
        XP.Commit();
    } else {
        // a row vector: solve for its transpose
        AbstractDistMatrix<F> xt(xPre.Grid(), MC, MR);
        Transpose(static_cast<const AbstractDistMatrix<F>&>(xPre), xt, false);
        TrsmLeft(uplo, o, diag, AP.Get(), xt, nullptr, true);
        Transpose(static_cast<const AbstractDistMatrix<F>&>(xt), xPre, false);
    }
}

// Algorithm selection as Trsm.cpp:94-375: width-1 left solves go to Trsv; LEFT: Large when width(B) > 5 p, else
// Medium, Small only on request; RIGHT: the default algorithm only.
template <typename F>
void Trsm(LeftOrRight side, UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, F alpha,
          const AbstractDistMatrix<F>& APre, AbstractDistMatrix<F>& BPre, bool checkIfSingular, TrsmAlgorithm alg) {
    AssertSameGrid(APre, BPre);
    if (APre.Height() != APre.Width()) LogicError("A must be square");
    if ((side == LEFT ? BPre.Height() : BPre.Width()) != APre.Height()) LogicError("Nonconformal Trsm");
    if (side == RIGHT && alg != TRSM_DEFAULT) LogicError("Unsupported TRSM algorithm");
    Scale(alpha, BPre);  // Trsm.cpp:94
    if (APre.Height() == 0 || BPre.Height() == 0 || BPre.Width() == 0) return;
    if (side == LEFT && BPre.Width() == 1 && !checkIfSingular) {
        Trsv(uplo, o, diag, APre, BPre);   // Trsm.cpp:97-101
        return;
    }
    std::unique_ptr<dev::DeviceFlag> flag;
    if (checkIfSingular && diag != UNIT) flag.reset(new dev::DeviceFlag());
    int* fdev = flag ? flag->dev_ : nullptr;
    const Int p = BPre.Grid().Size();
    if (side == LEFT && alg == TRSM_SMALL) {
        TrsmLeftSmall(uplo, o, diag, APre, BPre, fdev);
    } else {
        ReadProxy<F> AP(APre);
        ReadWriteProxy<F> XP(BPre);
        if (side == LEFT) {
            const bool medium = (alg == TRSM_MEDIUM) || (alg == TRSM_DEFAULT && !(BPre.Width() > 5 * p));
            TrsmLeft(uplo, o, diag, AP.Get(), XP.Get(), fdev, medium);
        } else {
            TrsmRight(uplo, o, diag, AP.Get(), XP.Get(), fdev);
        }
        XP.Commit();
    }
    // the diagonal blocks are replicated, so every process reads the same flag (no deadlock); the reference throws
    // at the first zero diagonal entry, before that block is solved (Trsm.cpp:54-60)
    if (flag && flag->Read() != 0) throw SingularMatrixException();
}

#define ELB_INST(T)                                                                                                  \
    template void Gemm(Orientation, Orientation, T, const Matrix<T>&, const Matrix<T>&, T, Matrix<T>&);              \
    template void Gemm(Orientation, Orientation, T, const Matrix<T>&, const Matrix<T>&, Matrix<T>&);                 \
    template void Gemm(Orientation, Orientation, T, const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, T,   \
                       AbstractDistMatrix<T>&, GemmAlgorithm);                                                       \
    template void Gemm(Orientation, Orientation, T, const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&,      \
                       AbstractDistMatrix<T>&, GemmAlgorithm);                                                       \
    template void LocalGemm(Orientation, Orientation, T, const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, \
                            T, AbstractDistMatrix<T>&);                                                              \
    template void LocalGemm(Orientation, Orientation, T, const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, \
                            AbstractDistMatrix<T>&);                                                                 \
    template void Trrk(UpperOrLower, Orientation, Orientation, T, const Matrix<T>&, const Matrix<T>&, T, Matrix<T>&); \
    template void Trrk(UpperOrLower, Orientation, Orientation, T, const AbstractDistMatrix<T>&,                      \
                       const AbstractDistMatrix<T>&, T, AbstractDistMatrix<T>&);                                     \
    template void LocalTrrk(UpperOrLower, Orientation, Orientation, T, const AbstractDistMatrix<T>&,                 \
                            const AbstractDistMatrix<T>&, T, AbstractDistMatrix<T>&);                                \
    template void Syrk(UpperOrLower, Orientation, T, const Matrix<T>&, T, Matrix<T>&, bool);                         \
    template void Syrk(UpperOrLower, Orientation, T, const AbstractDistMatrix<T>&, T, AbstractDistMatrix<T>&, bool); \
    template void Herk(UpperOrLower, Orientation, Base<T>, const Matrix<T>&, Base<T>, Matrix<T>&);                   \
    template void Herk(UpperOrLower, Orientation, Base<T>, const AbstractDistMatrix<T>&, Base<T>,                    \
                       AbstractDistMatrix<T>&);                                                                      \
    template void Syr2k(UpperOrLower, Orientation, T, const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, T, \
                        AbstractDistMatrix<T>&, bool);                                                               \
    template void Trr2k(UpperOrLower, Orientation, Orientation, Orientation, Orientation, T,                        \
                        const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, T, const AbstractDistMatrix<T>&, \
                        const AbstractDistMatrix<T>&, T, AbstractDistMatrix<T>&);                                    \
    template void LocalTrr2k(UpperOrLower, Orientation, Orientation, Orientation, Orientation, T,                   \
                             const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, T,                          \
                             const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, T, AbstractDistMatrix<T>&); \
    template void Her2k(UpperOrLower, Orientation, T, const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&,    \
                        Base<T>, AbstractDistMatrix<T>&);                                                            \
    template void Symm(LeftOrRight, UpperOrLower, T, const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, T,  \
                       AbstractDistMatrix<T>&, bool);                                                                \
    template void Hemm(LeftOrRight, UpperOrLower, T, const AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&, T,  \
                       AbstractDistMatrix<T>&);                                                                      \
    template void Trmm(LeftOrRight, UpperOrLower, Orientation, UnitOrNonUnit, T, const AbstractDistMatrix<T>&,       \
                       AbstractDistMatrix<T>&);                                                                      \
    template void Trsm(LeftOrRight, UpperOrLower, Orientation, UnitOrNonUnit, T, const Matrix<T>&, Matrix<T>&, bool); \
    template void Trsm(LeftOrRight, UpperOrLower, Orientation, UnitOrNonUnit, T, const AbstractDistMatrix<T>&,       \
                       AbstractDistMatrix<T>&, bool, TrsmAlgorithm);                                                 \
    template void TwoSidedTrsm(UpperOrLower, UnitOrNonUnit, AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&);   \
    template void TwoSidedTrmm(UpperOrLower, UnitOrNonUnit, AbstractDistMatrix<T>&, const AbstractDistMatrix<T>&);   \
    template void Trsv(UpperOrLower, Orientation, UnitOrNonUnit, const AbstractDistMatrix<T>&,                       \
                       AbstractDistMatrix<T>&);                                                                      \
    template void LocalTrsm(LeftOrRight, UpperOrLower, Orientation, UnitOrNonUnit, T, const AbstractDistMatrix<T>&,  \
                            AbstractDistMatrix<T>&, bool);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
