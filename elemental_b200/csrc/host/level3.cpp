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
    // output takes the row distribution of op(A)'s rows and column distribution of op(B)'s columns
    const Int m = (oA == NORMAL) ? A.Height() : A.Width();
    const Int n = (oB == NORMAL) ? B.Width() : B.Height();
    if (oA == NORMAL) C.AlignColsWith(A, true, true); else C.AlignColsWith(A, true, true);
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
    const int reserve = dev::PanelSms(8);
    const int gemmSms = std::max(1, elb200::sm_count() - reserve);
    AbstractDistMatrix<T> A1[2] = {AbstractDistMatrix<T>(g, MC, STAR), AbstractDistMatrix<T>(g, MC, STAR)};
    AbstractDistMatrix<T> B1[2] = {AbstractDistMatrix<T>(g, STAR, MR), AbstractDistMatrix<T>(g, STAR, MR)};
    dev::Event ready[2], freed[2], fork;
    for (int s = 0; s < 2; ++s) { A1[s].AlignWith(C); B1[s].AlignWith(C); }
    fork.Record(mainS);   // A, B (and the scaled C) are final on the main stream from here on
    fork.Wait(panelS);
    Int it = 0;
    for (Int k = 0; k < sumDim; k += bsize, ++it) {
        const Int nb = std::min(bsize, sumDim - k);
        const int slot = (int)(it & 1);
        {
            dev::StreamScope onPanel(panelS);
            if (it >= 2) freed[slot].Wait(panelS);
            FormPanel(oA, A, 0, k, m, nb, A1[slot]);
            FormPanel(oB, B, k, 0, nb, n, B1[slot]);
            ready[slot].Record(panelS);
        }
        ready[slot].Wait(mainS);
        {
            dev::SmLimitScope lim(gemmSms);
            LocalGemm(NORMAL, NORMAL, alpha, A1[slot], B1[slot], T(1), C);
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
    for (Int k = 0; k < n; k += bsize) {
        const Int nb = std::min(bsize, n - k);
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
    for (Int k = 0; k < m; k += bsize) {
        const Int nb = std::min(bsize, m - k);
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
template <typename T>
void Syrk(UpperOrLower uplo, Orientation o, T alpha, const AbstractDistMatrix<T>& A, T beta, AbstractDistMatrix<T>& C,
          bool conjugate) {
    const Orientation other = conjugate ? ADJOINT : TRANSPOSE;
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

// C_tri := alpha op(A) op(B)' + alphaSec op(B) op(A)' + beta C_tri, alphaSec = conj(alpha) for Her2k
// (Syr2k/LN.hpp:25).  The reference fuses both products in LocalTrr2k; two masked rank-k updates do
// the same arithmetic with one more pass over the triangle of C.
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
    if (normal) {
        Trrk(uplo, NORMAL, other, alpha, A, B, beta, C);
        Trrk(uplo, NORMAL, other, alphaSec, B, A, T(1), C);
    } else {
        Trrk(uplo, other, NORMAL, alpha, A, B, beta, C);
        Trrk(uplo, other, NORMAL, alphaSec, B, A, T(1), C);
    }
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

template <typename F>
void Trsm(LeftOrRight side, UpperOrLower uplo, Orientation o, UnitOrNonUnit diag, F alpha,
          const AbstractDistMatrix<F>& APre, AbstractDistMatrix<F>& BPre, bool checkIfSingular, TrsmAlgorithm) {
    AssertSameGrid(APre, BPre);
    if (APre.Height() != APre.Width()) LogicError("A must be square");
    if ((side == LEFT ? BPre.Height() : BPre.Width()) != APre.Height()) LogicError("Nonconformal Trsm");
    Scale(alpha, BPre);  // Trsm.cpp:94
    ReadProxy<F> AP(APre);
    ReadWriteProxy<F> XP(BPre);
    const auto& L = AP.Get();
    auto& X = XP.Get();
    const Grid& g = X.Grid();
    const Int mTri = L.Height();
    const Int bsize = Blocksize();
    const bool effLower = (uplo == LOWER) == (o == NORMAL);
    const bool forward = (side == LEFT) ? effLower : !effLower;
    AbstractDistMatrix<F> L11(g, STAR, STAR);
    const Int nblk = (mTri + bsize - 1) / bsize;
    for (Int step = 0; step < nblk; ++step) {
        const Int kb = forward ? step : nblk - 1 - step;
        const Int k = kb * bsize;
        const Int nb = std::min(bsize, mTri - k);
        const Int r0 = forward ? k + nb : 0;            // the not-yet-solved part
        const Int rl = forward ? mTri - (k + nb) : k;
        {
            auto L11v = LockedView(L, k, k, nb, nb);
            Copy(static_cast<const AbstractDistMatrix<F>&>(L11v), L11);  // L11[*,*] <- L11[MC,MR]
        }
        if (side == LEFT) {
            // Trsm/LLN.hpp:18-70 (forward) / LLT.hpp:20-80 (backward) and the LUN/LUT mirrors
            auto X1 = View(X, k, 0, nb, X.Width());
            AbstractDistMatrix<F> X1_STAR_VR(g, STAR, VR), X1_STAR_MR(g, STAR, MR);
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1), X1_STAR_VR);
            LocalTrsm(LEFT, uplo, o, diag, F(1), L11, X1_STAR_VR, checkIfSingular);
            if (rl > 0) {
                auto X2 = View(X, r0, 0, rl, X.Width());
                X1_STAR_MR.AlignWith(X2);
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_STAR_VR), X1_STAR_MR);
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_STAR_MR), X1);
                if (o == NORMAL) {
                    AbstractDistMatrix<F> Lp(g, MC, STAR);
                    Lp.AlignWith(X2);
                    auto Lv = LockedView(L, r0, k, rl, nb);
                    Copy(static_cast<const AbstractDistMatrix<F>&>(Lv), Lp);
                    LocalGemm(NORMAL, NORMAL, F(-1), Lp, X1_STAR_MR, F(1), X2);
                } else {
                    AbstractDistMatrix<F> Lp(g, STAR, MC);
                    Lp.AlignWith(X2);
                    auto Lv = LockedView(L, k, r0, nb, rl);
                    Copy(static_cast<const AbstractDistMatrix<F>&>(Lv), Lp);
                    LocalGemm(o, NORMAL, F(-1), Lp, X1_STAR_MR, F(1), X2);
                }
            } else {
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_STAR_VR), X1);
            }
        } else {
            // Trsm/RLN.hpp, RLT.hpp, RUN.hpp, RUT.hpp
            auto X1 = View(X, 0, k, X.Height(), nb);
            AbstractDistMatrix<F> X1_VC_STAR(g, VC, STAR), X1_MC_STAR(g, MC, STAR);
            Copy(static_cast<const AbstractDistMatrix<F>&>(X1), X1_VC_STAR);
            LocalTrsm(RIGHT, uplo, o, diag, F(1), L11, X1_VC_STAR, checkIfSingular);
            if (rl > 0) {
                auto X2 = View(X, 0, r0, X.Height(), rl);
                X1_MC_STAR.AlignWith(X2);
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_VC_STAR), X1_MC_STAR);
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_MC_STAR), X1);
                if (o == NORMAL) {
                    AbstractDistMatrix<F> Lp(g, STAR, MR);
                    Lp.AlignWith(X2);
                    auto Lv = LockedView(L, k, r0, nb, rl);
                    Copy(static_cast<const AbstractDistMatrix<F>&>(Lv), Lp);
                    LocalGemm(NORMAL, NORMAL, F(-1), X1_MC_STAR, Lp, F(1), X2);
                } else {
                    AbstractDistMatrix<F> Lp(g, MR, STAR);
                    Lp.AlignWith(X2);
                    auto Lv = LockedView(L, r0, k, rl, nb);
                    Copy(static_cast<const AbstractDistMatrix<F>&>(Lv), Lp);
                    LocalGemm(NORMAL, o, F(-1), X1_MC_STAR, Lp, F(1), X2);
                }
            } else {
                Copy(static_cast<const AbstractDistMatrix<F>&>(X1_VC_STAR), X1);
            }
        }
    }
    XP.Commit();
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
    template void LocalTrsm(LeftOrRight, UpperOrLower, Orientation, UnitOrNonUnit, T, const AbstractDistMatrix<T>&,  \
                            AbstractDistMatrix<T>&, bool);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
