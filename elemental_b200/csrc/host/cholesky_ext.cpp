// The rest of the reference's Cholesky family on device-resident matrices:
//   CholeskyMod          src/lapack_like/factor/Cholesky.cpp:143-173, Cholesky/LowerMod.hpp, UpperMod.hpp
//   Cholesky(uplo, A, P) Cholesky.cpp:40-53,112-121, Cholesky/PivotedLowerVariant3.hpp, PivotedUpperVariant3.hpp
//   cholesky::SolveAfter(uplo, o, A, P, B)   Cholesky/SolveAfter.hpp:108-141
#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "dev.hpp"
#include "elb200/lu.hpp"

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
template <typename T>
const AbstractDistMatrix<T>& C(const AbstractDistMatrix<T>& A) { return A; }

template <typename T>
AbstractDistMatrix<T> AsMcMr(Matrix<T>& A) {
    AbstractDistMatrix<T> D(El::Grid::Default(), MC, MR);
    D.Attach(A.Height(), A.Width(), El::Grid::Default(), 0, 0, A.Buffer(), A.LDim());
    return D;
}

// T [MC,MR] holds the factor in its uplo triangle; V [MC,MR] is n x w.  The panel of the factor and all of V are
// replicated; the kernel works on column panels of L = T (LOWER) or L = T^H (UPPER: U'^H U' = U^H U + alpha V V^H is
// the same statement about L = U^H, and the factor with a positive diagonal is unique).
template <typename F>
void CholeskyModMcMr(UpperOrLower uplo, AbstractDistMatrix<F>& T, Base<F> alpha, AbstractDistMatrix<F>& V) {
    typedef dev::D<F> D;
    typedef Base<F> Real;
    const El::Grid& g = T.Grid();
    const Int n = T.Height(), w = V.Width();
    if (T.Width() != n) LogicError("Cholesky factors must be square");
    if (V.Height() != n) LogicError("V is the wrong height");
    if (alpha == Real(0) || n == 0) return;
    cudaStream_t s = dev::stream();
    const bool down = alpha < Real(0);
    Scale(F(std::sqrt(down ? -alpha : alpha)), V);   // LowerMod.hpp:246-256
    AbstractDistMatrix<F> Vs(g, STAR, STAR), panel(g, STAR, STAR);
    Copy(C(V), Vs);
    dev::DeviceFlag info;
    const Int bsize = Blocksize();
    for (Int k = 0; k < n; k += bsize) {
        const Int nb = std::min(bsize, n - k);
        if (uplo == LOWER) {
            auto TB1 = View(T, k, k, n - k, nb);
            Copy(C(TB1), panel);
            elb200::cholmod_panel_device<D>(down, n - k, nb, dev::ptr(panel.Buffer()), panel.LDim(),
                                            dev::ptr(Vs.Buffer()) + k, Vs.LDim(), w, info.dev_, s);
            Copy(C(panel), TB1);
        } else {
            auto T1R = View(T, k, k, nb, n - k);
            Adjoint(C(T1R), panel);
            elb200::cholmod_panel_device<D>(down, n - k, nb, dev::ptr(panel.Buffer()), panel.LDim(),
                                            dev::ptr(Vs.Buffer()) + k, Vs.LDim(), w, info.dev_, s);
            Adjoint(C(panel), T1R);
        }
    }
    Copy(C(Vs), V);   // V is workspace on return, as in the reference
    if (info.Read() != 0) LogicError("Attempted to square-root a negative number");   // Hyperbolic/Row.hpp:39-40
}

// ---------------------------------------------------------------------------------------------------------------
// Diagonally pivoted Cholesky: P A P^T = L L^H (LOWER) / U^H U (UPPER)
// ---------------------------------------------------------------------------------------------------------------
// The reference's lazy panel (PivotedLowerVariant3.hpp:230-290) swaps the stored triangle symmetrically and
// broadcasts along rows and columns for EVERY column.  B200-first restatement of the same algorithm:
//   * the matrix is kept as a full Hermitian matrix W for the duration (HBM holds it twice), so a symmetric
//     interchange is a row interchange plus a column interchange and a pivot column is one gather W(:, c) -> [*,*];
//   * the whole state of a panel -- its computed columns X, the running diagonal d, the position map -- is
//     replicated, so pivot search, the lazy column update, the scaling and the diagonal update are local kernels;
//   * physical interchanges are applied once per PANEL (the slot machinery of LU), not once per column;
//   * the trailing update is one GEMM W22 -= X21 X21^H on both triangles (twice the flops of the masked update:
//     the price of the cheap interchanges).
// One host round trip per column remains: the pivot index selects which column the redistribution engine gathers.
// The pivot rule is the algorithm's: the largest remaining diagonal entry, first occurrence (VectorMaxAbsLoc).
template <typename F>
struct SlotInterchange {
    typedef dev::D<F> D;
    typedef long long i64;
    i64* slotIdx = nullptr;
    int* srcSlot = nullptr;
    explicit SlotInterchange(Int maxNb) {
        slotIdx = (i64*)elb200::scratch_alloc((sizeof(i64) + sizeof(int)) * 2 * (size_t)maxNb, dev::stream());
        srcSlot = (int*)(slotIdx + 2 * maxNb);
    }
    ~SlotInterchange() { if (slotIdx) cudaFreeAsync(slotIdx, dev::stream()); }
    // indices k + j <-> k + ipiv[j], j = 0 .. nb-1, applied to the rows AND the columns of W
    void Symmetric(AbstractDistMatrix<F>& W, Int k, Int nb, const i64* ipiv) {
        cudaStream_t s = dev::stream();
        const El::Grid& g = W.Grid();
        const int S = 2 * (int)nb;
        elb200::swap_plan_device((int)nb, ipiv, k, slotIdx, srcSlot, s);
        const i64 mloc = W.LocalHeight(), nloc = W.LocalWidth();
        if (mloc == 0 || nloc == 0) return;
        const int r = W.ColStride(), c = W.RowStride();
        const size_t perR = (size_t)S * (size_t)nloc, perC = (size_t)S * (size_t)mloc;
        const size_t need = std::max(perR * (size_t)(r > 1 ? r + 1 : 1), perC * (size_t)(c > 1 ? c + 1 : 1));
        D* buf = (D*)elb200::scratch_alloc(sizeof(D) * need, s);
        // rows
        elb200::pack_rows_device<D>(S, slotIdx, srcSlot, dev::ptr(W.LockedBuffer()), W.LDim(), nloc, W.ColAlign(), r, W.ColRank(),
                                    W.ColShift(), buf, s);
        const D* all = buf;
        if (r > 1) {
            ELB_NCCL(ncclAllGather(buf, buf + perR, perR * sizeof(D), ncclInt8, (ncclComm_t)g.MCComm().nccl, s));
            GetRedistStats().allGathers++;
            all = buf + perR;
        }
        elb200::unpack_rows_device<D>(S, slotIdx, srcSlot, dev::ptr(W.Buffer()), W.LDim(), nloc, W.ColAlign(), r, W.ColRank(),
                                      W.ColShift(), all, (i64)perR, s);
        // columns
        elb200::pack_cols_device<D>(S, slotIdx, srcSlot, dev::ptr(W.LockedBuffer()), W.LDim(), mloc, W.RowAlign(), c, W.RowRank(),
                                    W.RowShift(), buf, s);
        all = buf;
        if (c > 1) {
            ELB_NCCL(ncclAllGather(buf, buf + perC, perC * sizeof(D), ncclInt8, (ncclComm_t)g.MRComm().nccl, s));
            GetRedistStats().allGathers++;
            all = buf + perC;
        }
        elb200::unpack_cols_device<D>(S, slotIdx, srcSlot, dev::ptr(W.Buffer()), W.LDim(), mloc, W.RowAlign(), c, W.RowRank(),
                                      W.RowShift(), all, (i64)perC, s);
        elb200::scratch_free(buf, s);
    }
};

template <typename F>
void PivotedCholeskyMcMr(UpperOrLower uplo, AbstractDistMatrix<F>& A, DistPermutation& P) {
    typedef dev::D<F> D;
    typedef long long i64;
    const El::Grid& g = A.Grid();
    const Int n = A.Height();
    if (A.Width() != n) LogicError("A must be square");
    P.SetGrid(g);
    P.MakeIdentity(n);
    P.ReserveSwaps(n);
    if (n == 0) return;
    const Int bsize = Blocksize();
    if (bsize > 512) LogicError("Pivoted Cholesky: Blocksize() above 512 is not supported");
    cudaStream_t s = dev::stream();
    const UpperOrLower other = (uplo == LOWER) ? UPPER : LOWER;

    // W := the Hermitian matrix the uplo triangle of A stands for
    AbstractDistMatrix<F> W(g, MC, MR), Wh(g, MC, MR);
    W.Align(A.ColAlign(), A.RowAlign());
    Wh.Align(A.ColAlign(), A.RowAlign());
    Copy(C(A), W);
    MakeTrapezoidal(uplo, W);
    Adjoint(C(W), Wh);
    AxpyTrapezoid(other, F(1), C(Wh), W, uplo == LOWER ? 1 : -1);
    Wh.Empty();

    // replicated running diagonal
    double* d = (double*)elb200::scratch_alloc(sizeof(double) * (size_t)n, s);
    ELB_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)n, s));
    elb200::diag_extract_device<D>(W.LocalHeight(), W.LocalWidth(), dev::ptr(W.LockedBuffer()), W.LDim(), W.ColShift(), W.ColStride(),
                                   W.RowShift(), W.RowStride(), d, s);
    if (g.Size() > 1) ELB_NCCL(ncclAllReduce(d, d, (size_t)n, ncclDouble, ncclSum, (ncclComm_t)g.VCComm().nccl, s));

    i64* pos = (i64*)elb200::scratch_alloc(sizeof(i64) * ((size_t)n + (size_t)bsize + 1), s);
    i64* ipiv = pos + n;
    i64* pivDev = ipiv + bsize;
    std::vector<i64> ident(n), posHost(n);
    for (Int i = 0; i < n; ++i) ident[i] = i;
    AbstractDistMatrix<F> X(g, STAR, STAR), h(g, STAR, STAR), X21_MC(g, MC, STAR), X21_MR(g, MR, STAR);
    SlotInterchange<F> swapper(bsize);
    dev::DeviceFlag info;

    for (Int off = 0; off < n; off += bsize) {
        const Int nb = std::min(bsize, n - off), M = n - off;
        X.Resize(M, nb);
        ELB_CUDA(cudaMemcpyAsync(pos, ident.data(), sizeof(i64) * (size_t)M, cudaMemcpyHostToDevice, s));
        std::copy(ident.begin(), ident.begin() + M, posHost.begin());
        for (Int k = 0; k < nb; ++k) {
            // pivot: the largest remaining diagonal entry (PanelFull / VectorMaxAbsLoc); its index picks the column
            elb200::argmax_abs_device(d, off + k, n, pivDev, s);
            i64 piv = 0;
            ELB_CUDA(cudaMemcpyAsync(&piv, pivDev, sizeof(i64), cudaMemcpyDeviceToHost, s));
            ELB_CUDA(cudaStreamSynchronize(s));
            const i64 f = piv - off;
            elb200::pivot_swap_device<D>((int)k, f, d + off, dev::ptr(X.Buffer()), X.LDim(), pos, ipiv, s);
            std::swap(posHost[k], posHost[f]);
            // the pivot column of the matrix as it stood at the start of the panel, replicated
            auto col = LockedView(C(W), off, off + (Int)posHost[k], M, 1);
            Copy(C(col), h);
            elb200::pivot_column_device<D>(M, (int)k, dev::ptr(h.LockedBuffer()), pos, dev::ptr(X.Buffer()), X.LDim(), d + off,
                                           info.dev_, off + k, s);
        }
        P.AppendDeviceSwaps(ipiv, nb, off);
        // the interchanges of the panel, physically and symmetrically; then the panel's columns of the factor
        swapper.Symmetric(W, off, nb, ipiv);
        auto WB1 = View(W, off, off, M, nb);
        Copy(C(X), WB1);
        if (off + nb < n) {
            auto W22 = View(W, off + nb, off + nb, M - nb, M - nb);
            auto X21 = LockedView(C(X), nb, 0, M - nb, nb);
            X21_MC.AlignWith(W22);
            X21_MR.AlignWith(W22);
            Copy(C(X21), X21_MC);
            Copy(C(X21), X21_MR);
            LocalGemm(NORMAL, ADJOINT, F(-1), C(X21_MC), C(X21_MR), F(1), W22);
        }
    }
    elb200::scratch_free(d, s);
    elb200::scratch_free(pos, s);
    // the factor into the uplo triangle of A; the other triangle of A keeps its values
    ScaleTrapezoid(F(0), uplo, A);
    if (uplo == LOWER) {
        AxpyTrapezoid(LOWER, F(1), C(W), A);
    } else {
        AbstractDistMatrix<F> U(g, MC, MR);
        U.Align(A.ColAlign(), A.RowAlign());
        Adjoint(C(W), U);
        AxpyTrapezoid(UPPER, F(1), C(U), A);
    }
    if (info.Read() != 0) throw NonHPDMatrixException("A was not numerically HPD");
}

}  // namespace

template <typename F>
void Cholesky(UpperOrLower uplo, AbstractDistMatrix<F>& APre, DistPermutation& P) {
    if (APre.ColDist() == MC && APre.RowDist() == MR) {
        PivotedCholeskyMcMr(uplo, APre, P);
    } else {
        AbstractDistMatrix<F> A(APre.Grid(), MC, MR);
        Copy(C(APre), A);
        PivotedCholeskyMcMr(uplo, A, P);
        Copy(C(A), APre);
    }
}
template <typename F>
void Cholesky(UpperOrLower uplo, Matrix<F>& A, Permutation& P) {
    auto D = AsMcMr(A);
    PivotedCholeskyMcMr(uplo, D, P);
}
namespace cholesky {
// Cholesky/SolveAfter.hpp:108-141
template <typename F>
void SolveAfter(UpperOrLower uplo, Orientation o, const AbstractDistMatrix<F>& A, const DistPermutation& P,
                AbstractDistMatrix<F>& B) {
    if (A.Height() != A.Width()) LogicError("A must be square");
    if (A.Height() != B.Height()) LogicError("A and B must be the same height");
    P.PermuteRows(B);
    SolveAfter(uplo, o, A, B);
    P.InversePermuteRows(B);
}
}  // namespace cholesky

template <typename F>
void CholeskyMod(UpperOrLower uplo, AbstractDistMatrix<F>& TPre, Base<F> alpha, AbstractDistMatrix<F>& VPre) {
    if (&TPre.Grid() != &VPre.Grid()) LogicError("Grids must match");
    std::unique_ptr<AbstractDistMatrix<F>> Tc, Vc;
    AbstractDistMatrix<F>* T = &TPre;
    AbstractDistMatrix<F>* V = &VPre;
    if (TPre.ColDist() != MC || TPre.RowDist() != MR) { Tc.reset(new AbstractDistMatrix<F>(TPre.Grid(), MC, MR)); Copy(C(TPre), *Tc); T = Tc.get(); }
    if (VPre.ColDist() != MC || VPre.RowDist() != MR) { Vc.reset(new AbstractDistMatrix<F>(VPre.Grid(), MC, MR)); Copy(C(VPre), *Vc); V = Vc.get(); }
    CholeskyModMcMr(uplo, *T, alpha, *V);
    if (Tc) Copy(C(*Tc), TPre);
    if (Vc) Copy(C(*Vc), VPre);
}
template <typename F>
void CholeskyMod(UpperOrLower uplo, Matrix<F>& T, Base<F> alpha, Matrix<F>& V) {
    auto DT = AsMcMr(T);
    auto DV = AsMcMr(V);
    CholeskyModMcMr(uplo, DT, alpha, DV);
}

#define ELB_INST(F)                                                                              \
    template void Cholesky(UpperOrLower, AbstractDistMatrix<F>&, DistPermutation&);              \
    template void Cholesky(UpperOrLower, Matrix<F>&, Permutation&);                              \
    template void cholesky::SolveAfter(UpperOrLower, Orientation, const AbstractDistMatrix<F>&, const DistPermutation&, \
                                       AbstractDistMatrix<F>&);                                   \
    template void CholeskyMod(UpperOrLower, AbstractDistMatrix<F>&, Base<F>, AbstractDistMatrix<F>&); \
    template void CholeskyMod(UpperOrLower, Matrix<F>&, Base<F>, Matrix<F>&);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
