// Right-looking blocked Cholesky (lower and upper), SolveAfter and HPDSolve on
// DistMatrix<F,MC,MR>.  Follows the reference's variant-3 panel loop
//   src/lapack_like/factor/Cholesky/LowerVariant3.hpp:70-126, UpperVariant3.hpp:75-123,
//   Cholesky/SolveAfter.hpp:78-107, src/lapack_like/solve/HPD.cpp:27-69
// with these B200-first changes:
//   * the replicated diagonal block is factored by ONE single-CTA kernel
//     (potrf.cu) instead of nb level-2 sweeps; a pivot failure is recorded in a device
//     flag and raised as NonHPDMatrixException once, after the sweep, on every rank
//     (the block is replicated, so every rank sees the same flag: no deadlock);
//   * A21[VC,*] is turned into the [MC,*] and [MR,*] panels of the update by two direct
//     redistributions (no [VR,*] exchange, no explicit transposes): the masked GEMM
//     takes A21[MC,*] and (A21[MR,*])^H as they are;
//   * the trailing update is one masked tensor-pipe GEMM (LocalTrrk).
#include <algorithm>

#include "dev.hpp"
#include "elb200/factor.hpp"

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

struct InfoFlag {
    int* dev_ = nullptr;
    InfoFlag() {
        dev_ = (int*)elb200::scratch_alloc(sizeof(int), dev::stream());
        ELB_CUDA(cudaMemsetAsync(dev_, 0, sizeof(int), dev::stream()));
    }
    ~InfoFlag() { if (dev_) cudaFreeAsync(dev_, dev::stream()); }
    int Read() {
        int h = 0;
        ELB_CUDA(cudaMemcpyAsync(&h, dev_, sizeof(int), cudaMemcpyDeviceToHost, dev::stream()));
        ELB_CUDA(cudaStreamSynchronize(dev::stream()));
        return h;
    }
};

template <typename F>
void LocalPotrf(UpperOrLower uplo, Matrix<F>& A, int* info, Int colOffset) {
    if (A.Height() != A.Width()) LogicError("Can only compute Cholesky factor of square matrices");
    elb200::potrf_device<dev::D<F>>(UpperOrLowerToChar(uplo), A.Height(), dev::ptr(A.Buffer()), A.LDim(), info,
                                    colOffset, dev::stream());
}

template <typename F>
void LowerVariant3Blocked(AbstractDistMatrix<F>& A, InfoFlag& info) {
    const Grid& g = A.Grid();
    const Int n = A.Height();
    const Int bsize = Blocksize();
    AbstractDistMatrix<F> A11_STAR_STAR(g, STAR, STAR);
    for (Int k = 0; k < n; k += bsize) {
        const Int nb = std::min(bsize, n - k);
        const Int m2 = n - (k + nb);
        auto A11 = View(A, k, k, nb, nb);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A11), A11_STAR_STAR);
        LocalPotrf(LOWER, A11_STAR_STAR.Matrix(), info.dev_, k);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A11_STAR_STAR), A11);
        if (m2 <= 0) break;
        auto A21 = View(A, k + nb, k, m2, nb);
        auto A22 = View(A, k + nb, k + nb, m2, m2);
        AbstractDistMatrix<F> A21_VC_STAR(g, VC, STAR), A21_MC_STAR(g, MC, STAR), A21_MR_STAR(g, MR, STAR);
        A21_VC_STAR.AlignWith(A22);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A21), A21_VC_STAR);
        LocalTrsm(RIGHT, LOWER, ADJOINT, NON_UNIT, F(1), A11_STAR_STAR, A21_VC_STAR);
        A21_MC_STAR.AlignWith(A22);
        A21_MR_STAR.AlignWith(A22);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A21_VC_STAR), A21_MC_STAR);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A21_VC_STAR), A21_MR_STAR);
        // A22 -= A21 A21^H on the lower staircase
        LocalTrrk(LOWER, NORMAL, ADJOINT, F(-1), A21_MC_STAR, A21_MR_STAR, F(1), A22);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A21_MC_STAR), A21);
    }
}

template <typename F>
void UpperVariant3Blocked(AbstractDistMatrix<F>& A, InfoFlag& info) {
    const Grid& g = A.Grid();
    const Int n = A.Height();
    const Int bsize = Blocksize();
    AbstractDistMatrix<F> A11_STAR_STAR(g, STAR, STAR);
    for (Int k = 0; k < n; k += bsize) {
        const Int nb = std::min(bsize, n - k);
        const Int m2 = n - (k + nb);
        auto A11 = View(A, k, k, nb, nb);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A11), A11_STAR_STAR);
        LocalPotrf(UPPER, A11_STAR_STAR.Matrix(), info.dev_, k);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A11_STAR_STAR), A11);
        if (m2 <= 0) break;
        auto A12 = View(A, k, k + nb, nb, m2);
        auto A22 = View(A, k + nb, k + nb, m2, m2);
        AbstractDistMatrix<F> A12_STAR_VR(g, STAR, VR), A12_STAR_MC(g, STAR, MC), A12_STAR_MR(g, STAR, MR);
        A12_STAR_VR.AlignWith(A22);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A12), A12_STAR_VR);
        LocalTrsm(LEFT, UPPER, ADJOINT, NON_UNIT, F(1), A11_STAR_STAR, A12_STAR_VR);
        A12_STAR_MC.AlignWith(A22);
        A12_STAR_MR.AlignWith(A22);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A12_STAR_VR), A12_STAR_MC);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A12_STAR_VR), A12_STAR_MR);
        // A22 -= A12^H A12 on the upper staircase
        LocalTrrk(UPPER, ADJOINT, NORMAL, F(-1), A12_STAR_MC, A12_STAR_MR, F(1), A22);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A12_STAR_MR), A12);
    }
}

}  // namespace

template <typename F>
void Cholesky(UpperOrLower uplo, Matrix<F>& A) {
    InfoFlag info;
    LocalPotrf(uplo, A, info.dev_, 0);
    if (info.Read() != 0) throw NonHPDMatrixException("A was not numerically HPD");
}

template <typename F>
void Cholesky(UpperOrLower uplo, AbstractDistMatrix<F>& APre, bool scalapack) {
    if (scalapack) LogicError("The ScaLAPACK path is out of scope of this build (EL_DISABLE_SCALAPACK)");
    if (APre.Height() != APre.Width()) LogicError("Can only compute Cholesky factor of square matrices");
    InfoFlag info;
    if (APre.ColDist() == STAR && APre.RowDist() == STAR) {
        // Cholesky(uplo, DistMatrix<F,STAR,STAR>&): redundant local factorisation (Cholesky.cpp:123-126)
        LocalPotrf(uplo, APre.Matrix(), info.dev_, 0);
    } else if (APre.ColDist() == MC && APre.RowDist() == MR) {
        if (uplo == LOWER) LowerVariant3Blocked(APre, info);
        else UpperVariant3Blocked(APre, info);
    } else {
        AbstractDistMatrix<F> A(APre.Grid(), MC, MR);
        Copy(static_cast<const AbstractDistMatrix<F>&>(APre), A);
        if (uplo == LOWER) LowerVariant3Blocked(A, info);
        else UpperVariant3Blocked(A, info);
        Copy(static_cast<const AbstractDistMatrix<F>&>(A), APre);
    }
    if (info.Read() != 0) throw NonHPDMatrixException("A was not numerically HPD");
}

namespace cholesky {
template <typename F>
void SolveAfter(UpperOrLower uplo, Orientation o, const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B) {
    if (A.Height() != A.Width()) LogicError("A must be square");
    if (A.Height() != B.Height()) LogicError("A and B must be the same height");
    if (o == TRANSPOSE) Conjugate(B);
    if (uplo == LOWER) {
        Trsm(LEFT, LOWER, NORMAL, NON_UNIT, F(1), A, B);
        Trsm(LEFT, LOWER, ADJOINT, NON_UNIT, F(1), A, B);
    } else {
        Trsm(LEFT, UPPER, ADJOINT, NON_UNIT, F(1), A, B);
        Trsm(LEFT, UPPER, NORMAL, NON_UNIT, F(1), A, B);
    }
    if (o == TRANSPOSE) Conjugate(B);
}
template <typename F>
void SolveAfter(UpperOrLower uplo, Orientation o, const Matrix<F>& A, Matrix<F>& B) {
    if (A.Height() != A.Width()) LogicError("A must be square");
    if (A.Height() != B.Height()) LogicError("A and B must be the same height");
    auto conj = [&]() {
        if (!IsComplex<F>::value || B.Height() == 0 || B.Width() == 0) return;
        elb200::lattice_copy_device<dev::D<F>>(dev::ptr(B.LockedBuffer()), dev::ptr(B.Buffer()), B.Height(), B.Width(), 0,
                                               1, B.LDim(), 0, 1, B.LDim(), true, nullptr, false, dev::stream());
    };
    if (o == TRANSPOSE) conj();
    if (uplo == LOWER) {
        Trsm(LEFT, LOWER, NORMAL, NON_UNIT, F(1), A, B);
        Trsm(LEFT, LOWER, ADJOINT, NON_UNIT, F(1), A, B);
    } else {
        Trsm(LEFT, UPPER, ADJOINT, NON_UNIT, F(1), A, B);
        Trsm(LEFT, UPPER, NORMAL, NON_UNIT, F(1), A, B);
    }
    if (o == TRANSPOSE) conj();
}
}  // namespace cholesky

namespace hpd_solve {
template <typename F>
void Overwrite(UpperOrLower uplo, Orientation o, AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B) {
    Cholesky(uplo, A);
    cholesky::SolveAfter(uplo, o, static_cast<const AbstractDistMatrix<F>&>(A), B);
}
template <typename F>
void Overwrite(UpperOrLower uplo, Orientation o, Matrix<F>& A, Matrix<F>& B) {
    Cholesky(uplo, A);
    cholesky::SolveAfter(uplo, o, static_cast<const Matrix<F>&>(A), B);
}
}  // namespace hpd_solve

template <typename F>
void HPDSolve(UpperOrLower uplo, Orientation o, const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B) {
    // HPD.cpp:59-69: factor a copy of A ([MC,MR]), then two triangular solves
    AbstractDistMatrix<F> ACopy(A.Grid(), MC, MR);
    Copy(A, ACopy);
    hpd_solve::Overwrite(uplo, o, ACopy, B);
}
template <typename F>
void HPDSolve(UpperOrLower uplo, Orientation o, const Matrix<F>& A, Matrix<F>& B) {
    Matrix<F> ACopy(A);
    hpd_solve::Overwrite(uplo, o, ACopy, B);
}

#define ELB_INST(T)                                                                                              \
    template void Cholesky(UpperOrLower, Matrix<T>&);                                                            \
    template void Cholesky(UpperOrLower, AbstractDistMatrix<T>&, bool);                                          \
    template void cholesky::SolveAfter(UpperOrLower, Orientation, const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&); \
    template void cholesky::SolveAfter(UpperOrLower, Orientation, const Matrix<T>&, Matrix<T>&);                 \
    template void hpd_solve::Overwrite(UpperOrLower, Orientation, AbstractDistMatrix<T>&, AbstractDistMatrix<T>&); \
    template void hpd_solve::Overwrite(UpperOrLower, Orientation, Matrix<T>&, Matrix<T>&);                       \
    template void HPDSolve(UpperOrLower, Orientation, const AbstractDistMatrix<T>&, AbstractDistMatrix<T>&);     \
    template void HPDSolve(UpperOrLower, Orientation, const Matrix<T>&, Matrix<T>&);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
