// elb200 host layer, part 3: Cholesky and HPDSolve with the reference's signatures
// (include/El/lapack_like/factor.hpp:21-27,60-90; include/El/lapack_like/solve.hpp:189-228).
#pragma once
#include "elb200/level3.hpp"

namespace El {

// In-place Cholesky factor of the uplo triangle; the other triangle is left untouched.
// Throws NonHPDMatrixException when a pivot is <= 0 (LowerVariant3.hpp:29-30).
template <typename F> void Cholesky(UpperOrLower uplo, Matrix<F>& A);
template <typename F> void Cholesky(UpperOrLower uplo, AbstractDistMatrix<F>& A, bool scalapack = false);

// Reverse factorisations A = L^H L (LOWER) / A = U U^H (UPPER), src/lapack_like/factor/Cholesky.cpp:55-141
// (cholesky::ReverseLowerVariant3Blocked / ReverseUpperVariant3Blocked)
template <typename F> void ReverseCholesky(UpperOrLower uplo, Matrix<F>& A);
template <typename F> void ReverseCholesky(UpperOrLower uplo, AbstractDistMatrix<F>& A);

namespace cholesky {
// the left-looking blocked variants (Cholesky/LowerVariant2.hpp:43-110, UpperVariant2.hpp); same factor as Cholesky()
template <typename F> void LowerVariant2Blocked(AbstractDistMatrix<F>& A);
template <typename F> void UpperVariant2Blocked(AbstractDistMatrix<F>& A);
// A holds a Cholesky factor; B := inv(A) B (or the transposed system), SolveAfter.hpp:78-107
template <typename F>
void SolveAfter(UpperOrLower uplo, Orientation orientation, const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B);
template <typename F>
void SolveAfter(UpperOrLower uplo, Orientation orientation, const Matrix<F>& A, Matrix<F>& B);
}  // namespace cholesky

// B := inv(A) B for Hermitian positive-definite A (src/lapack_like/solve/HPD.cpp:47-69)
template <typename F>
void HPDSolve(UpperOrLower uplo, Orientation orientation, const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B);
template <typename F>
void HPDSolve(UpperOrLower uplo, Orientation orientation, const Matrix<F>& A, Matrix<F>& B);
namespace hpd_solve {
// overwrites A with its factor (HPD.cpp:27-43)
template <typename F>
void Overwrite(UpperOrLower uplo, Orientation orientation, AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B);
template <typename F>
void Overwrite(UpperOrLower uplo, Orientation orientation, Matrix<F>& A, Matrix<F>& B);
}  // namespace hpd_solve

}  // namespace El
