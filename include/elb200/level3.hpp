// elb200 host layer, part 2: the distributed level-3 routines of the hot path with the
// reference's signatures (include/El/blas_like/level3.hpp:34-68 Gemm/LocalGemm, :85-124
// Herk/Syrk, :455-477 Trsm/LocalTrsm, :551-590 Trrk/LocalTrrk).
#pragma once
#include "elb200/core.hpp"

namespace El {

// ---- Gemm: C := alpha op(A) op(B) + beta C (src/blas_like/level3/Gemm.cpp:19-133) ----
template <typename T>
void Gemm(Orientation orientA, Orientation orientB, T alpha, const Matrix<T>& A, const Matrix<T>& B, T beta,
          Matrix<T>& C);
template <typename T>
void Gemm(Orientation orientA, Orientation orientB, T alpha, const Matrix<T>& A, const Matrix<T>& B, Matrix<T>& C);
template <typename T>
void Gemm(Orientation orientA, Orientation orientB, T alpha, const AbstractDistMatrix<T>& A,
          const AbstractDistMatrix<T>& B, T beta, AbstractDistMatrix<T>& C, GemmAlgorithm alg = GEMM_DEFAULT);
template <typename T>
void Gemm(Orientation orientA, Orientation orientB, T alpha, const AbstractDistMatrix<T>& A,
          const AbstractDistMatrix<T>& B, AbstractDistMatrix<T>& C, GemmAlgorithm alg = GEMM_DEFAULT);
// local product of the LOCAL matrices of compatibly distributed operands (Gemm.cpp:135-259)
template <typename T>
void LocalGemm(Orientation orientA, Orientation orientB, T alpha, const AbstractDistMatrix<T>& A,
               const AbstractDistMatrix<T>& B, T beta, AbstractDistMatrix<T>& C);
template <typename T>
void LocalGemm(Orientation orientA, Orientation orientB, T alpha, const AbstractDistMatrix<T>& A,
               const AbstractDistMatrix<T>& B, AbstractDistMatrix<T>& C);
// which SUMMA variant GEMM_DEFAULT resolves to (Gemm/NN.hpp:304-313)
GemmAlgorithm GemmDefaultAlgorithm(Int m, Int n, Int k);

// ---- GemmHost: the same product for operands whose [MC,MR] local matrices are in HOST memory ----
// (what a caller of the reference holds: DistMatrix buffers are host allocations there).  height / width are the
// GLOBAL dimensions, buffer / ldim the column-major local matrix of this process (alignments 0, i.e. what
// DistMatrix<T>(grid) gives).  C is streamed through HBM in column bands, A in chunks of the summation index, host
// copies overlapped with the SUMMA updates (csrc/host/stream_gemm.cpp); pinned (page-locked) buffers are needed for
// the overlap, pageable ones still give the right result.  On return C is complete in host memory.
template <typename T>
struct HostLocalMatrix {
    Int height, width;
    T* buffer;
    Int ldim;
};
struct GemmHostStats { int bands, chunks; Int bandWidth, chunkWidth; };
template <typename T>
void GemmHost(Orientation orientA, Orientation orientB, T alpha, const Grid& grid, const HostLocalMatrix<T>& A,
              const HostLocalMatrix<T>& B, T beta, const HostLocalMatrix<T>& C, GemmAlgorithm alg = GEMM_DEFAULT,
              GemmHostStats* stats = nullptr);

// ---- Trrk: triangular rank-k update (src/blas_like/level3/Trrk.cpp, Trrk/Local.hpp) ----
template <typename T>
void Trrk(UpperOrLower uplo, Orientation orientA, Orientation orientB, T alpha, const Matrix<T>& A,
          const Matrix<T>& B, T beta, Matrix<T>& C);
template <typename T>
void Trrk(UpperOrLower uplo, Orientation orientA, Orientation orientB, T alpha, const AbstractDistMatrix<T>& A,
          const AbstractDistMatrix<T>& B, T beta, AbstractDistMatrix<T>& C);
// C[MC,MR] triangle += alpha op(A_local) op(B_local) with the global staircase mask
template <typename T>
void LocalTrrk(UpperOrLower uplo, Orientation orientA, Orientation orientB, T alpha,
               const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B, T beta, AbstractDistMatrix<T>& C);

// ---- Herk / Syrk (src/blas_like/level3/Herk.cpp:14-57, Syrk.cpp:20-86) ----
template <typename T>
void Syrk(UpperOrLower uplo, Orientation orientation, T alpha, const Matrix<T>& A, T beta, Matrix<T>& C,
          bool conjugate = false);
template <typename T>
void Syrk(UpperOrLower uplo, Orientation orientation, T alpha, const AbstractDistMatrix<T>& A, T beta,
          AbstractDistMatrix<T>& C, bool conjugate = false);
template <typename T>
void Herk(UpperOrLower uplo, Orientation orientation, Base<T> alpha, const Matrix<T>& A, Base<T> beta, Matrix<T>& C);
template <typename T>
void Herk(UpperOrLower uplo, Orientation orientation, Base<T> alpha, const AbstractDistMatrix<T>& A, Base<T> beta,
          AbstractDistMatrix<T>& C);

// ---- Syr2k / Her2k, Symm / Hemm, Trmm (siblings over the same leaves) ----
// Reference: src/blas_like/level3/Syr2k.cpp + Syr2k/{LN,LT,UN,UT}.hpp (LocalTrr2k), Symm.cpp + Symm/*.hpp,
// Trmm.cpp + Trmm/*.hpp.  Here: Syr2k = Trr2k (one masked GEMM over the stacked panels per step); Symm = Gemm on the mirrored
// triangle (an n x n temporary: sized for HBM, not for a CPU cache); Trmm = Gemm on the trapezoid copy.
template <typename T>
void Syr2k(UpperOrLower uplo, Orientation orientation, T alpha, const AbstractDistMatrix<T>& A,
           const AbstractDistMatrix<T>& B, T beta, AbstractDistMatrix<T>& C, bool conjugate = false);
// E_tri := alpha op(A) op(B) + beta op(C) op(D) + gamma E_tri (src/blas_like/level3/Trr2k.cpp:34-...,
// include/El/blas_like/level3.hpp:592-640); LocalTrr2k takes the panels already in [MC,*] / [*,MR] (or their
// transposes) and runs ONE masked GEMM over the stacked summation index (Trr2k/Local.hpp does two products per leaf)
template <typename T>
void Trr2k(UpperOrLower uplo, Orientation orientA, Orientation orientB, Orientation orientC, Orientation orientD,
           T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B, T beta,
           const AbstractDistMatrix<T>& C, const AbstractDistMatrix<T>& D, T gamma, AbstractDistMatrix<T>& E);
template <typename T>
void LocalTrr2k(UpperOrLower uplo, Orientation orientA, Orientation orientB, Orientation orientC,
                Orientation orientD, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B, T beta,
                const AbstractDistMatrix<T>& C, const AbstractDistMatrix<T>& D, T gamma, AbstractDistMatrix<T>& E);
template <typename T>
void Her2k(UpperOrLower uplo, Orientation orientation, T alpha, const AbstractDistMatrix<T>& A,
           const AbstractDistMatrix<T>& B, Base<T> beta, AbstractDistMatrix<T>& C);
template <typename T>
void Symm(LeftOrRight side, UpperOrLower uplo, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
          T beta, AbstractDistMatrix<T>& C, bool conjugate = false);
template <typename T>
void Hemm(LeftOrRight side, UpperOrLower uplo, T alpha, const AbstractDistMatrix<T>& A, const AbstractDistMatrix<T>& B,
          T beta, AbstractDistMatrix<T>& C);
template <typename T>
void Trmm(LeftOrRight side, UpperOrLower uplo, Orientation orientation, UnitOrNonUnit diag, T alpha,
          const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B);

// ---- TwoSidedTrsm / TwoSidedTrmm (src/blas_like/level3/TwoSidedTrsm.cpp:17-40, TwoSidedTrmm.cpp:18-42) ----
// A := inv(L) A inv(L)^H | inv(U)^H A inv(U)   resp.   A := L^H A L | U A U^H on the `uplo` triangle of Hermitian A
template <typename F>
void TwoSidedTrsm(UpperOrLower uplo, UnitOrNonUnit diag, AbstractDistMatrix<F>& A, const AbstractDistMatrix<F>& B);
template <typename F>
void TwoSidedTrmm(UpperOrLower uplo, UnitOrNonUnit diag, AbstractDistMatrix<F>& A, const AbstractDistMatrix<F>& B);

// ---- Trsm (src/blas_like/level3/Trsm.cpp:24-398, Trsm/{LLN,LLT,LUN,LUT,RLN,RLT,RUN,RUT}.hpp) ----
template <typename F>
void Trsm(LeftOrRight side, UpperOrLower uplo, Orientation orientation, UnitOrNonUnit diag, F alpha,
          const Matrix<F>& A, Matrix<F>& B, bool checkIfSingular = false);
template <typename F>
void Trsm(LeftOrRight side, UpperOrLower uplo, Orientation orientation, UnitOrNonUnit diag, F alpha,
          const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B, bool checkIfSingular = false,
          TrsmAlgorithm alg = TRSM_DEFAULT);
template <typename F>
void LocalTrsm(LeftOrRight side, UpperOrLower uplo, Orientation orientation, UnitOrNonUnit diag, F alpha,
               const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& X, bool checkIfSingular = false);
// single right-hand side (src/blas_like/level2/Trsv.cpp:47-68); Trsm(LEFT, ...) dispatches here for width 1
template <typename F>
void Trsv(UpperOrLower uplo, Orientation orientation, UnitOrNonUnit diag, const AbstractDistMatrix<F>& A,
          AbstractDistMatrix<F>& x);

}  // namespace El
