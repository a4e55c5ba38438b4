// elb200 host layer, part 5: permutations, LU with and without partial pivoting, the solves after it and
// LinearSolve, with the reference's signatures (include/El/core/DistPermutation.hpp:70-170,
// include/El/lapack_like/factor.hpp LU / lu::SolveAfter, include/El/lapack_like/solve.hpp:14-30).
#pragma once
#include <vector>

#include "elb200/factor.hpp"

namespace El {

// A permutation of `size` indices held as a sequence of swaps (origin_j <-> dest_j, applied in order), exactly the
// representation the reference's factorisations append to (Permutation.cpp:207-250).  B200-first: the swap list lives
// in DEVICE memory, so a factorisation appends its pivots without a host round trip; the explicit vectors are
// composed on the host only when a caller asks for them (Image / Preimage / Parity) or applies the permutation
// to a whole matrix.  Row i of P A is row Preimage(i) of A.
class DistPermutation {
public:
    explicit DistPermutation(const El::Grid& g = El::Grid::Default());
    ~DistPermutation();
    DistPermutation(const DistPermutation&) = delete;
    DistPermutation& operator=(const DistPermutation&) = delete;

    void SetGrid(const El::Grid& g) { grid_ = &g; }
    const El::Grid& Grid() const { return *grid_; }
    void Empty();
    void MakeIdentity(Int size);
    void ReserveSwaps(Int maxSwaps);
    void Swap(Int origin, Int dest);
    void SwapSequence(const DistPermutation& P, Int offset = 0);
    Int Height() const { return size_; }
    Int Width() const { return size_; }
    Int NumSwaps() const { return numSwaps_; }
    bool IsSwapSequence() const { return true; }
    bool IsImplicitSwapSequence() const { return implicit_; }
    bool Parity() const;               // true when the permutation is odd
    Int Image(Int origin) const;       // where row `origin` goes
    Int Preimage(Int dest) const;      // which row lands in `dest`
    std::vector<Int> Preimages() const;   // the whole vector (host)

    // A := P A, A := P^{-1} A, A := A P^T (columns permuted the same way), and the inverse; rows / columns
    // offset .. offset + Height() - 1 of A are affected (Permutation.cpp:545-720)
    template <typename T> void PermuteRows(AbstractDistMatrix<T>& A, Int offset = 0) const;
    template <typename T> void InversePermuteRows(AbstractDistMatrix<T>& A, Int offset = 0) const;
    template <typename T> void PermuteCols(AbstractDistMatrix<T>& A, Int offset = 0) const;
    template <typename T> void InversePermuteCols(AbstractDistMatrix<T>& A, Int offset = 0) const;
    template <typename T> void PermuteRows(Matrix<T>& A, Int offset = 0) const;
    template <typename T> void InversePermuteRows(Matrix<T>& A, Int offset = 0) const;
    template <typename T> void PermuteCols(Matrix<T>& A, Int offset = 0) const;
    template <typename T> void InversePermuteCols(Matrix<T>& A, Int offset = 0) const;

    // used by the factorisations: append `count` swaps (offset + j) <-> (offset + ipivDev[j]), j = 0 .. count-1,
    // read from DEVICE memory on the layer's current stream
    void AppendDeviceSwaps(const long long* ipivDev, Int count, Int offset);
    // device vectors of length Height(): preimages (inverse = false) or images (inverse = true)
    const long long* DeviceVector(bool inverse) const;

private:
    const El::Grid* grid_;
    Int size_ = 0, numSwaps_ = 0, capacity_ = 0;
    bool implicit_ = true;
    long long* swaps_ = nullptr;   // device: origins at [0, capacity), destinations at [capacity, 2 capacity)
    mutable bool stale_ = true;
    mutable std::vector<long long> pre_, img_;
    mutable long long* vec_ = nullptr;   // device: preimages then images
    mutable Int vecSize_ = 0;
    void Compose() const;
    template <typename T> void Apply(AbstractDistMatrix<T>& A, Int offset, bool rows, bool inverse) const;
};
typedef DistPermutation Permutation;   // local matrices are device-resident too: one class serves both

// LU without pivoting (src/lapack_like/factor/LU.cpp:21-99): A = L U, unit-lower L and U packed in place.
// Throws SingularMatrixException on an exactly zero pivot (LU/Local.hpp:53-55).
template <typename F> void LU(Matrix<F>& A);
template <typename F> void LU(AbstractDistMatrix<F>& A);
// LU with partial pivoting (LU.cpp:104-220): P A = L U
template <typename F> void LU(Matrix<F>& A, Permutation& P);
template <typename F> void LU(AbstractDistMatrix<F>& A, DistPermutation& P);

namespace lu {
// B := op(A)^{-1} B from the packed factors (LU/SolveAfter.hpp:14-129)
template <typename F> void SolveAfter(Orientation orientation, const Matrix<F>& A, Matrix<F>& B);
template <typename F> void SolveAfter(Orientation orientation, const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B);
template <typename F> void SolveAfter(Orientation orientation, const Matrix<F>& A, const Permutation& P, Matrix<F>& B);
template <typename F>
void SolveAfter(Orientation orientation, const AbstractDistMatrix<F>& A, const DistPermutation& P, AbstractDistMatrix<F>& B);
}  // namespace lu

// B := A^{-1} B by LU with partial pivoting on a copy of A (src/lapack_like/solve/Linear.cpp:170-230)
template <typename F> void LinearSolve(const Matrix<F>& A, Matrix<F>& B);
template <typename F> void LinearSolve(const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B, bool scalapack = false);

// ---- the rest of the Cholesky family (include/El/lapack_like/factor.hpp:36-90) ----
// Diagonally pivoted factorisation P A P^T = L L^H (LOWER) / U^H U (UPPER) of a Hermitian positive (semi)definite
// matrix, pivoting on the largest remaining diagonal entry; only the uplo triangle of A is read and written.
template <typename F> void Cholesky(UpperOrLower uplo, Matrix<F>& A, Permutation& P);
template <typename F> void Cholesky(UpperOrLower uplo, AbstractDistMatrix<F>& A, DistPermutation& P);
namespace cholesky {
template <typename F>
void SolveAfter(UpperOrLower uplo, Orientation orientation, const AbstractDistMatrix<F>& A, const DistPermutation& P,
                AbstractDistMatrix<F>& B);
}  // namespace cholesky
// T := factor of T T^H + alpha V V^H (LOWER) / T^H T + alpha V V^H (UPPER); V (n x w) is overwritten with workspace.
// A downdate that would make the matrix indefinite throws std::logic_error, as the reference's reflector does.
template <typename F> void CholeskyMod(UpperOrLower uplo, Matrix<F>& T, Base<F> alpha, Matrix<F>& V);
template <typename F> void CholeskyMod(UpperOrLower uplo, AbstractDistMatrix<F>& T, Base<F> alpha, AbstractDistMatrix<F>& V);

}  // namespace El
