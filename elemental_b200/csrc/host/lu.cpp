// Permutations, LU with / without partial pivoting, lu::SolveAfter and LinearSolve on device-resident matrices.
// Reference: src/lapack_like/perm/Permutation.cpp, DistPermutation.cpp; src/lapack_like/factor/LU.cpp:21-220,
// LU/Panel.hpp, LU/Local.hpp, LU/SolveAfter.hpp; src/lapack_like/solve/Linear.cpp.
//
// The blocked loop is the reference's (LU.cpp:170-220): panel, interchange of the rows outside the panel,
// A12 := L11^{-1} A12 on [*,VR], rank-nb update of A22 from [MC,*] x [*,MR].  B200-first differences:
//   * the panel A(k:m, k:k+nb) is gathered to [*,*] and factored by one cooperative kernel (kernels/lu.cu) -- the
//     reference keeps it [MC,*] and pays a MaxLoc all-reduce plus a row broadcast per COLUMN;
//   * pivots never visit the host: the kernel leaves them in device memory, the interchange kernels and the
//     DistPermutation read them there, and a zero pivot raises SingularMatrixException once, after the sweep
//     (the panel is replicated, so every process sees the same flag);
//   * the interchange of a panel touches at most 2 nb rows: they are packed, all-gathered inside the process
//     column (NCCL, one call) and scattered to their new owners; the reference sends every row through a
//     general permutation (DistPermutation::PermuteRows -> PermutationMeta all-to-all).
#include <algorithm>
#include <memory>
#include <numeric>

#include "dev.hpp"
#include "elb200/lu.hpp"
#include "elb200_plan.h"

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

typedef long long i64;

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// DistPermutation
// ---------------------------------------------------------------------------------------------------------------
DistPermutation::DistPermutation(const El::Grid& g) : grid_(&g) {}
DistPermutation::~DistPermutation() {
    if (swaps_) cudaFreeAsync(swaps_, dev::stream());
    if (vec_) cudaFreeAsync(vec_, dev::stream());
}
void DistPermutation::Empty() {
    if (swaps_) elb200::scratch_free(swaps_, dev::stream());
    if (vec_) elb200::scratch_free(vec_, dev::stream());
    swaps_ = nullptr; vec_ = nullptr;
    size_ = numSwaps_ = capacity_ = vecSize_ = 0;
    implicit_ = true; stale_ = true;
    pre_.clear(); img_.clear();
}
void DistPermutation::MakeIdentity(Int size) {
    if (size < 0) LogicError("Permutation size must be non-negative");
    size_ = size; numSwaps_ = 0; implicit_ = true; stale_ = true;
}
void DistPermutation::ReserveSwaps(Int maxSwaps) {
    if (maxSwaps <= capacity_) return;
    cudaStream_t s = dev::stream();
    i64* fresh = (i64*)elb200::scratch_alloc(sizeof(i64) * 2 * (size_t)maxSwaps, s);
    if (numSwaps_ > 0) {
        ELB_CUDA(cudaMemcpyAsync(fresh, swaps_, sizeof(i64) * (size_t)numSwaps_, cudaMemcpyDeviceToDevice, s));
        ELB_CUDA(cudaMemcpyAsync(fresh + maxSwaps, swaps_ + capacity_, sizeof(i64) * (size_t)numSwaps_,
                                 cudaMemcpyDeviceToDevice, s));
    }
    if (swaps_) elb200::scratch_free(swaps_, s);
    swaps_ = fresh;
    capacity_ = maxSwaps;
}
void DistPermutation::Swap(Int origin, Int dest) {
    if (origin < 0 || origin >= size_ || dest < 0 || dest >= size_) LogicError("Swap index out of range");
    if (numSwaps_ == capacity_) ReserveSwaps(std::max<Int>(2 * capacity_, 16));
    const i64 o = origin, d = dest;
    cudaStream_t s = dev::stream();
    // pageable sources: the copies are staged before the calls return
    ELB_CUDA(cudaMemcpyAsync(swaps_ + numSwaps_, &o, sizeof(i64), cudaMemcpyHostToDevice, s));
    ELB_CUDA(cudaMemcpyAsync(swaps_ + capacity_ + numSwaps_, &d, sizeof(i64), cudaMemcpyHostToDevice, s));
    if (origin != numSwaps_) implicit_ = false;
    ++numSwaps_;
    stale_ = true;
}
void DistPermutation::SwapSequence(const DistPermutation& P, Int offset) {
    const Int count = P.numSwaps_;
    if (count == 0) return;
    if (numSwaps_ + count > capacity_) ReserveSwaps(std::max<Int>(2 * capacity_, numSwaps_ + count));
    cudaStream_t s = dev::stream();
    std::vector<i64> o(count), d(count);
    ELB_CUDA(cudaMemcpyAsync(o.data(), P.swaps_, sizeof(i64) * (size_t)count, cudaMemcpyDeviceToHost, s));
    ELB_CUDA(cudaMemcpyAsync(d.data(), P.swaps_ + P.capacity_, sizeof(i64) * (size_t)count, cudaMemcpyDeviceToHost, s));
    ELB_CUDA(cudaStreamSynchronize(s));
    for (Int j = 0; j < count; ++j) {
        o[j] += offset; d[j] += offset;
        if (o[j] < 0 || o[j] >= size_ || d[j] < 0 || d[j] >= size_) LogicError("Swap index out of range");
        if (o[j] != numSwaps_ + j) implicit_ = false;
    }
    ELB_CUDA(cudaMemcpyAsync(swaps_ + numSwaps_, o.data(), sizeof(i64) * (size_t)count, cudaMemcpyHostToDevice, s));
    ELB_CUDA(cudaMemcpyAsync(swaps_ + capacity_ + numSwaps_, d.data(), sizeof(i64) * (size_t)count, cudaMemcpyHostToDevice, s));
    ELB_CUDA(cudaStreamSynchronize(s));   // o, d die with this frame
    numSwaps_ += count;
    stale_ = true;
}
void DistPermutation::AppendDeviceSwaps(const long long* ipivDev, Int count, Int offset) {
    if (count <= 0) return;
    if (numSwaps_ + count > capacity_) ReserveSwaps(std::max<Int>(2 * capacity_, numSwaps_ + count));
    elb200::append_swaps_device(swaps_, swaps_ + capacity_, numSwaps_, ipivDev, count, offset, dev::stream());
    if (offset != numSwaps_) implicit_ = false;
    numSwaps_ += count;
    stale_ = true;
}
void DistPermutation::Compose() const {
    if (!stale_) return;
    cudaStream_t s = dev::stream();
    std::vector<i64> o(numSwaps_), d(numSwaps_);
    if (numSwaps_ > 0) {
        ELB_CUDA(cudaMemcpyAsync(o.data(), swaps_, sizeof(i64) * (size_t)numSwaps_, cudaMemcpyDeviceToHost, s));
        ELB_CUDA(cudaMemcpyAsync(d.data(), swaps_ + capacity_, sizeof(i64) * (size_t)numSwaps_, cudaMemcpyDeviceToHost, s));
        ELB_CUDA(cudaStreamSynchronize(s));
    }
    pre_.resize(size_);
    img_.resize(size_);
    // the bookkeeping itself is plain host code with a C entry point (include/elb200_plan.h), tested on the CPU
    static_assert(sizeof(i64) == sizeof(int64_t), "swap lists are 64-bit");
    if (elb200_perm_compose(size_, numSwaps_, (const int64_t*)o.data(), (const int64_t*)d.data(), (int64_t*)pre_.data(),
                            (int64_t*)img_.data()) != 0)
        RuntimeError("Corrupt swap sequence");
    if (vecSize_ < size_) {
        if (vec_) elb200::scratch_free(vec_, s);
        vec_ = (i64*)elb200::scratch_alloc(sizeof(i64) * 2 * (size_t)std::max<Int>(size_, 1), s);
        vecSize_ = size_;
    }
    if (size_ > 0) {
        ELB_CUDA(cudaMemcpyAsync(vec_, pre_.data(), sizeof(i64) * (size_t)size_, cudaMemcpyHostToDevice, s));
        ELB_CUDA(cudaMemcpyAsync(vec_ + vecSize_, img_.data(), sizeof(i64) * (size_t)size_, cudaMemcpyHostToDevice, s));
        ELB_CUDA(cudaStreamSynchronize(s));
    }
    stale_ = false;
}
const long long* DistPermutation::DeviceVector(bool inverse) const {
    Compose();
    return inverse ? vec_ + vecSize_ : vec_;
}
bool DistPermutation::Parity() const {
    Compose();
    return elb200_perm_parity(size_, (const int64_t*)pre_.data()) != 0;
}
Int DistPermutation::Image(Int origin) const {
    if (origin < 0 || origin >= size_) LogicError("Index out of range");
    Compose();
    return (Int)img_[origin];
}
Int DistPermutation::Preimage(Int dest) const {
    if (dest < 0 || dest >= size_) LogicError("Index out of range");
    Compose();
    return (Int)pre_[dest];
}
std::vector<Int> DistPermutation::Preimages() const {
    Compose();
    return std::vector<Int>(pre_.begin(), pre_.end());
}

// rows: A(offset + i, :) := A_old(offset + v[i], :); columns likewise.  The old rows are read from a copy whose
// permuted dimension is not distributed ([*,V] / [U,*]: an all-gather inside the process column / row).
template <typename T>
void DistPermutation::Apply(AbstractDistMatrix<T>& A, Int offset, bool rows, bool inverse) const {
    if (size_ == 0) return;
    if (offset < 0 || offset + size_ > (rows ? A.Height() : A.Width())) LogicError("Permutation does not fit the matrix");
    const long long* v = DeviceVector(inverse);
    const El::Grid& g = A.Grid();
    cudaStream_t s = dev::stream();
    if (rows) {
        auto V = View(A, offset, 0, size_, A.Width());
        AbstractDistMatrix<T> old(g, STAR, A.RowDist());
        old.AlignRows(V.RowAlign());
        Copy(C(V), old);
        elb200::permute_device<dev::D<T>>(true, V.LocalHeight(), V.LocalWidth(), v, V.ColShift(), V.ColStride(),
                                          dev::ptr(old.LockedBuffer()), old.LDim(), dev::ptr(V.Buffer()), V.LDim(), s);
    } else {
        auto V = View(A, 0, offset, A.Height(), size_);
        AbstractDistMatrix<T> old(g, A.ColDist(), STAR);
        old.AlignCols(V.ColAlign());
        Copy(C(V), old);
        elb200::permute_device<dev::D<T>>(false, V.LocalHeight(), V.LocalWidth(), v, V.RowShift(), V.RowStride(),
                                          dev::ptr(old.LockedBuffer()), old.LDim(), dev::ptr(V.Buffer()), V.LDim(), s);
    }
}
template <typename T> void DistPermutation::PermuteRows(AbstractDistMatrix<T>& A, Int offset) const { Apply(A, offset, true, false); }
template <typename T> void DistPermutation::InversePermuteRows(AbstractDistMatrix<T>& A, Int offset) const { Apply(A, offset, true, true); }
template <typename T> void DistPermutation::PermuteCols(AbstractDistMatrix<T>& A, Int offset) const { Apply(A, offset, false, false); }
template <typename T> void DistPermutation::InversePermuteCols(AbstractDistMatrix<T>& A, Int offset) const { Apply(A, offset, false, true); }

namespace {
// a local matrix seen as a [*,*] matrix of the trivial grid
template <typename T>
AbstractDistMatrix<T> AsStarStar(Matrix<T>& A) {
    AbstractDistMatrix<T> D(El::Grid::Default(), STAR, STAR);
    D.Attach(A.Height(), A.Width(), El::Grid::Default(), 0, 0, A.Buffer(), A.LDim());
    return D;
}
template <typename T>
AbstractDistMatrix<T> AsMcMr(Matrix<T>& A) {
    AbstractDistMatrix<T> D(El::Grid::Default(), MC, MR);
    D.Attach(A.Height(), A.Width(), El::Grid::Default(), 0, 0, A.Buffer(), A.LDim());
    return D;
}
template <typename T>
AbstractDistMatrix<T> AsMcMr(const Matrix<T>& A) {
    AbstractDistMatrix<T> D(El::Grid::Default(), MC, MR);
    D.LockedAttach(A.Height(), A.Width(), El::Grid::Default(), 0, 0, A.LockedBuffer(), A.LDim());
    return D;
}
}  // namespace
template <typename T> void DistPermutation::PermuteRows(Matrix<T>& A, Int offset) const { auto D = AsStarStar(A); Apply(D, offset, true, false); }
template <typename T> void DistPermutation::InversePermuteRows(Matrix<T>& A, Int offset) const { auto D = AsStarStar(A); Apply(D, offset, true, true); }
template <typename T> void DistPermutation::PermuteCols(Matrix<T>& A, Int offset) const { auto D = AsStarStar(A); Apply(D, offset, false, false); }
template <typename T> void DistPermutation::InversePermuteCols(Matrix<T>& A, Int offset) const { auto D = AsStarStar(A); Apply(D, offset, false, true); }

// ---------------------------------------------------------------------------------------------------------------
// LU
// ---------------------------------------------------------------------------------------------------------------
namespace {

// rows k .. of every local column of A take the interchanges of one panel (ipiv: device, relative to row k)
template <typename F>
struct PanelSwapper {
    typedef dev::D<F> D;
    i64* slotRow = nullptr;
    int* srcSlot = nullptr;
    Int cap;
    explicit PanelSwapper(Int maxNb) : cap(maxNb) {
        cudaStream_t s = dev::stream();
        slotRow = (i64*)elb200::scratch_alloc(sizeof(i64) * 2 * (size_t)cap + sizeof(int) * 2 * (size_t)cap, s);
        srcSlot = (int*)(slotRow + 2 * cap);
    }
    ~PanelSwapper() { if (slotRow) cudaFreeAsync(slotRow, dev::stream()); }
    void Run(AbstractDistMatrix<F>& A, Int k, Int nb, const i64* ipiv) {
        cudaStream_t s = dev::stream();
        const El::Grid& g = A.Grid();
        const i64 nloc = A.LocalWidth();
        const int S = 2 * (int)nb, r = A.ColStride();
        elb200::swap_plan_device((int)nb, ipiv, k, slotRow, srcSlot, s);
        if (nloc == 0) return;
        const size_t per = (size_t)S * (size_t)nloc;
        D* buf = (D*)elb200::scratch_alloc(sizeof(D) * per * (size_t)(r > 1 ? r + 1 : 1), s);
        elb200::pack_rows_device<D>(S, slotRow, srcSlot, dev::ptr(A.LockedBuffer()), A.LDim(), nloc, A.ColAlign(), r,
                                    A.ColRank(), A.ColShift(), buf, s);
        const D* all = buf;
        if (r > 1) {
            // every process of the column contributes its slots (fixed size: unowned slots travel as padding)
            const Comm& col = A.ColDist() == MC ? g.MCComm() : g.MRComm();
            ELB_NCCL(ncclAllGather(buf, buf + per, per * sizeof(D), ncclInt8, (ncclComm_t)col.nccl, s));
            GetRedistStats().allGathers++;
            all = buf + per;
        }
        elb200::unpack_rows_device<D>(S, slotRow, srcSlot, dev::ptr(A.Buffer()), A.LDim(), nloc, A.ColAlign(), r, A.ColRank(),
                                      A.ColShift(), all, (i64)per, s);
        elb200::scratch_free(buf, s);
    }
};

// A [MC,MR]; P == nullptr: no pivoting
template <typename F>
void LUBlocked(AbstractDistMatrix<F>& A, DistPermutation* P, dev::DeviceFlag& info) {
    typedef dev::D<F> D;
    const El::Grid& g = A.Grid();
    const Int m = A.Height(), n = A.Width(), minDim = std::min(m, n);
    const Int bsize = Blocksize();
    if (bsize > 512) LogicError("LU: Blocksize() above 512 is not supported by the panel kernel");
    cudaStream_t s = dev::stream();
    if (P) {
        P->SetGrid(g);
        P->MakeIdentity(m);
        P->ReserveSwaps(minDim);
    }
    if (minDim == 0) return;
    AbstractDistMatrix<F> panel(g, STAR, STAR), A12_STAR_VR(g, STAR, VR), A12_STAR_MR(g, STAR, MR), A21_MC_STAR(g, MC, STAR);
    i64* ipiv = P ? (i64*)elb200::scratch_alloc(sizeof(i64) * (size_t)bsize, s) : nullptr;
    std::unique_ptr<PanelSwapper<F>> swapper;
    if (P) swapper.reset(new PanelSwapper<F>(bsize));
    dev::PhaseTimer tm;   // ELB200_TRACE=1: serialised per-phase device times
    for (Int k = 0; k < minDim; k += bsize) {
        const Int nb = std::min(bsize, minDim - k);
        auto AB1 = View(A, k, k, m - k, nb);
        tm.Begin(s);
        Copy(C(AB1), panel);
        tm.End(s, "panel -> [*,*]");
        tm.Begin(s);
        elb200::getrf_panel_device<D>(m - k, nb, dev::ptr(panel.Buffer()), panel.LDim(), ipiv, P != nullptr, info.dev_, k, s);
        tm.End(s, "getrf(panel)");
        if (P) {
            tm.Begin(s);
            P->AppendDeviceSwaps(ipiv, nb, k);
            swapper->Run(A, k, nb, ipiv);   // all columns: the panel's own are overwritten next
            tm.End(s, "row interchanges");
        }
        tm.Begin(s);
        Copy(C(panel), AB1);
        tm.End(s, "panel <- [*,*]");
        if (k + nb < n) {
            auto A12 = View(A, k, k + nb, nb, n - k - nb);
            auto A11 = LockedView(C(panel), 0, 0, nb, nb);
            A12_STAR_VR.AlignWith(A12);
            tm.Begin(s);
            Copy(C(A12), A12_STAR_VR);
            tm.End(s, "A12 -> [*,VR]");
            tm.Begin(s);
            LocalTrsm(LEFT, LOWER, NORMAL, UNIT, F(1), C(A11), A12_STAR_VR);
            tm.End(s, "trsm(A12)");
            A12_STAR_MR.AlignWith(A12);
            tm.Begin(s);
            Copy(C(A12_STAR_VR), A12_STAR_MR);
            tm.End(s, "A12 [*,VR] -> [*,MR]");
            if (k + nb < m) {
                auto A22 = View(A, k + nb, k + nb, m - k - nb, n - k - nb);
                auto L21 = LockedView(C(panel), nb, 0, m - k - nb, nb);
                A21_MC_STAR.AlignWith(A22);
                tm.Begin(s);
                Copy(C(L21), A21_MC_STAR);
                tm.End(s, "L21 [*,*] -> [MC,*]");
                tm.Begin(s);
                LocalGemm(NORMAL, NORMAL, F(-1), C(A21_MC_STAR), C(A12_STAR_MR), F(1), A22);
                tm.End(s, "trailing update (gemm)");
            }
            tm.Begin(s);
            Copy(C(A12_STAR_MR), A12);
            tm.End(s, "A12 <- [*,MR]");
        }
    }
    tm.Report("LU");
    if (ipiv) elb200::scratch_free(ipiv, s);
}

template <typename F>
void LUDriver(AbstractDistMatrix<F>& APre, DistPermutation* P) {
    dev::DeviceFlag info;
    if (APre.ColDist() == MC && APre.RowDist() == MR) {
        LUBlocked(APre, P, info);
    } else {
        AbstractDistMatrix<F> A(APre.Grid(), MC, MR);
        Copy(C(APre), A);
        LUBlocked(A, P, info);
        Copy(C(A), APre);
    }
    if (info.Read() != 0) throw SingularMatrixException();
}

}  // namespace

template <typename F> void LU(AbstractDistMatrix<F>& A) { LUDriver<F>(A, nullptr); }
template <typename F> void LU(AbstractDistMatrix<F>& A, DistPermutation& P) { LUDriver<F>(A, &P); }
template <typename F> void LU(Matrix<F>& A) { auto D = AsMcMr(A); LUDriver<F>(D, nullptr); }
template <typename F> void LU(Matrix<F>& A, Permutation& P) { auto D = AsMcMr(A); LUDriver<F>(D, &P); P.SetGrid(El::Grid::Default()); }

namespace lu {

template <typename F>
void SolveAfter(Orientation o, const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B) {
    if (A.Height() != A.Width()) LogicError("A must be square");
    if (A.Height() != B.Height()) LogicError("A and B must be the same height");
    if (o == NORMAL) {
        Trsm(LEFT, LOWER, NORMAL, UNIT, F(1), A, B);
        Trsm(LEFT, UPPER, NORMAL, NON_UNIT, F(1), A, B);
    } else {
        Trsm(LEFT, UPPER, o, NON_UNIT, F(1), A, B);
        Trsm(LEFT, LOWER, o, UNIT, F(1), A, B);
    }
}
template <typename F>
void SolveAfter(Orientation o, const AbstractDistMatrix<F>& A, const DistPermutation& P, AbstractDistMatrix<F>& B) {
    if (A.Height() != A.Width()) LogicError("A must be square");
    if (A.Height() != B.Height()) LogicError("A and B must be the same height");
    if (o == NORMAL) {
        P.PermuteRows(B);
        Trsm(LEFT, LOWER, NORMAL, UNIT, F(1), A, B);
        Trsm(LEFT, UPPER, NORMAL, NON_UNIT, F(1), A, B);
    } else {
        Trsm(LEFT, UPPER, o, NON_UNIT, F(1), A, B);
        Trsm(LEFT, LOWER, o, UNIT, F(1), A, B);
        P.InversePermuteRows(B);
    }
}
template <typename F>
void SolveAfter(Orientation o, const Matrix<F>& A, Matrix<F>& B) {
    auto DA = AsMcMr(A);
    auto DB = AsMcMr(B);
    SolveAfter(o, C(DA), DB);
}
template <typename F>
void SolveAfter(Orientation o, const Matrix<F>& A, const Permutation& P, Matrix<F>& B) {
    auto DA = AsMcMr(A);
    auto DB = AsMcMr(B);
    SolveAfter(o, C(DA), P, DB);
}

}  // namespace lu

template <typename F>
void LinearSolve(const AbstractDistMatrix<F>& A, AbstractDistMatrix<F>& B, bool) {
    if (A.Height() != A.Width()) LogicError("A must be square");
    if (A.Height() != B.Height()) LogicError("A and B must be the same height");
    // Linear.cpp factors [A B] by RowEchelon and back-substitutes; the same row operations reach B here through
    // P, L^{-1} and U^{-1} applied after the factorisation of a copy of A
    AbstractDistMatrix<F> F_(A.Grid(), MC, MR);
    Copy(A, F_);
    DistPermutation P(A.Grid());
    LU(F_, P);
    lu::SolveAfter(NORMAL, C(F_), P, B);
}
template <typename F>
void LinearSolve(const Matrix<F>& A, Matrix<F>& B) {
    auto DA = AsMcMr(A);
    auto DB = AsMcMr(B);
    LinearSolve(C(DA), DB, false);
}

#define ELB_INST(F)                                                                                                     \
    template void DistPermutation::PermuteRows(AbstractDistMatrix<F>&, Int) const;                                       \
    template void DistPermutation::InversePermuteRows(AbstractDistMatrix<F>&, Int) const;                                \
    template void DistPermutation::PermuteCols(AbstractDistMatrix<F>&, Int) const;                                       \
    template void DistPermutation::InversePermuteCols(AbstractDistMatrix<F>&, Int) const;                                \
    template void DistPermutation::PermuteRows(Matrix<F>&, Int) const;                                                   \
    template void DistPermutation::InversePermuteRows(Matrix<F>&, Int) const;                                            \
    template void DistPermutation::PermuteCols(Matrix<F>&, Int) const;                                                   \
    template void DistPermutation::InversePermuteCols(Matrix<F>&, Int) const;                                            \
    template void LU(Matrix<F>&);                                                                                        \
    template void LU(AbstractDistMatrix<F>&);                                                                            \
    template void LU(Matrix<F>&, Permutation&);                                                                          \
    template void LU(AbstractDistMatrix<F>&, DistPermutation&);                                                          \
    template void lu::SolveAfter(Orientation, const Matrix<F>&, Matrix<F>&);                                             \
    template void lu::SolveAfter(Orientation, const AbstractDistMatrix<F>&, AbstractDistMatrix<F>&);                     \
    template void lu::SolveAfter(Orientation, const Matrix<F>&, const Permutation&, Matrix<F>&);                         \
    template void lu::SolveAfter(Orientation, const AbstractDistMatrix<F>&, const DistPermutation&, AbstractDistMatrix<F>&); \
    template void LinearSolve(const Matrix<F>&, Matrix<F>&);                                                             \
    template void LinearSolve(const AbstractDistMatrix<F>&, AbstractDistMatrix<F>&, bool);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
