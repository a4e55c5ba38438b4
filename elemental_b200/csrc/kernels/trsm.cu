// Triangular solve with many right-hand sides, all side/uplo/trans/diag combos,
// every scalar type.  Replaces blas::Trsm (reference include/El/core/imports/blas.hpp:884-913
// -> ?trsm_, src/core/imports/blas/Trsm.hpp:359-393) as called from LocalTrsm
// (src/blas_like/level3/Trsm.cpp:377-398).
//
// Method (block substitution with inverted diagonal blocks, the standard GPU
// formulation): (1) one warp per 32x32 diagonal block inverts it in registers /
// shared memory; (2) for every block step, a small in-place kernel applies the
// inverted block to its 32 rows (LEFT) or columns (RIGHT) of B and (3) one GEMM
// on the tensor pipe (gemm_device) updates the rest of B.  All the O(m n^2)
// flops of the solve are in step (3).
#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int TB = 32;

// Inverse of each TB x TB diagonal block of the stored triangle of A (n x n).
// inv holds nblk dense TB x TB column-major blocks (other triangle zero).
template <class T>
__global__ void __launch_bounds__(32) trtri_diag_kernel(int lower, int unit, i64 n, const T* __restrict__ A,
                                                        i64 lda, T* __restrict__ inv) {
    typedef scalar_traits<T> st;
    typedef typename st::real R;
    __shared__ T sL[TB][TB + 1];
    __shared__ T sX[TB][TB + 1];
    const i64 b0 = (i64)blockIdx.x * TB;
    const int nb = (int)((n - b0 < TB) ? (n - b0) : TB);
    const int j = threadIdx.x;
    // load as a LOWER triangular block: for upper storage read the transpose (no conj):
    // inv(U) = (inv(U^T))^T and U^T is lower.
    for (int c = 0; c < TB; ++c) {
        T v = st::zero();
        if (j < nb && c < nb) {
            if (lower) { if (j >= c) v = A[(b0 + j) + (b0 + c) * lda]; }
            else { if (j >= c) v = A[(b0 + c) + (b0 + j) * lda]; }
        }
        if (j == c) { if (unit || j >= nb) v = st::from_real(R(1)); }
        sL[j][c] = v;
        sX[j][c] = st::zero();
    }
    __syncwarp();
    // thread j computes column j of X = L^-1 by forward substitution
    {
        // 1 / l_jj
        const T d = sL[j][j];
        const R den = st::abs2(d);
        const T dinv = st::conj(d) * (R(1) / den);
        sX[j][j] = dinv;
        for (int i = j + 1; i < TB; ++i) {
            T acc = st::zero();
            for (int k = j; k < i; ++k) acc += sL[i][k] * sX[k][j];
            const T di = sL[i][i];
            const T diinv = st::conj(di) * (R(1) / st::abs2(di));
            sX[i][j] = -(acc * diinv);
        }
    }
    __syncwarp();
    T* out = inv + (i64)blockIdx.x * TB * TB;
    for (int c = 0; c < TB; ++c) {
        // out(row=j, col=c): lower -> X[j][c]; upper -> transpose back
        out[j + c * TB] = lower ? sX[j][c] : sX[c][j];
    }
}

// In-place application of one inverted diagonal block.
// LEFT : B(r0:r0+nb, :) := op(Inv) * B(r0:r0+nb, :)     (CTA per 32-column slab)
// RIGHT: B(:, c0:c0+nb) := B(:, c0:c0+nb) * op(Inv)     (CTA per 32-row slab)
// top: 0 none, 1 transpose, 2 conjugate-transpose of the stored inverse block.
template <class T>
__global__ void __launch_bounds__(256) apply_inv_kernel(int left, int top, int nb, i64 m, i64 n, T alpha,
                                                        const T* __restrict__ invblk, T* __restrict__ B, i64 ldb,
                                                        i64 off) {
    typedef scalar_traits<T> st;
    __shared__ T sI[TB][TB + 1];   // sI[r][c] = op(Inv)(r,c)
    __shared__ T sB[TB][TB + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int c = ty; c < TB; c += 8) {
        T v = invblk[tx + c * TB];  // Inv(tx, c)
        if (top == 0) sI[tx][c] = v;
        else sI[c][tx] = (top == 2) ? st::conj(v) : v;
    }
    if (left) {
        // slab of 32 columns starting at j0; rows off..off+nb
        const i64 j0 = (i64)blockIdx.x * TB;
        for (int c = ty; c < TB; c += 8) {
            T v = st::zero();
            if (tx < nb && j0 + c < n) v = B[(off + tx) + (j0 + c) * ldb];
            sB[tx][c] = v;  // sB[row][col]
        }
        __syncthreads();
        for (int c = ty; c < TB; c += 8) {
            if (tx < nb && j0 + c < n) {
                T acc = st::zero();
                for (int k = 0; k < nb; ++k) acc += sI[tx][k] * sB[k][c];
                B[(off + tx) + (j0 + c) * ldb] = alpha * acc;
            }
        }
    } else {
        // slab of 32 rows starting at i0; columns off..off+nb
        const i64 i0 = (i64)blockIdx.x * TB;
        for (int c = ty; c < TB; c += 8) {
            T v = st::zero();
            if (i0 + tx < m && c < nb) v = B[(i0 + tx) + (off + c) * ldb];
            sB[tx][c] = v;
        }
        __syncthreads();
        for (int c = ty; c < TB; c += 8) {
            if (i0 + tx < m && c < nb) {
                T acc = st::zero();
                for (int k = 0; k < nb; ++k) acc += sB[tx][k] * sI[k][c];
                B[(i0 + tx) + (off + c) * ldb] = alpha * acc;
            }
        }
    }
}

}  // namespace

void* scratch_alloc(size_t bytes, cudaStream_t s) {
    static bool tuned = false;
    if (!tuned) {
        int dev = 0;
        ELB_CUDA(cudaGetDevice(&dev));
        cudaMemPool_t pool;
        ELB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long thr = ~0ULL;
        ELB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        tuned = true;
    }
    void* p = nullptr;
    ELB_CUDA(cudaMallocAsync(&p, bytes ? bytes : 16, s));
    return p;
}
void scratch_free(void* p, cudaStream_t s) {
    if (p) ELB_CUDA(cudaFreeAsync(p, s));
}

// Block-substitution core; B must already carry alpha.
template <class T>
static void trsm_core(char side, char uplo, char trans, char diag, i64 m, i64 n, const T* A, i64 lda,
                      T* B, i64 ldb, cudaStream_t s) {
    typedef scalar_traits<T> st;
    const bool left = side == 'L';
    const i64 na = left ? m : n;  // order of A
    const i64 nblk = ceil_div(na, TB);
    T* inv = (T*)scratch_alloc(sizeof(T) * (size_t)nblk * TB * TB, s);
    trtri_diag_kernel<T><<<(unsigned)nblk, 32, 0, s>>>(uplo == 'L', diag == 'U', na, A, lda, inv);
    ELB_LAUNCH_CHECK();

    const bool tr = trans != 'N';
    const int top = trans == 'N' ? 0 : (trans == 'T' ? 1 : 2);
    const bool eff_lower = (uplo == 'L') != tr;  // op(A) is lower triangular
    const T one = st::from_real(1), minus_one = st::from_real(-1);
    // forward sweep (block 0 first) when: LEFT & op(A) lower, or RIGHT & op(A) upper
    const bool forward = left ? eff_lower : !eff_lower;
    for (i64 step = 0; step < nblk; ++step) {
        const i64 kb = forward ? step : (nblk - 1 - step);
        const i64 k0 = kb * TB;
        const int nb = (int)((na - k0 < TB) ? (na - k0) : TB);
        if (left) {
            // X_k := op(A_kk)^-1 B_k
            apply_inv_kernel<T><<<(unsigned)ceil_div(n, TB), 256, 0, s>>>(1, top, nb, m, n, one,
                                                                         inv + kb * TB * TB, B, ldb, k0);
            ELB_LAUNCH_CHECK();
            // B_rest -= op(A)(rest,k) X_k
            const i64 r0 = forward ? k0 + nb : 0;
            const i64 rl = forward ? m - (k0 + nb) : k0;
            if (rl > 0) {
                if (!tr)
                    gemm_device<T>(0, 'N', 'N', rl, n, nb, minus_one, A + r0 + k0 * lda, lda, B + k0, ldb,
                                   one, B + r0, ldb, 0, 1, 0, 1, s);
                else
                    gemm_device<T>(0, trans, 'N', rl, n, nb, minus_one, A + k0 + r0 * lda, lda, B + k0, ldb,
                                   one, B + r0, ldb, 0, 1, 0, 1, s);
            }
        } else {
            // X_k := B_k op(A_kk)^-1
            apply_inv_kernel<T><<<(unsigned)ceil_div(m, TB), 256, 0, s>>>(0, top, nb, m, n, one,
                                                                         inv + kb * TB * TB, B, ldb, k0);
            ELB_LAUNCH_CHECK();
            // B(:,rest) -= X_k op(A)(k,rest)
            const i64 c0 = forward ? k0 + nb : 0;
            const i64 cl = forward ? n - (k0 + nb) : k0;
            if (cl > 0) {
                if (!tr)
                    gemm_device<T>(0, 'N', 'N', m, cl, nb, minus_one, B + k0 * ldb, ldb, A + k0 + c0 * lda,
                                   lda, one, B + c0 * ldb, ldb, 0, 1, 0, 1, s);
                else
                    gemm_device<T>(0, 'N', trans, m, cl, nb, minus_one, B + k0 * ldb, ldb, A + c0 + k0 * lda,
                                   lda, one, B + c0 * ldb, ldb, 0, 1, 0, 1, s);
            }
        }
    }
    scratch_free(inv, s);
}

template <class T>
void trsm_device(char side_, char uplo_, char trans_, char diag_, i64 m, i64 n, T alpha, const T* A,
                 i64 lda, T* B, i64 ldb, cudaStream_t s) {
    typedef scalar_traits<T> st;
    const char side = up(side_), uplo = up(uplo_), diag = up(diag_);
    char trans = up(trans_);
    if (side != 'L' && side != 'R') throw std::logic_error("trsm: side must be 'L' or 'R'");
    if (uplo != 'L' && uplo != 'U') throw std::logic_error("trsm: uplo must be 'L' or 'U'");
    if (trans != 'N' && trans != 'T' && trans != 'C') throw std::logic_error("trsm: invalid trans");
    if (diag != 'N' && diag != 'U') throw std::logic_error("trsm: diag must be 'N' or 'U'");
    if (m < 0 || n < 0) throw std::logic_error("trsm: negative dimension");
    if (!st::is_complex && trans == 'C') trans = 'T';
    if (m == 0 || n == 0) return;
    const i64 na = side == 'L' ? m : n;
    if (lda < (na > 1 ? na : 1) || ldb < (m > 1 ? m : 1))
        throw std::logic_error("trsm: leading dimension too small");
    // B := alpha B once, then an alpha-free substitution
    if (!st::is_one(alpha))
        lattice_copy_device<T>(B, B, m, n, 0, 1, ldb, 0, 1, ldb, false, &alpha, false, s);
    trsm_core<T>(side, uplo, trans, diag, m, n, A, lda, B, ldb, s);
}

template void trsm_device<float>(char, char, char, char, i64, i64, float, const float*, i64, float*, i64, cudaStream_t);
template void trsm_device<double>(char, char, char, char, i64, i64, double, const double*, i64, double*, i64, cudaStream_t);
template void trsm_device<c32_t>(char, char, char, char, i64, i64, c32_t, const c32_t*, i64, c32_t*, i64, cudaStream_t);
template void trsm_device<c64_t>(char, char, char, char, i64, i64, c64_t, const c64_t*, i64, c64_t*, i64, cudaStream_t);

}  // namespace elb200

extern "C" {
using namespace elb200;
int elb200_dtrsm(char side, char uplo, char trans, char diag, int64_t m, int64_t n, double alpha,
                 const double* A, int64_t lda, double* B, int64_t ldb, elb200_stream_t s) {
    return guarded([&] { trsm_device<double>(side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, (cudaStream_t)s); });
}
int elb200_strsm(char side, char uplo, char trans, char diag, int64_t m, int64_t n, float alpha,
                 const float* A, int64_t lda, float* B, int64_t ldb, elb200_stream_t s) {
    return guarded([&] { trsm_device<float>(side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, (cudaStream_t)s); });
}
int elb200_ztrsm(char side, char uplo, char trans, char diag, int64_t m, int64_t n, elb200_c64 alpha,
                 const elb200_c64* A, int64_t lda, elb200_c64* B, int64_t ldb, elb200_stream_t s) {
    return guarded([&] {
        trsm_device<c64_t>(side, uplo, trans, diag, m, n, mk(alpha.re, alpha.im), (const c64_t*)A, lda, (c64_t*)B,
                          ldb, (cudaStream_t)s);
    });
}
int elb200_ctrsm(char side, char uplo, char trans, char diag, int64_t m, int64_t n, elb200_c32 alpha,
                 const elb200_c32* A, int64_t lda, elb200_c32* B, int64_t ldb, elb200_stream_t s) {
    return guarded([&] {
        trsm_device<c32_t>(side, uplo, trans, diag, m, n, mk(alpha.re, alpha.im), (const c32_t*)A, lda, (c32_t*)B,
                          ldb, (cudaStream_t)s);
    });
}
}
