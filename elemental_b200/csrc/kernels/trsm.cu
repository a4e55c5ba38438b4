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
#include <type_traits>

#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int TB = 32;
int g_trsm_flags = 0;  // bit 0: never use the fused slab kernel (debug / A-B timing)

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


// ---------------------------------------------------------------------------------------------
// Fused right-side solve for the Cholesky panel shape (double): X op(A) = alpha B with op(A) UPPER
// triangular of order n <= 256 (A lower and transposed: the A21 L11^-T of the right-looking
// factorisation) and B tall.  The rows of a right-side solve are independent, so ONE launch does the
// whole block substitution: each CTA keeps a 64-row slab of B in shared memory, and per 32-column
// block applies the inverted diagonal block and downdates the remaining columns of its slab, both on
// the FP64 tensor pipe (DMMA.8x8x4 fragments read from shared memory with pitches = 4 mod 16
// doubles: conflict-free).  The generic path above needs 2 launches per 32-column block; for the
// 256-wide panel of the hot path that was 17 dependent launches (~165 us whatever the height).
// ---------------------------------------------------------------------------------------------
constexpr int SLAB_ROWS = 64, SLAB_MAXN = 256, SLAB_THREADS = 512;
constexpr int SLAB_PM = TB + 4;  // pitch of the 32-wide operand panels

__device__ __forceinline__ void dmma884_t(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

// tr: op(A) = A^T (A stored lower) ; else op(A) = A (A stored upper).  inv: TB x TB inverses of the
// stored diagonal blocks (trtri_diag_kernel).
__global__ void __launch_bounds__(SLAB_THREADS, 1) trsm_right_slab_kernel(int tr, i64 m, i64 n, double alpha,
                                                                          const double* __restrict__ A, i64 lda,
                                                                          const double* __restrict__ inv, double* B,
                                                                          i64 ldb) {
    extern __shared__ __align__(16) unsigned char slab_smem[];
    const int nblk = (int)((n + TB - 1) / TB);
    const int ncol = nblk * TB;            // padded width
    const int P = ncol + 4;                // slab pitch (= 4 mod 16 doubles since ncol % 32 == 0)
    double* S = (double*)slab_smem;        // [SLAB_ROWS][P]
    double* sI = S + SLAB_ROWS * P;        // [TB][SLAB_PM]      sI[nn][kk] = inv(M_kk)(kk, nn)
    double* sM = sI + TB * SLAB_PM;        // [ncol - TB][SLAB_PM]  sM[nn][kk] = M(k0 + kk, k0 + TB + nn)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, tq = lane & 3;
    const i64 i0 = (i64)blockIdx.x * SLAB_ROWS;

    // slab <- alpha * B (zero beyond m and n)
    for (int e = tid; e < SLAB_ROWS * ncol; e += SLAB_THREADS) {
        const int r = e % SLAB_ROWS, c = e / SLAB_ROWS;
        double v = 0.0;
        if (i0 + r < m && c < n) v = alpha * B[(i0 + r) + (i64)c * ldb];
        S[r * P + c] = v;
    }
    const int rf = warp & 7, hf = warp >> 3;   // row fragment (8 rows) and half (0/1) of this warp
    const double* Srow = S + (rf * 8 + g) * P;
    for (int kb = 0; kb < nblk; ++kb) {
        const int k0 = kb * TB, c0 = k0 + TB, rest = ncol - c0;
        // operand panels of this block step
        for (int e = tid; e < TB * TB; e += SLAB_THREADS) {
            const int nn = e & 31, kk = e >> 5;
            const double* blk = inv + (i64)kb * TB * TB;
            sI[nn * SLAB_PM + kk] = tr ? blk[nn + kk * TB] : blk[kk + nn * TB];
        }
        for (int e = tid; e < rest * TB; e += SLAB_THREADS) {
            const int nn = e % rest, kk = e / rest;
            const i64 row = k0 + kk, col = c0 + nn;   // M(row, col), row < col
            double v = 0.0;
            if (row < n && col < n) v = tr ? A[col + row * lda] : A[row + col * lda];
            sM[nn * SLAB_PM + kk] = v;
        }
        __syncthreads();
        // X_k = S_k inv(M_kk): warp -> 8 rows x 16 columns
        double x[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
        for (int ks = 0; ks < TB / 4; ++ks) {
            const double a = Srow[k0 + 4 * ks + tq];
#pragma unroll
            for (int f = 0; f < 2; ++f) dmma884_t(x[f], a, sI[((hf * 2 + f) * 8 + g) * SLAB_PM + 4 * ks + tq]);
        }
        __syncthreads();   // every warp has read S_k
#pragma unroll
        for (int f = 0; f < 2; ++f)
#pragma unroll
            for (int e = 0; e < 2; ++e) S[(rf * 8 + g) * P + k0 + (hf * 2 + f) * 8 + 2 * tq + e] = x[f][e];
        __syncthreads();
        // S_rest -= X_k M(k, rest): groups of 4 column fragments per warp
        const int ngroups = (rest / 8 + 3) / 4;
        for (int gi = hf; gi < ngroups; gi += 2) {
            double acc[4][2];
#pragma unroll
            for (int f = 0; f < 4; ++f) acc[f][0] = acc[f][1] = 0.0;
            const int cf0 = gi * 4;
#pragma unroll
            for (int ks = 0; ks < TB / 4; ++ks) {
                const double a = Srow[k0 + 4 * ks + tq];
#pragma unroll
                for (int f = 0; f < 4; ++f)
                    if ((cf0 + f) * 8 < rest) dmma884_t(acc[f], a, sM[((cf0 + f) * 8 + g) * SLAB_PM + 4 * ks + tq]);
            }
#pragma unroll
            for (int f = 0; f < 4; ++f)
                if ((cf0 + f) * 8 < rest) {
                    double* d = S + (rf * 8 + g) * P + c0 + (cf0 + f) * 8 + 2 * tq;
                    d[0] -= acc[f][0];
                    d[1] -= acc[f][1];
                }
        }
        __syncthreads();
    }
    for (int e = tid; e < SLAB_ROWS * ncol; e += SLAB_THREADS) {
        const int r = e % SLAB_ROWS, c = e / SLAB_ROWS;
        if (i0 + r < m && c < n) B[(i0 + r) + (i64)c * ldb] = S[r * P + c];
    }
}

size_t slab_smem_bytes(i64 n) {
    const i64 ncol = ceil_div(n, TB) * TB;
    return sizeof(double) * (size_t)(SLAB_ROWS * (ncol + 4) + TB * SLAB_PM + (ncol - TB) * SLAB_PM);
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
    if constexpr (std::is_same<T, double>::value) {
        // the Cholesky panel shape: one fused launch (see trsm_right_slab_kernel)
        if (!left && !eff_lower && na <= SLAB_MAXN && !(g_trsm_flags & 1)) {
            const size_t smem = slab_smem_bytes(na);
            static size_t configured = 0;
            if (smem > configured) {
                ELB_CUDA(cudaFuncSetAttribute(trsm_right_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured = smem;
            }
            trsm_right_slab_kernel<<<(unsigned)ceil_div(m, (i64)SLAB_ROWS), SLAB_THREADS, smem, s>>>(
                tr ? 1 : 0, m, n, 1.0, A, lda, inv, B, ldb);
            ELB_LAUNCH_CHECK();
            scratch_free(inv, s);
            return;
        }
    }
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
    // alpha == 0: B := 0 without referencing A (the BLAS definition; 0 * B would keep NaN / Inf)
    if (st::is_zero(alpha)) {
        ELB_CUDA(cudaMemset2DAsync(B, sizeof(T) * (size_t)ldb, 0, sizeof(T) * (size_t)m, (size_t)n, s));
        return;
    }
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
void elb200_trsm_set_debug_flags(int f) { g_trsm_flags = f; }
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
