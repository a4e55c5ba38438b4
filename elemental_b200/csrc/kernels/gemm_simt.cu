// Generic SIMT GEMM / TRRK for every scalar type (exact FFMA/DFMA arithmetic, all
// N/T/C combinations, conjugation in the loader) plus the typed dispatch
// gemm_device<T>: double -> DMMA kernel (gemm_f64.cu), Complex<double> -> DMMA
// kernel (gemm_c64.cu), float/Complex<float> -> this kernel (float's exact-FFMA
// path of BASELINE.json config 5; the 3xTF32 tcgen05 path is gemm_tf32.cu).
//
// CTA tile 64x64x16, 256 threads, 4x4 register micro-tile per thread, operands
// staged through shared memory as s[k][64+pad] so that both global reads
// (coalesced along the contiguous dimension) and fragment reads (consecutive
// threads -> consecutive addresses) are conflict-free.
#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int TM = 64, TN = 64, TK = 16;

struct SimtArgs {
    i64 m, n, k, lda, ldb, ldc;
    i64 gi0, gis, gj0, gjs;
    int ta, tb;  // 0 N, 1 T, 2 C
    int mode;
    int realDiag;  // HERK: force Im(c_ii) = 0
};

template <class T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtArgs p, T alpha, T beta,
                                                        const T* __restrict__ A,
                                                        const T* __restrict__ B, T* __restrict__ C) {
    __shared__ T sA[TK][TM + 1];
    __shared__ T sB[TK][TN + 1];
    const i64 m0 = (i64)blockIdx.x * TM, n0 = (i64)blockIdx.y * TN;
    if (p.mode != 0) {
        const i64 mlast = (m0 + TM - 1 < p.m - 1) ? (m0 + TM - 1) : (p.m - 1);
        const i64 nlast = (n0 + TN - 1 < p.n - 1) ? (n0 + TN - 1) : (p.n - 1);
        if (p.mode == 1 && p.gi0 + mlast * p.gis < p.gj0 + n0 * p.gjs) return;
        if (p.mode == 2 && p.gi0 + m0 * p.gis > p.gj0 + nlast * p.gjs) return;
    }
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    T acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = scalar_traits<T>::zero();

    for (i64 k0 = 0; k0 < p.k; k0 += TK) {
#pragma unroll
        for (int q = 0; q < (TM * TK) / 256; ++q) {
            const int e = tid + q * 256;
            int i, kk;
            if (p.ta == 0) { i = e % TM; kk = e / TM; } else { kk = e % TK; i = e / TK; }
            T v = scalar_traits<T>::zero();
            if (m0 + i < p.m && k0 + kk < p.k) {
                v = (p.ta == 0) ? A[(m0 + i) + (k0 + kk) * p.lda] : A[(k0 + kk) + (m0 + i) * p.lda];
                if (p.ta == 2) v = scalar_traits<T>::conj(v);
            }
            sA[kk][i] = v;
        }
#pragma unroll
        for (int q = 0; q < (TN * TK) / 256; ++q) {
            const int e = tid + q * 256;
            int j, kk;
            if (p.tb == 0) { kk = e % TK; j = e / TK; } else { j = e % TN; kk = e / TN; }
            T v = scalar_traits<T>::zero();
            if (n0 + j < p.n && k0 + kk < p.k) {
                v = (p.tb == 0) ? B[(k0 + kk) + (n0 + j) * p.ldb] : B[(n0 + j) + (k0 + kk) * p.ldb];
                if (p.tb == 2) v = scalar_traits<T>::conj(v);
            }
            sB[kk][j] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            T a[4], b[4];
#pragma unroll
            for (int x = 0; x < 4; ++x) a[x] = sA[kk][tx + 16 * x];
#pragma unroll
            for (int y = 0; y < 4; ++y) b[y] = sB[kk][ty + 16 * y];
#pragma unroll
            for (int x = 0; x < 4; ++x)
#pragma unroll
                for (int y = 0; y < 4; ++y) acc[x][y] += a[x] * b[y];
        }
        __syncthreads();
    }
    const bool has_beta = !scalar_traits<T>::is_zero(beta);
#pragma unroll
    for (int y = 0; y < 4; ++y) {
        const i64 col = n0 + ty + 16 * y;
        if (col >= p.n) continue;
        const i64 gj = p.gj0 + col * p.gjs;
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            const i64 row = m0 + tx + 16 * x;
            if (row >= p.m) continue;
            const i64 gi = p.gi0 + row * p.gis;
            if (p.mode == 1 && gi < gj) continue;
            if (p.mode == 2 && gi > gj) continue;
            T* c = C + row + col * p.ldc;
            T v = alpha * acc[x][y];
            if (has_beta) v += beta * (*c);
            if (p.realDiag && gi == gj) v = scalar_traits<T>::from_real(scalar_traits<T>::real_part(v));
            *c = v;
        }
    }
}

int trans_code(char c, const char* what) {
    c = up(c);
    if (c == 'N') return 0;
    if (c == 'T') return 1;
    if (c == 'C') return 2;
    throw std::logic_error(std::string("invalid orientation for ") + what);
}

}  // namespace

template <class T>
void gemm_simt_device(int mode, char transA, char transB, i64 m, i64 n, i64 k, T alpha, const T* A,
                      i64 lda, const T* B, i64 ldb, T beta, T* C, i64 ldc, i64 gi0, i64 gis,
                      i64 gj0, i64 gjs, cudaStream_t s, bool realDiag) {
    if (m < 0 || n < 0 || k < 0) throw std::logic_error("gemm: negative dimension");
    SimtArgs p;
    p.realDiag = realDiag ? 1 : 0;
    p.ta = trans_code(transA, "A");
    p.tb = trans_code(transB, "B");
    if (!scalar_traits<T>::is_complex) {
        if (p.ta == 2) p.ta = 1;
        if (p.tb == 2) p.tb = 1;
    }
    if (m == 0 || n == 0) return;
    p.m = m; p.n = n; p.k = scalar_traits<T>::is_zero(alpha) ? 0 : k;
    p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.gi0 = gi0; p.gis = gis; p.gj0 = gj0; p.gjs = gjs;
    p.mode = mode;
    dim3 grid((unsigned)ceil_div(m, TM), (unsigned)ceil_div(n, TN));
    if (grid.y > 65535) throw std::logic_error("gemm_simt: n too large for this kernel");
    gemm_simt_kernel<T><<<grid, 256, 0, s>>>(p, alpha, beta, A, B, C);
    ELB_LAUNCH_CHECK();
}

#define ELB_SIMT_INST(T)                                                                                   \
    template void gemm_simt_device<T>(int, char, char, i64, i64, i64, T, const T*, i64, const T*, i64, T, T*, \
                                      i64, i64, i64, i64, i64, cudaStream_t, bool);
ELB_SIMT_INST(float)
ELB_SIMT_INST(double)
ELB_SIMT_INST(c32_t)
ELB_SIMT_INST(c64_t)

// ---- typed dispatch -------------------------------------------------------
template <>
void gemm_device<double>(int mode, char ta, char tb, i64 m, i64 n, i64 k, double alpha, const double* A,
                         i64 lda, const double* B, i64 ldb, double beta, double* C, i64 ldc, i64 gi0,
                         i64 gis, i64 gj0, i64 gjs, cudaStream_t s) {
    dgemm_device(mode, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs, s);
}
template <>
void gemm_device<c64_t>(int mode, char ta, char tb, i64 m, i64 n, i64 k, c64_t alpha, const c64_t* A,
                        i64 lda, const c64_t* B, i64 ldb, c64_t beta, c64_t* C, i64 ldc, i64 gi0,
                        i64 gis, i64 gj0, i64 gjs, cudaStream_t s) {
    zgemm_device(mode, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs, s);
}
bool sgemm_ffma_device(int mode, bool ta, bool tb, i64 m, i64 n, i64 k, float alpha, const float* A, i64 lda,
                       const float* B, i64 ldb, float beta, float* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                       cudaStream_t s);
int g_sgemm_ffma_path = 0;   // 0 automatic, 1 always the generic SIMT kernel (elb200_sgemm_set_ffma_path)
int g_sgemm_ffma_last = 0;   // 1 generic SIMT kernel, 2 register-tiled float kernel
template <>
void gemm_device<float>(int mode, char ta, char tb, i64 m, i64 n, i64 k, float alpha, const float* A,
                        i64 lda, const float* B, i64 ldb, float beta, float* C, i64 ldc, i64 gi0,
                        i64 gis, i64 gj0, i64 gjs, cudaStream_t s) {
    // float: exact FFMA (default) or, when elb200_sgemm_set_mode(1) is in force and the operands meet
    // TMA's alignment rules, the 3xTF32 tcgen05 kernel (full GEMMs only; the masked TRRK form stays SIMT)
    if (mode == 0 && sgemm_mode() == 1 && k > 0 &&
        sgemm_3xtf32_device(ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, s))
        return;
    sgemm_note_simt();
    // exact FFMA: the register-tiled float kernel (gemm_f32_ffma.cu) whenever the operands allow 16-byte loads,
    // the generic SIMT kernel of this file otherwise
    {
        const char ua = up(ta), ub = up(tb);
        if (g_sgemm_ffma_path != 1 && k > 0 &&
            sgemm_ffma_device(mode, ua != 'N', ub != 'N', m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs, s)) {
            g_sgemm_ffma_last = 2;
            return;
        }
    }
    g_sgemm_ffma_last = 1;
    gemm_simt_device<float>(mode, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs, s);
}
template <>
void gemm_device<c32_t>(int mode, char ta, char tb, i64 m, i64 n, i64 k, c32_t alpha, const c32_t* A,
                        i64 lda, const c32_t* B, i64 ldb, c32_t beta, c32_t* C, i64 ldc, i64 gi0,
                        i64 gis, i64 gj0, i64 gjs, cudaStream_t s) {
    gemm_simt_device<c32_t>(mode, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs, s);
}

namespace {
int uplo_mode(char uplo, const char* what) {
    const char u = up(uplo);
    if (u == 'L') return 1;
    if (u == 'U') return 2;
    throw std::logic_error(std::string(what) + ": uplo must be 'L' or 'U'");
}
template <class T, class E>
inline T as(E z) { return mk(z.re, z.im); }
}  // namespace
}  // namespace elb200

extern "C" {
using namespace elb200;

int elb200_sgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha, const float* A,
                 int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc,
                 elb200_stream_t s) {
    return guarded([&] {
        gemm_device<float>(0, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, 1, 0, 1, (cudaStream_t)s);
    });
}
int elb200_zgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, elb200_c64 alpha,
                 const elb200_c64* A, int64_t lda, const elb200_c64* B, int64_t ldb, elb200_c64 beta,
                 elb200_c64* C, int64_t ldc, elb200_stream_t s) {
    return guarded([&] {
        gemm_device<c64_t>(0, ta, tb, m, n, k, mk(alpha.re, alpha.im), (const c64_t*)A, lda, (const c64_t*)B,
                           ldb, mk(beta.re, beta.im), (c64_t*)C, ldc, 0, 1, 0, 1, (cudaStream_t)s);
    });
}
int elb200_cgemm(char ta, char tb, int64_t m, int64_t n, int64_t k, elb200_c32 alpha,
                 const elb200_c32* A, int64_t lda, const elb200_c32* B, int64_t ldb, elb200_c32 beta,
                 elb200_c32* C, int64_t ldc, elb200_stream_t s) {
    return guarded([&] {
        gemm_device<c32_t>(0, ta, tb, m, n, k, mk(alpha.re, alpha.im), (const c32_t*)A, lda, (const c32_t*)B,
                           ldb, mk(beta.re, beta.im), (c32_t*)C, ldc, 0, 1, 0, 1, (cudaStream_t)s);
    });
}

int elb200_strrk(char uplo, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha,
                 const float* A, int64_t lda, const float* B, int64_t ldb, float beta, float* C,
                 int64_t ldc, int64_t rs, int64_t rst, int64_t cs, int64_t cst, elb200_stream_t s) {
    return guarded([&] {
        gemm_device<float>(uplo_mode(uplo, "strrk"), ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, rs,
                           rst, cs, cst, (cudaStream_t)s);
    });
}
int elb200_ztrrk(char uplo, char ta, char tb, int64_t m, int64_t n, int64_t k, elb200_c64 alpha,
                 const elb200_c64* A, int64_t lda, const elb200_c64* B, int64_t ldb, elb200_c64 beta,
                 elb200_c64* C, int64_t ldc, int64_t rs, int64_t rst, int64_t cs, int64_t cst,
                 elb200_stream_t s) {
    return guarded([&] {
        gemm_device<c64_t>(uplo_mode(uplo, "ztrrk"), ta, tb, m, n, k, mk(alpha.re, alpha.im), (const c64_t*)A,
                           lda, (const c64_t*)B, ldb, mk(beta.re, beta.im), (c64_t*)C, ldc, rs, rst, cs, cst,
                           (cudaStream_t)s);
    });
}
int elb200_ctrrk(char uplo, char ta, char tb, int64_t m, int64_t n, int64_t k, elb200_c32 alpha,
                 const elb200_c32* A, int64_t lda, const elb200_c32* B, int64_t ldb, elb200_c32 beta,
                 elb200_c32* C, int64_t ldc, int64_t rs, int64_t rst, int64_t cs, int64_t cst,
                 elb200_stream_t s) {
    return guarded([&] {
        gemm_device<c32_t>(uplo_mode(uplo, "ctrrk"), ta, tb, m, n, k, mk(alpha.re, alpha.im), (const c32_t*)A,
                           lda, (const c32_t*)B, ldb, mk(beta.re, beta.im), (c32_t*)C, ldc, rs, rst, cs, cst,
                           (cudaStream_t)s);
    });
}

// SYRK / HERK: 'N' -> A op(A)^T with A n x k; otherwise op(A)^T A with A k x n.
// HERK keeps the diagonal real by construction (a_i . conj(a_i)); alpha, beta real.
int elb200_ssyrk(char uplo, char trans, int64_t n, int64_t k, float alpha, const float* A,
                 int64_t lda, float beta, float* C, int64_t ldc, elb200_stream_t s) {
    return guarded([&] {
        const bool tr = up(trans) != 'N';
        gemm_device<float>(uplo_mode(uplo, "ssyrk"), tr ? 'T' : 'N', tr ? 'N' : 'T', n, n, k, alpha, A, lda, A,
                           lda, beta, C, ldc, 0, 1, 0, 1, (cudaStream_t)s);
    });
}
int elb200_zherk(char uplo, char trans, int64_t n, int64_t k, double alpha, const elb200_c64* A,
                 int64_t lda, double beta, elb200_c64* C, int64_t ldc, elb200_stream_t s) {
    return guarded([&] {
        const bool tr = up(trans) != 'N';
        zgemm_device_ex(uplo_mode(uplo, "zherk"), tr ? 'C' : 'N', tr ? 'N' : 'C', n, n, k, mk(alpha, 0.0),
                        (const c64_t*)A, lda, (const c64_t*)A, lda, mk(beta, 0.0), (c64_t*)C, ldc, 0, 1, 0, 1, true,
                        (cudaStream_t)s);
    });
}
int elb200_cherk(char uplo, char trans, int64_t n, int64_t k, float alpha, const elb200_c32* A,
                 int64_t lda, float beta, elb200_c32* C, int64_t ldc, elb200_stream_t s) {
    return guarded([&] {
        const bool tr = up(trans) != 'N';
        gemm_simt_device<c32_t>(uplo_mode(uplo, "cherk"), tr ? 'C' : 'N', tr ? 'N' : 'C', n, n, k, mk(alpha, 0.f),
                                (const c32_t*)A, lda, (const c32_t*)A, lda, mk(beta, 0.f), (c32_t*)C, ldc, 0, 1, 0, 1,
                                (cudaStream_t)s, true);
    });
}
int elb200_zsyrk(char uplo, char trans, int64_t n, int64_t k, elb200_c64 alpha, const elb200_c64* A,
                 int64_t lda, elb200_c64 beta, elb200_c64* C, int64_t ldc, elb200_stream_t s) {
    return guarded([&] {
        const bool tr = up(trans) != 'N';
        gemm_device<c64_t>(uplo_mode(uplo, "zsyrk"), tr ? 'T' : 'N', tr ? 'N' : 'T', n, n, k,
                           mk(alpha.re, alpha.im), (const c64_t*)A, lda, (const c64_t*)A, lda,
                           mk(beta.re, beta.im), (c64_t*)C, ldc, 0, 1, 0, 1, (cudaStream_t)s);
    });
}
int elb200_csyrk(char uplo, char trans, int64_t n, int64_t k, elb200_c32 alpha, const elb200_c32* A,
                 int64_t lda, elb200_c32 beta, elb200_c32* C, int64_t ldc, elb200_stream_t s) {
    return guarded([&] {
        const bool tr = up(trans) != 'N';
        gemm_device<c32_t>(uplo_mode(uplo, "csyrk"), tr ? 'T' : 'N', tr ? 'N' : 'T', n, n, k,
                           mk(alpha.re, alpha.im), (const c32_t*)A, lda, (const c32_t*)A, lda,
                           mk(beta.re, beta.im), (c32_t*)C, ldc, 0, 1, 0, 1, (cudaStream_t)s);
    });
}

}  // extern "C"

extern "C" {
// exact-FFMA float path: 0 automatic (register-tiled kernel when the operands allow 16-byte loads), 1 always the
// generic SIMT kernel; which one the last exact-FFMA call used: 1 generic, 2 register-tiled
void elb200_sgemm_set_ffma_path(int p) { elb200::g_sgemm_ffma_path = p; }
int elb200_sgemm_ffma_last_kernel(void) { return elb200::g_sgemm_ffma_last; }
}
