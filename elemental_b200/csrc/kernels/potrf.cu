// Cholesky factorisation of one (replicated) diagonal block, every scalar type.
//
// Replaces cholesky::LowerVariant3Unblocked / UpperVariant3Unblocked
// (reference src/lapack_like/factor/Cholesky/LowerVariant3.hpp:16-41,
// UpperVariant3.hpp:16-46: sqrt + Scal + Her per column, i.e. level-2 BLAS run
// redundantly on every rank) with ONE single-CTA kernel that keeps the working
// panel in shared memory:
//   for each 32-column block:   warp 0 factors the 32x32 diagonal block in
//   registers (shuffles, no memory traffic), every thread then solves one row
//   of the panel below against it (registers + broadcast smem reads), and the
//   whole CTA applies the rank-32 Hermitian update to the trailing triangle
//   from a shared-memory copy of the panel.
// Upper storage is handled as the conjugate-transposed view of the same
// lower algorithm (U = L^H), so the other triangle is never referenced, as in
// the reference.  Same arithmetic as the reference per column: alpha = sqrt(a_jj),
// scale by 1/alpha, rank-1 downdates; pivot test `a_jj <= 0` (also catches NaN).
//
// Blocks too large for one CTA's shared memory fall back to a host-blocked
// right-looking sweep over 256-column panels built from the same leaf plus
// trsm_device and the masked GEMM.
#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int JB = 32;
// 512 threads (128 regs each) for 4/8/8-byte scalars, 256 (255 regs) for Complex<double>
template <class T> struct potrf_threads { static constexpr int value = sizeof(T) >= 16 ? 256 : 512; };
constexpr int LDP = JB + 1;

template <class T, bool UPPER>
struct TriView {
    T* A;
    i64 lda;
    __device__ __forceinline__ T get(i64 i, i64 j) const {
        return UPPER ? scalar_traits<T>::conj(A[j + i * lda]) : A[i + j * lda];
    }
    __device__ __forceinline__ void set(i64 i, i64 j, T v) const {
        if (UPPER) A[j + i * lda] = scalar_traits<T>::conj(v);
        else A[i + j * lda] = v;
    }
};

template <class T>
__device__ __forceinline__ T shfl(T v, int src);
template <> __device__ __forceinline__ float shfl<float>(float v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ double shfl<double>(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <> __device__ __forceinline__ c32_t shfl<c32_t>(c32_t v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src));
}
template <> __device__ __forceinline__ c64_t shfl<c64_t>(c64_t v, int src) {
    return mk(__shfl_sync(0xffffffffu, v.re, src), __shfl_sync(0xffffffffu, v.im, src));
}

template <class T, bool UPPER>
__global__ void __launch_bounds__(potrf_threads<T>::value, 1) potrf_kernel(i64 n, T* Aptr, i64 lda, int* info, i64 col_offset) {
    constexpr int NT = potrf_threads<T>::value;
    typedef scalar_traits<T> st;
    typedef typename st::real R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* sD = (T*)smem_raw;                 // [JB][LDP] : L of the diagonal block
    T* sP = sD + JB * LDP;                // [rows][LDP]: solved panel
    __shared__ R rinv[JB];
    __shared__ int s_fail;
    const TriView<T, UPPER> V{Aptr, lda};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_fail = 0;
    __syncthreads();

    for (i64 c0 = 0; c0 < n; c0 += JB) {
        const int w = (int)((n - c0 < JB) ? (n - c0) : JB);
        // ---- (1) diagonal block: warp 0, lane i owns row i ----
        if (warp == 0) {
            T v[JB];
#pragma unroll
            for (int k = 0; k < JB; ++k) {
                T x = st::zero();
                if (lane < w && k <= lane) x = V.get(c0 + lane, c0 + k);
                if (lane >= w && k == lane) x = st::from_real(R(1));
                v[k] = x;
            }
            int fail = 0;
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                const R d = st::real_part(shfl<T>(v[j], j));
                if (j < w && !(d > R(0)) && fail == 0) fail = (int)(col_offset + c0 + j + 1);
                const R sq = sqrt(d);
                const R ri = R(1) / sq;
                if (lane == j) { v[j] = st::from_real(sq); }
                else if (lane > j) v[j] = v[j] * ri;
                if (lane == 0) rinv[j] = ri;
#pragma unroll
                for (int k = j + 1; k < JB; ++k) {
                    const T lkj = shfl<T>(v[j], k);
                    if (lane >= k) v[k] -= v[j] * st::conj(lkj);
                }
            }
            if (fail != 0) {
                if (lane == 0) { s_fail = fail; atomicCAS(info, 0, fail); }
            } else {
#pragma unroll
                for (int k = 0; k < JB; ++k) {
                    sD[lane * LDP + k] = (k <= lane) ? v[k] : st::zero();
                    if (lane < w && k <= lane) V.set(c0 + lane, c0 + k, v[k]);
                }
            }
        }
        __syncthreads();
        if (s_fail != 0) return;
        const i64 r0 = c0 + w;
        const i64 nt = n - r0;
        if (nt <= 0) break;
        // ---- (2) panel solve: one row per thread, x L^H = a ----
        const i64 T4 = (nt + 3) / 4;
        for (i64 t = tid; t < 4 * T4; t += NT) {
            T x[JB];
            if (t < nt) {
#pragma unroll
                for (int k = 0; k < JB; ++k) x[k] = (k < w) ? V.get(r0 + t, c0 + k) : st::zero();
#pragma unroll
                for (int k = 0; k < JB; ++k) {
                    T acc = x[k];
#pragma unroll
                    for (int q = 0; q < k; ++q) acc -= x[q] * st::conj(sD[k * LDP + q]);
                    x[k] = acc * rinv[k];
                }
#pragma unroll
                for (int k = 0; k < JB; ++k) {
                    if (k < w) V.set(r0 + t, c0 + k, x[k]);
                    sP[t * LDP + k] = x[k];
                }
            } else {
#pragma unroll
                for (int k = 0; k < JB; ++k) sP[t * LDP + k] = st::zero();
            }
        }
        __syncthreads();
        // ---- (3) trailing update: A22 -= P P^H on the lower triangle ----
        // thread tile = rows {ti + T4*a}, cols {tk + T4*b}; only a >= b can be lower.
        for (i64 idx = tid; idx < T4 * T4; idx += NT) {
            const i64 ti = idx % T4, tk = idx / T4;
            T acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = st::zero();
            for (int q = 0; q < w; ++q) {
                T pa[4], pb[4];
#pragma unroll
                for (int a = 0; a < 4; ++a) pa[a] = sP[(ti + T4 * a) * LDP + q];
#pragma unroll
                for (int b = 0; b < 4; ++b) pb[b] = st::conj(sP[(tk + T4 * b) * LDP + q]);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b <= a; ++b) acc[a][b] += pa[a] * pb[b];
            }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b <= a; ++b) {
                    const i64 i = ti + T4 * a, k = tk + T4 * b;
                    if (i < nt && k < nt && i >= k) {
                        const T old = V.get(r0 + i, r0 + k);
                        V.set(r0 + i, r0 + k, old - acc[a][b]);
                    }
                }
        }
        __syncthreads();
    }
}

template <class T>
size_t potrf_smem(i64 n) {
    const i64 rows = 4 * ((n + 3) / 4) + 4;
    return sizeof(T) * (size_t)(JB * LDP + rows * LDP);
}

template <class T, bool UPPER>
void launch_potrf(i64 n, T* A, i64 lda, int* info, i64 col_offset, cudaStream_t s) {
    const size_t smem = potrf_smem<T>(n);
    auto kern = potrf_kernel<T, UPPER>;
    static size_t configured = 0;
    if (smem > configured) {
        ELB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    kern<<<1, potrf_threads<T>::value, smem, s>>>(n, A, lda, info, col_offset);
    ELB_LAUNCH_CHECK();
}

}  // namespace

template <class T>
void potrf_device(char uplo_, i64 n, T* A, i64 lda, int* info, i64 col_offset, cudaStream_t s) {
    typedef scalar_traits<T> st;
    const char uplo = up(uplo_);
    if (uplo != 'L' && uplo != 'U') throw std::logic_error("potrf: uplo must be 'L' or 'U'");
    if (n < 0) throw std::logic_error("potrf: negative dimension");
    if (lda < (n > 1 ? n : 1)) throw std::logic_error("potrf: lda too small");
    if (!info) throw std::logic_error("potrf: info_dev must not be NULL");
    if (n == 0) return;
    const size_t limit = 200 * 1024;
    if (potrf_smem<T>(n) <= limit) {
        if (uplo == 'L') launch_potrf<T, false>(n, A, lda, info, col_offset, s);
        else launch_potrf<T, true>(n, A, lda, info, col_offset, s);
        return;
    }
    // host-blocked right-looking sweep (same structure as LowerVariant3Blocked(Matrix),
    // LowerVariant3.hpp:43-68) for blocks that do not fit one CTA
    const i64 nb = 256;
    const T one = st::from_real(1), minus_one = st::from_real(-1);
    for (i64 k = 0; k < n; k += nb) {
        const i64 b = (n - k < nb) ? (n - k) : nb;
        T* A11 = A + k + k * lda;
        if (uplo == 'L') launch_potrf<T, false>(b, A11, lda, info, col_offset + k, s);
        else launch_potrf<T, true>(b, A11, lda, info, col_offset + k, s);
        const i64 rest = n - k - b;
        if (rest <= 0) break;
        if (uplo == 'L') {
            T* A21 = A + (k + b) + k * lda;
            T* A22 = A + (k + b) + (k + b) * lda;
            trsm_device<T>('R', 'L', 'C', 'N', rest, b, one, A11, lda, A21, lda, s);
            gemm_device<T>(1, 'N', 'C', rest, rest, b, minus_one, A21, lda, A21, lda, one, A22, lda, 0, 1, 0, 1, s);
        } else {
            T* A12 = A + k + (k + b) * lda;
            T* A22 = A + (k + b) + (k + b) * lda;
            trsm_device<T>('L', 'U', 'C', 'N', b, rest, one, A11, lda, A12, lda, s);
            gemm_device<T>(2, 'C', 'N', rest, rest, b, minus_one, A12, lda, A12, lda, one, A22, lda, 0, 1, 0, 1, s);
        }
    }
}

template void potrf_device<float>(char, i64, float*, i64, int*, i64, cudaStream_t);
template void potrf_device<double>(char, i64, double*, i64, int*, i64, cudaStream_t);
template void potrf_device<c32_t>(char, i64, c32_t*, i64, int*, i64, cudaStream_t);
template void potrf_device<c64_t>(char, i64, c64_t*, i64, int*, i64, cudaStream_t);

}  // namespace elb200

extern "C" {
using namespace elb200;
int elb200_dpotrf(char uplo, int64_t n, double* A, int64_t lda, int* info_dev, elb200_stream_t s) {
    return guarded([&] { potrf_device<double>(uplo, n, A, lda, info_dev, 0, (cudaStream_t)s); });
}
int elb200_spotrf(char uplo, int64_t n, float* A, int64_t lda, int* info_dev, elb200_stream_t s) {
    return guarded([&] { potrf_device<float>(uplo, n, A, lda, info_dev, 0, (cudaStream_t)s); });
}
int elb200_zpotrf(char uplo, int64_t n, elb200_c64* A, int64_t lda, int* info_dev, elb200_stream_t s) {
    return guarded([&] { potrf_device<c64_t>(uplo, n, (c64_t*)A, lda, info_dev, 0, (cudaStream_t)s); });
}
int elb200_cpotrf(char uplo, int64_t n, elb200_c32* A, int64_t lda, int* info_dev, elb200_stream_t s) {
    return guarded([&] { potrf_device<c32_t>(uplo, n, (c32_t*)A, lda, info_dev, 0, (cudaStream_t)s); });
}
}
