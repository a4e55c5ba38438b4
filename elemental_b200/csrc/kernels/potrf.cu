// Cholesky factorisation of one (replicated) diagonal block, every scalar type.
//
// Replaces cholesky::LowerVariant3Unblocked / UpperVariant3Unblocked
// (reference src/lapack_like/factor/Cholesky/LowerVariant3.hpp:16-41,
// UpperVariant3.hpp:16-46: sqrt + Scal + Her per column, i.e. level-2 BLAS run
// redundantly on every rank) with ONE single-CTA kernel that keeps the working
// panel in shared memory:
//   for each 32-column block:   ALL threads eliminate the 32x32 diagonal block in
//   shared memory with one barrier per column (trailing entries use the unscaled
//   column and 1/d; the scaled column goes to a second array), every thread then
//   solves one row of the panel below against it (a 32-step dependency chain), and
//   the whole CTA applies the rank-32 Hermitian update to the trailing triangle
//   from a shared-memory copy of the panel (double: 32x32 blocks per warp on the
//   FP64 tensor pipe, DMMA.8x8x4 -- a quarter of the shared-memory loads of the
//   scalar form, which is what bounds it).  n = 256 double: 0.18 ms.  (Round 1's
//   first version factored the diagonal block in one warp's registers with shuffles:
//   2000 clocks per column, instruction-latency bound.)
// Upper storage is handled as the conjugate-transposed view of the same
// lower algorithm (U = L^H), so the other triangle is never referenced, as in
// the reference.  Same arithmetic as the reference per column: alpha = sqrt(a_jj),
// scale by 1/alpha, rank-1 downdates; pivot test `a_jj <= 0` (also catches NaN).
//
// Blocks too large for one CTA's shared memory fall back to a host-blocked
// right-looking sweep over 256-column panels built from the same leaf plus
// trsm_device and the masked GEMM.
#include <type_traits>

#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int JB = 32;
// 512 threads (128 regs each) for 4/8/8-byte scalars, 256 (255 regs) for Complex<double>
template <class T> struct potrf_threads { static constexpr int value = sizeof(T) >= 16 ? 256 : 512; };
// panel pitch: odd for the scalar update (row-per-lane reads), = 4 mod 16 doubles for the DMMA
// fragment reads (8 rows x 4 k per half-warp land on 32 distinct banks)
template <class T> struct potrf_ldp { static constexpr int value = JB + 1; };
template <> struct potrf_ldp<double> { static constexpr int value = JB + 4; };

__device__ __forceinline__ float rsq(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsq(double x) { return rsqrt(x); }

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

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

// per-phase clock counts of the last launches (thread 0; diag, solve, update, launches): a debugging
// aid read by elb200_potrf_phase_clocks.  Accumulated in registers and written once at the end of the
// kernel, and only while the profile is switched on (elb200_potrf_phase_clocks(out, reset = 2)).
__device__ unsigned long long g_potrf_clk[4];
bool g_potrf_prof = false;

template <class T, bool UPPER>
__global__ void __launch_bounds__(potrf_threads<T>::value, 1) potrf_kernel(i64 n, T* Aptr, i64 lda, int* info, i64 col_offset, int prof) {
    constexpr int NT = potrf_threads<T>::value;
    constexpr int LDP = potrf_ldp<T>::value;
    typedef scalar_traits<T> st;
    typedef typename st::real R;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LDD = JB + 1;           // odd pitch: column reads of the 32 x 32 block are conflict-free
    T* sD = (T*)smem_raw;                 // [JB][LDD] : L of the diagonal block
    T* sW = sD + JB * LDD;                // [JB][LDD] : working copy being eliminated
    T* sP = sW + JB * LDD;                // [rows][LDP]: solved panel
    __shared__ R rinv[JB];
    const TriView<T, UPPER> V{Aptr, lda};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    long long tclk = prof ? clock64() : 0;
    unsigned long long clk[3] = {0, 0, 0};
    auto lap = [&](int slot) {
        if (prof && tid == 0) { const long long now = clock64(); clk[slot] += (unsigned long long)(now - tclk); tclk = now; }
    };
    auto flush = [&]() {
        if (prof && tid == 0) {
            g_potrf_clk[0] += clk[0]; g_potrf_clk[1] += clk[1]; g_potrf_clk[2] += clk[2]; g_potrf_clk[3] += 1;
        }
    };
    for (i64 c0 = 0; c0 < n; c0 += JB) {
        const int w = (int)((n - c0 < JB) ? (n - c0) : JB);
        // ---- (1) diagonal block in shared memory, all threads, ONE barrier per column ----
        // (a single warp holding the block in registers is bound by instruction latency: ~2000 clocks
        // per column were measured.)  Column j: every thread reads d = a_jj and forms 1/sqrt(d) itself;
        // the trailing entries take  a_ik -= a_ij conj(a_kj) / d  from the UNSCALED column j, while the
        // scaled column l_ij = a_ij / sqrt(d) goes to a second array, so nothing read in step j is
        // written in step j.
        for (int e = tid; e < JB * JB; e += NT) {
            const int i = e & 31, k = e >> 5;
            T x = st::zero();
            if (i < w && k <= i) x = V.get(c0 + i, c0 + k);
            if (i >= w && k == i) x = st::from_real(R(1));
            sW[i * LDD + k] = x;
            sD[i * LDD + k] = st::zero();
        }
        __syncthreads();
        int fail = 0;
#pragma unroll 1
        for (int j = 0; j < JB; ++j) {
            const R d = st::real_part(sW[j * LDD + j]);
            if (j < w && !(d > R(0))) { fail = (int)(col_offset + c0 + j + 1); break; }  // uniform over the CTA
            const R ri = rsq(d);
            R sq = d * ri;
            sq = fma(R(0.5) * ri, fma(-sq, sq, d), sq);
            const R rid = ri * ri;
            if (tid == 0) rinv[j] = ri;
            for (int e = tid; e < JB * JB; e += NT) {
                const int i = e >> 5, k = e & 31;
                if (k > j && i >= k) {
                    sW[i * LDD + k] -= (sW[i * LDD + j] * st::conj(sW[k * LDD + j])) * rid;
                } else if (k == j && i >= j) {
                    sD[i * LDD + j] = (i == j) ? st::from_real(sq) : sW[i * LDD + j] * ri;
                }
            }
            __syncthreads();
        }
        if (fail != 0) {
            if (tid == 0) atomicCAS(info, 0, fail);
            return;
        }
        for (int e = tid; e < JB * JB; e += NT) {
            const int i = e & 31, k = e >> 5;
            if (i < w && k <= i) V.set(c0 + i, c0 + k, sD[i * LDD + k]);
        }
        lap(0);
        const i64 r0 = c0 + w;
        const i64 nt = n - r0;
        if (nt <= 0) break;
        // ---- (2) panel solve: one row per thread, x L^H = a ----
        const i64 T4 = (nt + 3) / 4;
        const i64 rowsFill = std::is_same<T, double>::value ? (nt + 31) / 32 * 32 : 4 * T4;
        for (i64 t = tid; t < rowsFill; t += NT) {
            T x[JB];
            if (t < nt) {
#pragma unroll
                for (int k = 0; k < JB; ++k) x[k] = (k < w) ? V.get(r0 + t, c0 + k) : st::zero();
                // right-looking substitution: x_q is final after its scaling, the 31 - q downdates that
                // follow are independent of each other (a 32-step dependency chain instead of 496)
#pragma unroll
                for (int q = 0; q < JB; ++q) {
                    x[q] = x[q] * rinv[q];
#pragma unroll
                    for (int k = q + 1; k < JB; ++k) x[k] -= x[q] * st::conj(sD[k * LDD + q]);
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
        lap(1);
        // ---- (3) trailing update: A22 -= P P^H on the lower triangle ----
        if constexpr (std::is_same<T, double>::value) {
            // one 32 x 32 block (bi >= bk) of the trailing matrix per warp: 4 x 4 DMMA fragments, 8 k-steps
            const int nbk = (int)((nt + 31) / 32), nblk = nbk * (nbk + 1) / 2;
            const int g = lane >> 2, tq = lane & 3;
            const double* P = reinterpret_cast<const double*>(sP);
            for (int b = warp; b < nblk; b += NT / 32) {
                int bi = (int)((sqrtf(8.f * (float)b + 1.f) - 1.f) * 0.5f);
                while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
                while (bi * (bi + 1) / 2 > b) --bi;
                const int bk = b - bi * (bi + 1) / 2;
                double acc[4][4][2];
#pragma unroll
                for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                    for (int fk = 0; fk < 4; ++fk) acc[fi][fk][0] = acc[fi][fk][1] = 0.0;
                const double* Pa = P + (bi * 32 + g) * LDP + tq;
                const double* Pb = P + (bk * 32 + g) * LDP + tq;
#pragma unroll
                for (int q0 = 0; q0 < JB; q0 += 4) {
                    double fa[4], fb[4];
#pragma unroll
                    for (int f = 0; f < 4; ++f) { fa[f] = Pa[f * 8 * LDP + q0]; fb[f] = Pb[f * 8 * LDP + q0]; }
#pragma unroll
                    for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                        for (int fk = 0; fk < 4; ++fk)
                            if (bi != bk || fk <= fi) dmma884(acc[fi][fk], fa[fi], fb[fk]);
                }
                // all loads of the read-modify-write first (the stores may alias them for the compiler)
#pragma unroll
                for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                    for (int fk = 0; fk < 4; ++fk) {
                        const i64 i = bi * 32 + fi * 8 + g;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const i64 k = bk * 32 + fk * 8 + 2 * tq + e;
                            if (i < nt && k < nt && i >= k) acc[fi][fk][e] = V.get(r0 + i, r0 + k) - acc[fi][fk][e];
                        }
                    }
#pragma unroll
                for (int fi = 0; fi < 4; ++fi)
#pragma unroll
                    for (int fk = 0; fk < 4; ++fk) {
                        const i64 i = bi * 32 + fi * 8 + g;
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const i64 k = bk * 32 + fk * 8 + 2 * tq + e;
                            if (i < nt && k < nt && i >= k) V.set(r0 + i, r0 + k, acc[fi][fk][e]);
                        }
                    }
            }
        } else
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
        lap(2);
    }
    flush();
}

template <class T>
size_t potrf_smem(i64 n) {
    const i64 rows = (n + 31) / 32 * 32 + 4;
    return sizeof(T) * (size_t)(2 * JB * (JB + 1) + rows * potrf_ldp<T>::value);
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
    kern<<<1, potrf_threads<T>::value, smem, s>>>(n, A, lda, info, col_offset, g_potrf_prof ? 1 : 0);
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
int elb200_potrf_phase_clocks(unsigned long long out[4], int reset) {
    return guarded([&] {
        ELB_CUDA(cudaMemcpyFromSymbol(out, g_potrf_clk, sizeof(unsigned long long) * 4));
        if (reset) {
            unsigned long long z[4] = {0, 0, 0, 0};
            ELB_CUDA(cudaMemcpyToSymbol(g_potrf_clk, z, sizeof(z)));
        }
        g_potrf_prof = reset == 2 ? true : (reset == 3 ? false : g_potrf_prof);
    });
}
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
