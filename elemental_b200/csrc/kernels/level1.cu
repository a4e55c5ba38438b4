// Memory-bound helpers of the hot path: strided block moves (pack / unpack /
// transpose / axpy / scale in one kernel), trapezoid scaling, counter-hash
// fills and norm reductions.  See include/elb200_level1.h for what each one
// replaces in the reference.  All of these are HBM-bound; they are written
// for coalesced 32-wide row access with a padded shared-memory tile when the
// source is read transposed.
#include "../common.hpp"
#include "cplx.cuh"
#include "elb200_level1.h"
#include "device_api.hpp"

namespace elb200 {
namespace {

constexpr int MAX_BATCH = 16;
struct LatticeBatch {
    elb200_lattice d[MAX_BATCH];
    int n;
};

template <class T, bool CONJ>
__global__ void __launch_bounds__(256) lattice_copy_kernel(const LatticeBatch b, const T alpha,
                                                           const int has_alpha, const int acc) {
    __shared__ T sm[32][33];
    const elb200_lattice d = b.d[blockIdx.y];
    const T* src = (const T*)d.src;  // may alias dst (in-place scale)
    T* dst = (T*)d.dst;
    const i64 tiles_r = (d.nrows + 31) / 32, tiles_c = (d.ncols + 31) / 32;
    const i64 ntiles = tiles_r * tiles_c;
    const bool row_fast = (d.s_rs <= d.s_cs);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 t0 = (tile % tiles_r) * 32, u0 = (tile / tiles_r) * 32;
        // read: fastest thread index follows the smaller source stride
#pragma unroll
        for (int k = ty; k < 32; k += 8) {
            const i64 t = row_fast ? t0 + tx : t0 + k;
            const i64 u = row_fast ? u0 + k : u0 + tx;
            if (t < d.nrows && u < d.ncols) {
                T v = src[d.s_off + t * d.s_rs + u * d.s_cs];
                if (CONJ) v = scalar_traits<T>::conj(v);
                if (row_fast) sm[k][tx] = v;  // sm[u][t]
                else sm[tx][k] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k = ty; k < 32; k += 8) {
            const i64 t = t0 + tx, u = u0 + k;
            if (t < d.nrows && u < d.ncols) {
                T v = sm[k][tx];
                if (has_alpha) v = alpha * v;
                T* p = dst + d.d_off + t * d.d_rs + u * d.d_cs;
                if (acc) v = *p + v;
                *p = v;
            }
        }
        __syncthreads();
    }
}

template <class T>
void lattice_copy_t(const elb200_lattice* descs, int ndesc, int conj, const void* alpha, int acc,
                    cudaStream_t s) {
    T a = scalar_traits<T>::from_real(1);
    int has_alpha = 0;
    if (alpha) {
        a = *(const T*)alpha;
        has_alpha = scalar_traits<T>::is_one(a) ? 0 : 1;
    }
    const int cap = sm_count() * 8;
    int i = 0;  // scan index: every batch resumes where the previous one stopped (empty descriptors are skipped)
    while (i < ndesc) {
        LatticeBatch b;
        b.n = 0;
        i64 maxtiles = 0;
        for (; i < ndesc && b.n < MAX_BATCH; ++i) {
            if (descs[i].nrows <= 0 || descs[i].ncols <= 0) continue;
            b.d[b.n++] = descs[i];
            i64 t = ceil_div(descs[i].nrows, 32) * ceil_div(descs[i].ncols, 32);
            if (t > maxtiles) maxtiles = t;
        }
        if (b.n == 0) continue;
        dim3 grid((unsigned)(maxtiles < cap ? maxtiles : cap), (unsigned)b.n);
        if (conj && scalar_traits<T>::is_complex)
            lattice_copy_kernel<T, true><<<grid, 256, 0, s>>>(b, a, has_alpha, acc);
        else
            lattice_copy_kernel<T, false><<<grid, 256, 0, s>>>(b, a, has_alpha, acc);
        ELB_LAUNCH_CHECK();
    }
}

// MODE 0: scale inside trapezoid; MODE 1: zero outside trapezoid
template <class T, int MODE>
__global__ void __launch_bounds__(256) trapezoid_kernel(T alpha, int lower, i64 m, i64 n, T* A, i64 lda,
                                                        i64 gi0, i64 gis, i64 gj0, i64 gjs, i64 offset) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const i64 gi = gi0 + i * gis;
    for (i64 j = blockIdx.y; j < n; j += gridDim.y) {
        const i64 gj = gj0 + j * gjs;
        const bool inside = lower ? (gj - gi <= offset) : (gj - gi >= offset);
        T* p = A + i + j * lda;
        if (MODE == 0) {
            if (inside) *p = alpha * (*p);
        } else {
            if (!inside) *p = scalar_traits<T>::zero();
        }
    }
}

// Y += alpha X inside the trapezoid (same predicate), X and Y local matrices of identically distributed operands
template <class T>
__global__ void __launch_bounds__(256) axpy_trapezoid_kernel(T alpha, int lower, i64 m, i64 n, const T* __restrict__ X,
                                                             i64 ldx, T* Y, i64 ldy, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                                                             i64 offset) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const i64 gi = gi0 + i * gis;
    for (i64 j = blockIdx.y; j < n; j += gridDim.y) {
        const i64 gj = gj0 + j * gjs;
        const bool inside = lower ? (gj - gi <= offset) : (gj - gi >= offset);
        if (inside) Y[i + j * ldy] = Y[i + j * ldy] + alpha * X[i + j * ldx];
    }
}
template <class T>
void axpy_trapezoid_t(const void* alpha, char uplo, i64 m, i64 n, const void* X, i64 ldx, void* Y, i64 ldy, i64 gi0,
                      i64 gis, i64 gj0, i64 gjs, i64 offset, cudaStream_t s) {
    if (m <= 0 || n <= 0) return;
    const char u = up(uplo);
    if (u != 'L' && u != 'U') throw std::logic_error("axpy_trapezoid: uplo must be 'L' or 'U'");
    T a = alpha ? *(const T*)alpha : scalar_traits<T>::from_real(1);
    dim3 grid((unsigned)ceil_div(m, 256), (unsigned)(n < 1024 ? n : 1024));
    axpy_trapezoid_kernel<T><<<grid, 256, 0, s>>>(a, u == 'L', m, n, (const T*)X, ldx, (T*)Y, ldy, gi0, gis, gj0, gjs, offset);
    ELB_LAUNCH_CHECK();
}

template <class T, int MODE>
void trapezoid_t(const void* alpha, char uplo, i64 m, i64 n, void* A, i64 lda, i64 gi0, i64 gis,
                 i64 gj0, i64 gjs, i64 offset, cudaStream_t s) {
    if (m <= 0 || n <= 0) return;
    const char u = up(uplo);
    if (u != 'L' && u != 'U') throw std::logic_error("trapezoid: uplo must be 'L' or 'U'");
    T a = alpha ? *(const T*)alpha : scalar_traits<T>::from_real(1);
    dim3 grid((unsigned)ceil_div(m, 256), (unsigned)(n < 1024 ? n : 1024));
    trapezoid_kernel<T, MODE><<<grid, 256, 0, s>>>(a, u == 'L', m, n, (T*)A, lda, gi0, gis, gj0, gjs, offset);
    ELB_LAUNCH_CHECK();
}

__host__ __device__ inline unsigned long long mix64(unsigned long long x) {
    x ^= x >> 30;
    x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27;
    x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
// u(i,j;seed) in [-1,1); restated verbatim in oracle/generator.py
__host__ __device__ inline double hash_uniform(unsigned long long i, unsigned long long j,
                                               unsigned long long seed) {
    unsigned long long x = mix64(seed + 0x9E3779B97F4A7C15ULL);
    x = mix64(x ^ (i * 0xD1B54A32D192ED03ULL + 0x2545F4914F6CDD1DULL));
    x = mix64(x ^ (j * 0x8CB92BA72F3D8DD7ULL + 0x9E6C63D0676A9A99ULL));
    return (double)(x >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

template <class T>
__device__ inline T make_entry(double re, double im);
template <> __device__ inline float make_entry<float>(double re, double) { return (float)re; }
template <> __device__ inline double make_entry<double>(double re, double) { return re; }
template <> __device__ inline c32_t make_entry<c32_t>(double re, double im) { return mk((float)re, (float)im); }
template <> __device__ inline c64_t make_entry<c64_t>(double re, double im) { return mk(re, im); }

template <class T>
__global__ void __launch_bounds__(256) fill_hash_kernel(int kind, i64 m, i64 n, T* A, i64 lda, i64 gi0,
                                                        i64 gis, i64 gj0, i64 gjs,
                                                        unsigned long long seed, double diag) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    const unsigned long long gi = (unsigned long long)(gi0 + i * gis);
    for (i64 j = blockIdx.y; j < n; j += gridDim.y) {
        const unsigned long long gj = (unsigned long long)(gj0 + j * gjs);
        double re, im = 0.0;
        if (kind == 0) {
            re = hash_uniform(gi, gj, seed);
            if (scalar_traits<T>::is_complex) im = hash_uniform(gi, gj, seed + 1);
        } else {
            const unsigned long long lo = gi < gj ? gi : gj, hi = gi < gj ? gj : gi;
            re = hash_uniform(lo, hi, seed);
            if (scalar_traits<T>::is_complex && gi != gj) {
                im = hash_uniform(lo, hi, seed + 1);
                if (gi > gj) im = -im;  // lower triangle holds the conjugate
            }
            if (gi == gj) re += diag;
        }
        A[i + j * lda] = make_entry<T>(re, im);
    }
}

template <class T>
void fill_hash_t(int kind, i64 m, i64 n, void* A, i64 lda, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                 unsigned long long seed, double diag, cudaStream_t s) {
    if (m <= 0 || n <= 0) return;
    dim3 grid((unsigned)ceil_div(m, 256), (unsigned)(n < 2048 ? n : 2048));
    fill_hash_kernel<T><<<grid, 256, 0, s>>>(kind, m, n, (T*)A, lda, gi0, gis, gj0, gjs, seed, diag);
    ELB_LAUNCH_CHECK();
}

__device__ inline void atomic_max_nonneg(double* addr, double v) {
    // valid for non-negative doubles: their bit patterns order like integers, +inf above every finite
    // value and the canonical quiet NaN above +inf, so a NaN anywhere survives the reduction
    if (v != v) v = __longlong_as_double(0x7ff8000000000000LL);
    atomicMax((unsigned long long*)addr, (unsigned long long)__double_as_longlong(v));
}
template <class T> __device__ inline double abs_entry(T x) { return fabs((double)x); }
template <> __device__ inline double abs_entry<c32_t>(c32_t x) { return hypot((double)x.re, (double)x.im); }
template <> __device__ inline double abs_entry<c64_t>(c64_t x) { return hypot(x.re, x.im); }
// NaN-propagating max: a NaN operand wins and is never dropped again
__device__ inline double nanmax(double acc, double v) { return (v > acc || v != v) ? v : acc; }

// MODE 0: sum of |a_ij / scale|^2 (scale read from device memory, 1 when the pointer is null or the
// value is 0 / not finite); MODE 1: max |a_ij|, NaN-propagating.  Together they give the scaled
// two-pass Frobenius norm max * sqrt(sum |a/max|^2), which neither overflows nor underflows
// (the reference's UpdateScaledSquare does the same in one pass).
template <class T, int MODE>
__global__ void __launch_bounds__(256) reduce_kernel(i64 m, i64 n, const T* __restrict__ A, i64 lda,
                                                     const double* __restrict__ scale_dev, double* out) {
    double acc = 0.0;
    double inv = 1.0;
    if (MODE == 0 && scale_dev) {
        const double sc = *scale_dev;
        if (sc > 0.0 && sc < __longlong_as_double(0x7ff0000000000000LL)) inv = 1.0 / sc;
    }
    for (i64 j = blockIdx.y; j < n; j += gridDim.y)
        for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < m; i += (i64)gridDim.x * 256) {
            if (MODE == 0) {
                const double ab = abs_entry<T>(A[i + j * lda]) * inv;
                acc += ab * ab;
            } else {
                acc = nanmax(acc, abs_entry<T>(A[i + j * lda]));
            }
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double other = __shfl_xor_sync(0xffffffffu, acc, o);
        if (MODE == 0) acc += other;
        else acc = nanmax(acc, other);
    }
    __shared__ double ws[8];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double r = ws[0];
        for (int w = 1; w < 8; ++w) {
            if (MODE == 0) r += ws[w];
            else r = nanmax(r, ws[w]);
        }
        if (MODE == 0) atomicAdd(out, r);
        else atomic_max_nonneg(out, r);
    }
}

template <class T, int MODE>
void reduce_t(i64 m, i64 n, const void* A, i64 lda, const double* scale_dev, double* out, cudaStream_t s) {
    if (m <= 0 || n <= 0) return;
    i64 gx = ceil_div(m, 256);
    if (gx > 64) gx = 64;
    dim3 grid((unsigned)gx, (unsigned)(n < 256 ? n : 256));
    reduce_kernel<T, MODE><<<grid, 256, 0, s>>>(m, n, (const T*)A, lda, scale_dev, out);
    ELB_LAUNCH_CHECK();
}

// first j with A(j,j) == 0 -> *flag = offset + j + 1 (only the smallest index survives; a flag that is
// already nonzero from an earlier block of the same solve is kept)
template <class T>
__global__ void __launch_bounds__(256) diag_zero_kernel(i64 n, const T* __restrict__ A, i64 lda, i64 offset, int* flag) {
    __shared__ int first;
    if (threadIdx.x == 0) first = 0x7fffffff;
    __syncthreads();
    for (i64 j = threadIdx.x; j < n; j += 256)
        if (scalar_traits<T>::is_zero(A[j + j * lda])) atomicMin(&first, (int)j);
    __syncthreads();
    if (threadIdx.x == 0 && first != 0x7fffffff) atomicCAS(flag, 0, (int)(offset + first + 1));
}
template <class T>
void diag_zero_t(i64 n, const void* A, i64 lda, i64 offset, int* flag, cudaStream_t s) {
    if (n <= 0) return;
    diag_zero_kernel<T><<<1, 256, 0, s>>>(n, (const T*)A, lda, offset, flag);
    ELB_LAUNCH_CHECK();
}

}  // namespace

template <class T>
void lattice_copy_device(const T* src, T* dst, i64 nrows, i64 ncols, i64 s_off, i64 s_rs, i64 s_cs,
                         i64 d_off, i64 d_rs, i64 d_cs, bool conj, const T* alpha, bool accumulate,
                         cudaStream_t s) {
    elb200_lattice d;
    d.src = src; d.dst = dst; d.nrows = nrows; d.ncols = ncols;
    d.s_off = s_off; d.s_rs = s_rs; d.s_cs = s_cs;
    d.d_off = d_off; d.d_rs = d_rs; d.d_cs = d_cs;
    lattice_copy_t<T>(&d, 1, conj ? 1 : 0, alpha, accumulate ? 1 : 0, s);
}
template void lattice_copy_device<float>(const float*, float*, i64, i64, i64, i64, i64, i64, i64, i64, bool, const float*, bool, cudaStream_t);
template void lattice_copy_device<double>(const double*, double*, i64, i64, i64, i64, i64, i64, i64, i64, bool, const double*, bool, cudaStream_t);
template void lattice_copy_device<c32_t>(const c32_t*, c32_t*, i64, i64, i64, i64, i64, i64, i64, i64, bool, const c32_t*, bool, cudaStream_t);
template void lattice_copy_device<c64_t>(const c64_t*, c64_t*, i64, i64, i64, i64, i64, i64, i64, i64, bool, const c64_t*, bool, cudaStream_t);

namespace {
#define DISPATCH_DTYPE(dtype, CALL)                                   \
    switch (dtype) {                                                  \
        case ELB200_F32: { typedef float T; CALL; } break;            \
        case ELB200_F64: { typedef double T; CALL; } break;           \
        case ELB200_C32: { typedef c32_t T; CALL; } break;            \
        case ELB200_C64: { typedef c64_t T; CALL; } break;            \
        default: throw std::logic_error("unknown dtype code");        \
    }

}  // namespace
}  // namespace elb200

extern "C" {
using namespace elb200;

int elb200_lattice_copy(int dtype, const elb200_lattice* descs, int ndesc, int conj,
                        const void* alpha, int accumulate, elb200_stream_t s) {
    return guarded([&] {
        DISPATCH_DTYPE(dtype, (lattice_copy_t<T>(descs, ndesc, conj, alpha, accumulate, (cudaStream_t)s)));
    });
}

int elb200_scale_trapezoid(int dtype, const void* alpha, char uplo, int64_t m, int64_t n, void* A,
                           int64_t lda, int64_t rowShift, int64_t rowStride, int64_t colShift,
                           int64_t colStride, int64_t offset, elb200_stream_t s) {
    return guarded([&] {
        DISPATCH_DTYPE(dtype, (trapezoid_t<T, 0>(alpha, uplo, m, n, A, lda, rowShift, rowStride, colShift,
                                                 colStride, offset, (cudaStream_t)s)));
    });
}

int elb200_axpy_trapezoid(int dtype, const void* alpha, char uplo, int64_t m, int64_t n, const void* X, int64_t ldx,
                          void* Y, int64_t ldy, int64_t rowShift, int64_t rowStride, int64_t colShift, int64_t colStride,
                          int64_t offset, elb200_stream_t s) {
    return guarded([&] {
        DISPATCH_DTYPE(dtype, (axpy_trapezoid_t<T>(alpha, uplo, m, n, X, ldx, Y, ldy, rowShift, rowStride, colShift,
                                                   colStride, offset, (cudaStream_t)s)));
    });
}

int elb200_make_trapezoidal(int dtype, char uplo, int64_t m, int64_t n, void* A, int64_t lda,
                            int64_t rowShift, int64_t rowStride, int64_t colShift,
                            int64_t colStride, int64_t offset, elb200_stream_t s) {
    return guarded([&] {
        DISPATCH_DTYPE(dtype, (trapezoid_t<T, 1>(nullptr, uplo, m, n, A, lda, rowShift, rowStride, colShift,
                                                 colStride, offset, (cudaStream_t)s)));
    });
}

int elb200_fill_hash(int dtype, int kind, int64_t m, int64_t n, void* A, int64_t lda,
                     int64_t rowShift, int64_t rowStride, int64_t colShift, int64_t colStride,
                     uint64_t seed, double diag, elb200_stream_t s) {
    return guarded([&] {
        DISPATCH_DTYPE(dtype, (fill_hash_t<T>(kind, m, n, A, lda, rowShift, rowStride, colShift, colStride,
                                              seed, diag, (cudaStream_t)s)));
    });
}

int elb200_diag_zero_check(int dtype, int64_t n, const void* A, int64_t lda, int64_t offset, int* flag_dev,
                           elb200_stream_t s) {
    return guarded([&] { DISPATCH_DTYPE(dtype, (diag_zero_t<T>(n, A, lda, offset, flag_dev, (cudaStream_t)s))); });
}

int elb200_sumsq(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, double* out_dev,
                 elb200_stream_t s) {
    return guarded([&] { DISPATCH_DTYPE(dtype, (reduce_t<T, 0>(m, n, A, lda, nullptr, out_dev, (cudaStream_t)s))); });
}

int elb200_sumsq_scaled(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, const double* scale_dev,
                        double* out_dev, elb200_stream_t s) {
    return guarded([&] { DISPATCH_DTYPE(dtype, (reduce_t<T, 0>(m, n, A, lda, scale_dev, out_dev, (cudaStream_t)s))); });
}

int elb200_maxabs(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, double* out_dev,
                  elb200_stream_t s) {
    return guarded([&] { DISPATCH_DTYPE(dtype, (reduce_t<T, 1>(m, n, A, lda, nullptr, out_dev, (cudaStream_t)s))); });
}

}  // extern "C"
