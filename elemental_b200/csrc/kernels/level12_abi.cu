// The remaining leaf symbols of the reference's lower boundary as device kernels:
//   ?syr_ / ?her_   <- blas::Syr / Her (src/core/imports/blas/Syr.hpp:12-30,152-173): the rank-1 updates of the
//                      unblocked Cholesky (LowerVariant3.hpp:16-41, UpperVariant3.hpp:16-46)
//   ?scal_          <- blas::Scal (Scal.hpp:12-19,118-125)
//   ?axpy_          <- blas::Axpy (Axpy.hpp:12-29)
//   ?lacpy_         <- lapack::Copy (src/core/imports/lapack.cpp:20-31,381-396): the strided copies of
//                      copy::util::{InterleaveMatrix, RowStridedPack, ...}
// Fortran-77 ABI (by-reference scalars on the HOST, column-major DEVICE arrays, BlasInt = int), launched on the
// layer's current stream, plus stream-taking elb200_* forms.  All HBM-bound: one thread per element, rows fastest.
#include "../common.hpp"
#include "cplx.cuh"
#include "elb200_blas.h"

namespace elb200 {
namespace {

template <class T> struct real_of { typedef T type; };
template <class R> struct real_of<cplx<R>> { typedef R type; };

// A(i,j) += alpha x_i op(x_j) on the uplo triangle; HER: op = conj, alpha real, Im(A(j,j)) := 0
template <class T, bool HER>
__global__ void __launch_bounds__(256) syr_kernel(int lower, i64 n, T alpha, const T* __restrict__ x, i64 incx, i64 x0,
                                                  T* A, i64 lda) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const T xi = x[x0 + i * incx];
    for (i64 j = blockIdx.y; j < n; j += gridDim.y) {
        if (lower ? (i < j) : (i > j)) continue;
        T xj = x[x0 + j * incx];
        if (HER) xj = scalar_traits<T>::conj(xj);
        T v = A[i + j * lda] + alpha * (xi * xj);
        if (HER && i == j) v = scalar_traits<T>::from_real(scalar_traits<T>::real_part(v));
        A[i + j * lda] = v;
    }
}
template <class T, bool HER>
void syr_t(char uplo, i64 n, T alpha, const T* x, i64 incx, T* A, i64 lda, cudaStream_t s) {
    const char u = up(uplo);
    if (u != 'L' && u != 'U') throw std::logic_error("syr/her: uplo must be 'L' or 'U'");
    if (n < 0 || incx == 0 || lda < (n > 1 ? n : 1)) throw std::logic_error("syr/her: invalid argument");
    if (n == 0 || scalar_traits<T>::is_zero(alpha)) return;
    const i64 x0 = incx < 0 ? (1 - n) * incx : 0;
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)(n < 1024 ? n : 1024));
    syr_kernel<T, HER><<<grid, 256, 0, s>>>(u == 'L', n, alpha, x, incx, x0, A, lda);
    ELB_LAUNCH_CHECK();
}

template <class T>
__global__ void __launch_bounds__(256) scal_kernel(i64 n, T alpha, T* x, i64 incx) {
    for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < n; i += (i64)gridDim.x * 256) x[i * incx] = alpha * x[i * incx];
}
template <class T>
void scal_t(i64 n, T alpha, T* x, i64 incx, cudaStream_t s) {
    if (n <= 0 || incx <= 0) return;   // reference BLAS: nothing to do for incx <= 0
    i64 g = ceil_div(n, 256);
    if (g > 148 * 16) g = 148 * 16;
    scal_kernel<T><<<(unsigned)g, 256, 0, s>>>(n, alpha, x, incx);
    ELB_LAUNCH_CHECK();
}

template <class T>
__global__ void __launch_bounds__(256) axpy_kernel(i64 n, T alpha, const T* __restrict__ x, i64 incx, i64 x0, T* y,
                                                   i64 incy, i64 y0) {
    for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < n; i += (i64)gridDim.x * 256)
        y[y0 + i * incy] = y[y0 + i * incy] + alpha * x[x0 + i * incx];
}
template <class T>
void axpy_t(i64 n, T alpha, const T* x, i64 incx, T* y, i64 incy, cudaStream_t s) {
    if (n <= 0 || scalar_traits<T>::is_zero(alpha)) return;
    const i64 x0 = incx < 0 ? (1 - n) * incx : 0, y0 = incy < 0 ? (1 - n) * incy : 0;
    i64 g = ceil_div(n, 256);
    if (g > 148 * 16) g = 148 * 16;
    axpy_kernel<T><<<(unsigned)g, 256, 0, s>>>(n, alpha, x, incx, x0, y, incy, y0);
    ELB_LAUNCH_CHECK();
}

// mode 0 full, 1 upper triangle (i <= j), 2 lower triangle (i >= j)
template <class T>
__global__ void __launch_bounds__(256) lacpy_kernel(int mode, i64 m, i64 n, const T* __restrict__ A, i64 lda, T* B, i64 ldb) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    for (i64 j = blockIdx.y; j < n; j += gridDim.y) {
        if (mode == 1 && i > j) continue;
        if (mode == 2 && i < j) continue;
        B[i + j * ldb] = A[i + j * lda];
    }
}
template <class T>
void lacpy_t(char uplo, i64 m, i64 n, const T* A, i64 lda, T* B, i64 ldb, cudaStream_t s) {
    if (m < 0 || n < 0 || lda < (m > 1 ? m : 1) || ldb < (m > 1 ? m : 1)) throw std::logic_error("lacpy: invalid argument");
    if (m == 0 || n == 0) return;
    const char u = up(uplo);
    const int mode = u == 'U' ? 1 : (u == 'L' ? 2 : 0);
    if (mode == 0) {   // plain 2-D copy: the copy engine
        ELB_CUDA(cudaMemcpy2DAsync(B, sizeof(T) * (size_t)ldb, A, sizeof(T) * (size_t)lda, sizeof(T) * (size_t)m, (size_t)n,
                                   cudaMemcpyDefault, s));   // Default: the arrays may be mapped host memory
        return;
    }
    dim3 grid((unsigned)ceil_div(m, 256), (unsigned)(n < 2048 ? n : 2048));
    lacpy_kernel<T><<<grid, 256, 0, s>>>(mode, m, n, A, lda, B, ldb);
    ELB_LAUNCH_CHECK();
}

void report(int rc, const char* name) {
    if (rc != 0) fprintf(stderr, "elb200 %s: %s\n", name, elb200_last_error());
    elb200::fortran_abi_fence();
}
cudaStream_t cur() { return current_stream(); }
inline c32_t C32(elb200_c32 a) { return mk(a.re, a.im); }
inline c64_t C64(elb200_c64 a) { return mk(a.re, a.im); }

}  // namespace
}  // namespace elb200

extern "C" {
using namespace elb200;

#define ELB_L12(P, T, CT, CONV, REALT)                                                                                  \
    int elb200_##P##scal(int64_t n, CT alpha, CT* x, int64_t incx, elb200_stream_t s) {                                 \
        return guarded([&] { scal_t<T>(n, CONV(alpha), (T*)x, incx, (cudaStream_t)s); });                               \
    }                                                                                                                   \
    int elb200_##P##axpy(int64_t n, CT alpha, const CT* x, int64_t incx, CT* y, int64_t incy, elb200_stream_t s) {      \
        return guarded([&] { axpy_t<T>(n, CONV(alpha), (const T*)x, incx, (T*)y, incy, (cudaStream_t)s); });            \
    }                                                                                                                   \
    int elb200_##P##lacpy(char uplo, int64_t m, int64_t n, const CT* A, int64_t lda, CT* B, int64_t ldb,                \
                          elb200_stream_t s) {                                                                          \
        return guarded([&] { lacpy_t<T>(uplo, m, n, (const T*)A, lda, (T*)B, ldb, (cudaStream_t)s); });                 \
    }                                                                                                                   \
    void P##scal_(const int* n, const CT* alpha, CT* x, const int* incx) {                                              \
        report(elb200_##P##scal(*n, *alpha, x, *incx, (elb200_stream_t)cur()), #P "scal_");                             \
    }                                                                                                                   \
    void P##axpy_(const int* n, const CT* alpha, const CT* x, const int* incx, CT* y, const int* incy) {                \
        report(elb200_##P##axpy(*n, *alpha, x, *incx, y, *incy, (elb200_stream_t)cur()), #P "axpy_");                   \
    }                                                                                                                   \
    void P##lacpy_(const char* uplo, const int* m, const int* n, const CT* A, const int* lda, CT* B, const int* ldb) {  \
        report(elb200_##P##lacpy(*uplo, *m, *n, A, *lda, B, *ldb, (elb200_stream_t)cur()), #P "lacpy_");                \
    }
inline float idf(float a) { return a; }
inline double idd(double a) { return a; }
ELB_L12(s, float, float, idf, float)
ELB_L12(d, double, double, idd, double)
ELB_L12(c, c32_t, elb200_c32, C32, float)
ELB_L12(z, c64_t, elb200_c64, C64, double)

int elb200_ssyr(char uplo, int64_t n, float alpha, const float* x, int64_t incx, float* A, int64_t lda, elb200_stream_t s) {
    return guarded([&] { syr_t<float, false>(uplo, n, alpha, x, incx, A, lda, (cudaStream_t)s); });
}
int elb200_dsyr(char uplo, int64_t n, double alpha, const double* x, int64_t incx, double* A, int64_t lda, elb200_stream_t s) {
    return guarded([&] { syr_t<double, false>(uplo, n, alpha, x, incx, A, lda, (cudaStream_t)s); });
}
int elb200_cher(char uplo, int64_t n, float alpha, const elb200_c32* x, int64_t incx, elb200_c32* A, int64_t lda,
                elb200_stream_t s) {
    return guarded([&] { syr_t<c32_t, true>(uplo, n, mk(alpha, 0.f), (const c32_t*)x, incx, (c32_t*)A, lda, (cudaStream_t)s); });
}
int elb200_zher(char uplo, int64_t n, double alpha, const elb200_c64* x, int64_t incx, elb200_c64* A, int64_t lda,
                elb200_stream_t s) {
    return guarded([&] { syr_t<c64_t, true>(uplo, n, mk(alpha, 0.0), (const c64_t*)x, incx, (c64_t*)A, lda, (cudaStream_t)s); });
}
void ssyr_(const char* uplo, const int* n, const float* alpha, const float* x, const int* incx, float* A, const int* lda) {
    report(elb200_ssyr(*uplo, *n, *alpha, x, *incx, A, *lda, (elb200_stream_t)cur()), "ssyr_");
}
void dsyr_(const char* uplo, const int* n, const double* alpha, const double* x, const int* incx, double* A, const int* lda) {
    report(elb200_dsyr(*uplo, *n, *alpha, x, *incx, A, *lda, (elb200_stream_t)cur()), "dsyr_");
}
void cher_(const char* uplo, const int* n, const float* alpha, const elb200_c32* x, const int* incx, elb200_c32* A,
           const int* lda) {
    report(elb200_cher(*uplo, *n, *alpha, x, *incx, A, *lda, (elb200_stream_t)cur()), "cher_");
}
void zher_(const char* uplo, const int* n, const double* alpha, const elb200_c64* x, const int* incx, elb200_c64* A,
           const int* lda) {
    report(elb200_zher(*uplo, *n, *alpha, x, *incx, A, *lda, (elb200_stream_t)cur()), "zher_");
}

}  // extern "C"
