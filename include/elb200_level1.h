/*
 * elb200_level1.h -- memory-bound device helpers of the hot path (C-ABI).
 *
 * These replace the host loops the reference runs between its BLAS calls and
 * its MPI calls (paths relative to the Elemental tree):
 *   elb200_lattice_copy   <- copy::util::{InterleaveMatrix,RowStridedPack/Unpack,
 *                            ColStridedPack/Unpack,PartialCol/RowStrided*,StridedPack/Unpack}
 *                            (include/El/blas_like/level1/Copy/util.hpp:16-458), the blocked
 *                            CPU Transpose (level1/Transpose.hpp:87-129), Copy(Matrix)
 *                            (level1/Copy.hpp:18-55), Axpy/InterleaveMatrixUpdate
 *                            (level1/AxpyContract.hpp:300-303) and Scale
 *   elb200_scale_trapezoid<- ScaleTrapezoid (level1/ScaleTrapezoid.hpp:14-86)
 *   elb200_make_trapezoidal <- MakeTrapezoidal
 *   elb200_axpy_trapezoid <- AxpyTrapezoid / LocalAxpyTrapezoid (level1/AxpyTrapezoid.hpp:14-49,73-125)
 *   elb200_fill_hash      <- test/bench input generation on global indices (grid independent;
 *                            stands in for Uniform / HermitianUniformSpectrum)
 *   elb200_sumsq / elb200_maxabs <- FrobeniusNorm / MaxNorm pieces used by the residual checks
 *
 * dtype codes: 0 = float, 1 = double, 2 = complex<float>, 3 = complex<double>.
 * All pointers are DEVICE pointers unless stated otherwise.
 */
#ifndef ELB200_LEVEL1_H
#define ELB200_LEVEL1_H

#include <stdint.h>
#include "elb200_blas.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { ELB200_F32 = 0, ELB200_F64 = 1, ELB200_C32 = 2, ELB200_C64 = 3 };

/* One strided 2-D block move: for t < nrows, u < ncols
 *   dst[d_off + t*d_rs + u*d_cs]  (op)=  src[s_off + t*s_rs + u*s_cs]
 * (offsets and strides in ELEMENTS).  A transpose is just swapped source strides. */
typedef struct {
    const void* src;
    void* dst;
    int64_t nrows, ncols;
    int64_t s_off, s_rs, s_cs;
    int64_t d_off, d_rs, d_cs;
} elb200_lattice;

/* dst = alpha*op(src)             (accumulate == 0)
 * dst = dst + alpha*op(src)       (accumulate != 0)
 * op = conj when conj != 0; alpha is a HOST pointer to one element of the dtype, NULL = 1. */
int elb200_lattice_copy(int dtype, const elb200_lattice* descs, int ndesc, int conj,
                        const void* alpha, int accumulate, elb200_stream_t s);

/* Scale the entries of the local block whose GLOBAL indices (rowShift + i*rowStride,
 * colShift + j*colStride) satisfy gj - gi <= offset ('L') or gj - gi >= offset ('U'). */
int elb200_scale_trapezoid(int dtype, const void* alpha, char uplo, int64_t m, int64_t n, void* A,
                           int64_t lda, int64_t rowShift, int64_t rowStride, int64_t colShift,
                           int64_t colStride, int64_t offset, elb200_stream_t s);
/* Y += alpha X restricted to the same trapezoid: AxpyTrapezoid / LocalAxpyTrapezoid
 * (include/El/blas_like/level1/AxpyTrapezoid.hpp:14-49,73-125); X, Y local matrices of equally distributed operands */
int elb200_axpy_trapezoid(int dtype, const void* alpha, char uplo, int64_t m, int64_t n, const void* X, int64_t ldx,
                          void* Y, int64_t ldy, int64_t rowShift, int64_t rowStride, int64_t colShift, int64_t colStride,
                          int64_t offset, elb200_stream_t s);
/* Zero everything OUTSIDE that trapezoid. */
int elb200_make_trapezoidal(int dtype, char uplo, int64_t m, int64_t n, void* A, int64_t lda,
                            int64_t rowShift, int64_t rowStride, int64_t colShift,
                            int64_t colStride, int64_t offset, elb200_stream_t s);

/* Counter-hash fill on global indices; identical formula in oracle/generator.py.
 * kind 0: general, a_ij = u(i,j)            (complex: re = u(i,j;seed), im = u(i,j;seed+1))
 * kind 1: Hermitian, a_ij = u(min,max) (+ i*sign*u(min,max;seed+1)), real diagonal, + diag on i==j */
int elb200_fill_hash(int dtype, int kind, int64_t m, int64_t n, void* A, int64_t lda,
                     int64_t rowShift, int64_t rowStride, int64_t colShift, int64_t colStride,
                     uint64_t seed, double diag, elb200_stream_t s);

/* The checkIfSingular scan of El::Trsm (src/blas_like/level3/Trsm.cpp:54-60: a host loop over A.Get(j,j)):
 * if some A(j,j) == 0 and *flag_dev (device int) is still 0, it receives offset + j + 1 for the smallest such j.
 * The host layer turns a nonzero flag into SingularMatrixException. */
int elb200_diag_zero_check(int dtype, int64_t n, const void* A, int64_t lda, int64_t offset, int* flag_dev,
                           elb200_stream_t s);

/* *out_dev (device double) += sum |a_ij|^2 over the local block */
int elb200_sumsq(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, double* out_dev,
                 elb200_stream_t s);
/* *out_dev += sum |a_ij / *scale_dev|^2 (scale_dev: device double, e.g. the result of elb200_maxabs;
 * NULL, 0 or a non-finite value means 1): second pass of the scaled Frobenius norm
 * (the reference's UpdateScaledSquare, src/lapack_like/props/Norm/Frobenius.cpp) */
int elb200_sumsq_scaled(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, const double* scale_dev,
                        double* out_dev, elb200_stream_t s);
/* *out_dev (device double) = max(*out_dev, max |a_ij|); *out_dev must be >= 0 on entry.
 * A NaN entry yields NaN (the reference's MaxNorm propagates NaN too). */
int elb200_maxabs(int dtype, int64_t m, int64_t n, const void* A, int64_t lda, double* out_dev,
                  elb200_stream_t s);

#ifdef __cplusplus
}
#endif
#endif
