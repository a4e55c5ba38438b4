/*
 * elb200_blas.h -- the leaf C-ABI of the B200-native Elemental hot path.
 *
 * Every entry point takes DEVICE pointers (column-major, leading dimension in
 * elements) and launches hand-written sm_100a kernels on the given CUDA
 * stream.  There is no CPU fallback: without a CUDA device every call fails
 * with a nonzero return code and a message in elb200_last_error().
 *
 * What each group replaces in the reference (paths relative to the Elemental
 * tree):
 *
 *   elb200_{s,d,c,z}gemm   <- El::blas::Gemm   include/El/core/imports/blas.hpp:570-603,
 *                             which binds sgemm_/dgemm_/cgemm_/zgemm_
 *                             (src/core/imports/blas/Gemm.hpp:13-40,388,431,471,511)
 *   elb200_{s,d,c,z}trsm   <- El::blas::Trsm   blas.hpp:884-913 (src/core/imports/blas/Trsm.hpp:12-33)
 *   elb200_{s,d}syrk, elb200_{c,z}herk, elb200_{c,z}syrk
 *                          <- El::blas::Syrk/Herk  blas.hpp:689-722,813-846
 *   elb200_{s,d,c,z}trrk   <- the LocalTrrk recursion, src/blas_like/level3/Trrk/Local.hpp:782-830
 *                             (one masked GEMM instead of gemm + AxpyTrapezoid leaves)
 *   elb200_{s,d,c,z}potrf  <- cholesky::LowerVariant3Unblocked / UpperVariant3Unblocked,
 *                             src/lapack_like/factor/Cholesky/LowerVariant3.hpp:16-41,
 *                             UpperVariant3.hpp:16-46 (the reference never calls LAPACK potrf)
 *   Fortran-ABI symbols (dgemm_, dtrsm_, dsyrk_, zherk_, ...) with the exact
 *   prototypes the reference declares; they forward to the functions above on
 *   the layer's current stream (elb200_set_stream).
 */
#ifndef ELB200_BLAS_H
#define ELB200_BLAS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* elb200_stream_t; /* == cudaStream_t */
typedef struct { float re, im; } elb200_c32;
typedef struct { double re, im; } elb200_c64;

/* ---- runtime ---------------------------------------------------------- */
const char* elb200_last_error(void);
int elb200_version(void);
/* 0 when a usable sm_100 device is present, nonzero (+last_error) otherwise */
int elb200_device_check(void);
void elb200_set_stream(elb200_stream_t s);
elb200_stream_t elb200_get_stream(void);
/* Cap on the CTAs a persistent GEMM launch may occupy (0 = one per SM).  The overlapped panel
 * loops of the host layer lower it so that NCCL / pack / potrf kernels of the panel stream find
 * free SMs beside the trailing update. */
void elb200_set_sm_limit(int n);
int elb200_get_sm_limit(void);

/* kernels launched by this library since the last reset (bench.py: "gpu_launches") */
unsigned long long elb200_launch_count(int reset);
/* Per-launch CUDA-event timing of the FP64 GEMM/TRRK kernel on its launching stream:
 * enable, run, then read the summed device time, launch count and algorithmic flops
 * (2mnk per launch; for TRRK only the tiles that are not skipped). */
void elb200_gemm_profile(int enable);
int elb200_gemm_profile_read(double* total_ms, long long* launches, double* flops);

/* FP64 GEMM kernel selection: 0 automatic (default: the persistent warp-specialised TMA kernel
 * whenever A and B are 16-byte aligned with even leading dimension, else the cp.async kernel),
 * 1 = cp.async kernel, 128x128 CTA tile, one CTA per SM, 2 = cp.async kernel, 128x64 CTA tile,
 * two CTAs per SM, 3 = same as 0 */
void elb200_dgemm_set_config(int cfg);
/* which kernel the last elb200_dgemm / dtrrk / dsyrk call launched: 1 cp.async, 2 TMA */
int elb200_dgemm_last_kernel(void);
/* Debugging aids of the persistent TMA kernel (gemm_f64_ws.cu).  Flags: bit 4 = never use the L2 reduction epilogue
 * (tests compare it with the load-add-store form bit for bit), bit 10 = phase-clock diagnostic build (NN only),
 * bits 20..25 = rasterisation band width, bits 26..29 = 1: no group stagger, v > 1: stagger of v - 1 microseconds.
 * The profile buffer is a device array of (grid x 16 warps x 8) unsigned 64-bit counters for the phase clocks.
 * elb200_dgemm_ws_last_maps: bit 0 / 1 = A / B of the last launch used the single-box 3-D tensor map. */
void elb200_dgemm_set_debug_flags(int flags);
void elb200_dgemm_set_profile_buffer(void* device_u64);
int elb200_dgemm_ws_last_maps(void);

/* ---- GEMM: C := alpha op(A) op(B) + beta C ---------------------------- */
/* trans in {'N','T','C'}; for real types 'C' == 'T' (blas/Gemm.hpp:386-387) */
int elb200_dgemm(char transA, char transB, int64_t m, int64_t n, int64_t k,
                 double alpha, const double* A, int64_t lda,
                 const double* B, int64_t ldb,
                 double beta, double* C, int64_t ldc, elb200_stream_t s);
int elb200_sgemm(char transA, char transB, int64_t m, int64_t n, int64_t k,
                 float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb,
                 float beta, float* C, int64_t ldc, elb200_stream_t s);
int elb200_zgemm(char transA, char transB, int64_t m, int64_t n, int64_t k,
                 elb200_c64 alpha, const elb200_c64* A, int64_t lda,
                 const elb200_c64* B, int64_t ldb,
                 elb200_c64 beta, elb200_c64* C, int64_t ldc, elb200_stream_t s);
int elb200_cgemm(char transA, char transB, int64_t m, int64_t n, int64_t k,
                 elb200_c32 alpha, const elb200_c32* A, int64_t lda,
                 const elb200_c32* B, int64_t ldb,
                 elb200_c32 beta, elb200_c32* C, int64_t ldc, elb200_stream_t s);
/* Complex<double> products: 0 = automatic (large ones run on the real persistent kernel through the
 * (re, im)-row-pair identity of kernels/gemm_c64_real.cu), 1 = always the dedicated complex kernel.
 * elb200_zgemm_last_kernel: 1 complex cp.async kernel, 2 real persistent kernel. */
void elb200_zgemm_set_path(int path);
int elb200_zgemm_last_kernel(void);
/* float GEMM on tcgen05 (kind::tf32) with the 3xTF32 operand split;
 * relative error per product ~2^-21 instead of 2^-24 */
/* the 3xTF32 kernel's tile rasterisation (host copy of the device function; tests check it is a bijection) */
void elb200_tf32_tile_coords(int64_t tile, int64_t tilesM, int64_t tilesN, int64_t* tileRow, int64_t* tileCol);
int elb200_sgemm_3xtf32(char transA, char transB, int64_t m, int64_t n, int64_t k,
                 float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb,
                 float beta, float* C, int64_t ldc, elb200_stream_t s);
/* Which arithmetic elb200_sgemm / sgemm_ / El::Gemm<float> use for full GEMMs: 0 = exact FFMA
 * (default), 1 = 3xTF32 on tcgen05 whenever the operands are 16-byte aligned with ld % 4 == 0
 * (otherwise, and for the masked TRRK form, exact FFMA).  elb200_sgemm_last_kernel: 1 SIMT, 2 tcgen05. */
void elb200_sgemm_set_mode(int mode);
/* exact-FFMA path: 0 automatic (register-tiled kernel gemm_f32_ffma.cu when A, B are 16-byte aligned with
 * ld % 4 == 0), 1 always the generic SIMT kernel; last_kernel: 1 generic, 2 register-tiled */
void elb200_sgemm_set_ffma_path(int path);
int elb200_sgemm_ffma_last_kernel(void);
int elb200_sgemm_get_mode(void);
int elb200_sgemm_last_kernel(void);

/* ---- TRRK: triangle-restricted rank-k update --------------------------- */
/* C_tri := alpha op(A) op(B) + beta C_tri, touching only entries whose GLOBAL
 * indices (gi = rowShift + i*rowStride, gj = colShift + j*colStride) lie in
 * the uplo triangle (gi >= gj for 'L', gi <= gj for 'U').  With shift 0 and
 * stride 1 this is a plain local Trrk; with the [MC,MR] shifts/strides it is
 * the staircase update of LocalTrrk (Trrk/Local.hpp:782-830). */
int elb200_dtrrk(char uplo, char transA, char transB, int64_t m, int64_t n, int64_t k,
                 double alpha, const double* A, int64_t lda,
                 const double* B, int64_t ldb,
                 double beta, double* C, int64_t ldc,
                 int64_t rowShift, int64_t rowStride,
                 int64_t colShift, int64_t colStride, elb200_stream_t s);
int elb200_strrk(char uplo, char transA, char transB, int64_t m, int64_t n, int64_t k,
                 float alpha, const float* A, int64_t lda,
                 const float* B, int64_t ldb,
                 float beta, float* C, int64_t ldc,
                 int64_t rowShift, int64_t rowStride,
                 int64_t colShift, int64_t colStride, elb200_stream_t s);
int elb200_ztrrk(char uplo, char transA, char transB, int64_t m, int64_t n, int64_t k,
                 elb200_c64 alpha, const elb200_c64* A, int64_t lda,
                 const elb200_c64* B, int64_t ldb,
                 elb200_c64 beta, elb200_c64* C, int64_t ldc,
                 int64_t rowShift, int64_t rowStride,
                 int64_t colShift, int64_t colStride, elb200_stream_t s);
int elb200_ctrrk(char uplo, char transA, char transB, int64_t m, int64_t n, int64_t k,
                 elb200_c32 alpha, const elb200_c32* A, int64_t lda,
                 const elb200_c32* B, int64_t ldb,
                 elb200_c32 beta, elb200_c32* C, int64_t ldc,
                 int64_t rowShift, int64_t rowStride,
                 int64_t colShift, int64_t colStride, elb200_stream_t s);

/* ---- SYRK / HERK (local, blas.hpp:689-722,813-846) --------------------- */
/* trans 'N': C := alpha A A^T + beta C (A is n x k); 'T'/'C': A^T A (A is k x n) */
int elb200_dsyrk(char uplo, char trans, int64_t n, int64_t k, double alpha,
                 const double* A, int64_t lda, double beta, double* C, int64_t ldc,
                 elb200_stream_t s);
int elb200_ssyrk(char uplo, char trans, int64_t n, int64_t k, float alpha,
                 const float* A, int64_t lda, float beta, float* C, int64_t ldc,
                 elb200_stream_t s);
int elb200_zherk(char uplo, char trans, int64_t n, int64_t k, double alpha,
                 const elb200_c64* A, int64_t lda, double beta, elb200_c64* C, int64_t ldc,
                 elb200_stream_t s);
int elb200_cherk(char uplo, char trans, int64_t n, int64_t k, float alpha,
                 const elb200_c32* A, int64_t lda, float beta, elb200_c32* C, int64_t ldc,
                 elb200_stream_t s);
int elb200_zsyrk(char uplo, char trans, int64_t n, int64_t k, elb200_c64 alpha,
                 const elb200_c64* A, int64_t lda, elb200_c64 beta, elb200_c64* C, int64_t ldc,
                 elb200_stream_t s);
int elb200_csyrk(char uplo, char trans, int64_t n, int64_t k, elb200_c32 alpha,
                 const elb200_c32* A, int64_t lda, elb200_c32 beta, elb200_c32* C, int64_t ldc,
                 elb200_stream_t s);

/* ---- TRSM: B := alpha op(A)^-1 B (side 'L') or alpha B op(A)^-1 ('R') --- */
/* debugging aid: bit 0 = never use the fused single-launch kernel of the right-side panel solve */
void elb200_trsm_set_debug_flags(int flags);
int elb200_dtrsm(char side, char uplo, char trans, char diag, int64_t m, int64_t n,
                 double alpha, const double* A, int64_t lda, double* B, int64_t ldb,
                 elb200_stream_t s);
int elb200_strsm(char side, char uplo, char trans, char diag, int64_t m, int64_t n,
                 float alpha, const float* A, int64_t lda, float* B, int64_t ldb,
                 elb200_stream_t s);
int elb200_ztrsm(char side, char uplo, char trans, char diag, int64_t m, int64_t n,
                 elb200_c64 alpha, const elb200_c64* A, int64_t lda, elb200_c64* B, int64_t ldb,
                 elb200_stream_t s);
int elb200_ctrsm(char side, char uplo, char trans, char diag, int64_t m, int64_t n,
                 elb200_c32 alpha, const elb200_c32* A, int64_t lda, elb200_c32* B, int64_t ldb,
                 elb200_stream_t s);

/* ---- POTRF: unblocked-equivalent Cholesky of one n x n block ----------- */
/* Factors the uplo triangle in place; the other triangle is not referenced.
 * info_dev (device int, may not be NULL) is left untouched on success and
 * receives j+1 for the first column j whose pivot is <= 0 or NaN (only if it
 * still holds 0, so one flag can be shared by a whole factorisation).  The
 * host layer turns a nonzero flag into NonHPDMatrixException
 * (LowerVariant3.hpp:29-30). */
/* debugging aid: SM clocks spent by thread 0 of the single-CTA potrf kernel in {diagonal block, panel
 * solve, trailing update} and the number of launches since the last reset.  reset: 0 read only, 1 read and
 * clear, 2 read, clear and switch the profile ON (off by default), 3 read, clear and switch it off */
int elb200_potrf_phase_clocks(unsigned long long out[4], int reset);
int elb200_dpotrf(char uplo, int64_t n, double* A, int64_t lda, int* info_dev, elb200_stream_t s);
int elb200_spotrf(char uplo, int64_t n, float* A, int64_t lda, int* info_dev, elb200_stream_t s);
int elb200_zpotrf(char uplo, int64_t n, elb200_c64* A, int64_t lda, int* info_dev, elb200_stream_t s);
int elb200_cpotrf(char uplo, int64_t n, elb200_c32* A, int64_t lda, int* info_dev, elb200_stream_t s);

/* LU of a tall panel with partial pivoting <- lu::Panel (src/lapack_like/factor/LU/Panel.hpp:14-55) / lu::Unb
 * (LU/Local.hpp:44-61, pivot = 0).  A: m x n DEVICE array, m >= n, n <= 512, overwritten by unit-lower L and U;
 * ipiv: DEVICE array of n int64, ipiv[j] = row of the panel exchanged with row j (the i?amax rule: largest |x|,
 * |re| + |im| for complex, first occurrence); *info_dev (DEVICE int, must start at 0) = j + 1 of the first exactly
 * zero pivot.  One cooperative kernel: row slabs per CTA, two grid barriers per column, no host round trip. */
int elb200_dgetrf_panel(int64_t m, int64_t n, double* A, int64_t lda, int64_t* ipiv, int pivot, int* info_dev, elb200_stream_t s);
int elb200_sgetrf_panel(int64_t m, int64_t n, float* A, int64_t lda, int64_t* ipiv, int pivot, int* info_dev, elb200_stream_t s);
int elb200_zgetrf_panel(int64_t m, int64_t n, elb200_c64* A, int64_t lda, int64_t* ipiv, int pivot, int* info_dev, elb200_stream_t s);
int elb200_cgetrf_panel(int64_t m, int64_t n, elb200_c32* A, int64_t lda, int64_t* ipiv, int pivot, int* info_dev, elb200_stream_t s);

/* ---- FP64 tensor-pipe ceiling probe ------------------------------------ */
/* Runs a register-resident DMMA (mma.sync m8n8k4 f64) loop on every SM and
 * returns the achieved FLOP/s through *flops_per_s: the measured FP64 tensor
 * roofline denominator used by bench.py. */
int elb200_dmma_peak(int iters, double* flops_per_s, float* ms);

/* ---- Fortran-77 BLAS ABI (device pointers, layer's current stream) ----- */
/* Prototypes follow the reference's own declarations:
 *   src/core/imports/blas/Gemm.hpp:13-40, Trsm.hpp:12-33, Syrk.hpp:12-50 */
void sgemm_(const char* transA, const char* transB, const int* m, const int* n, const int* k,
            const float* alpha, const float* A, const int* lda, const float* B, const int* ldb,
            const float* beta, float* C, const int* ldc);
void dgemm_(const char* transA, const char* transB, const int* m, const int* n, const int* k,
            const double* alpha, const double* A, const int* lda, const double* B, const int* ldb,
            const double* beta, double* C, const int* ldc);
void cgemm_(const char* transA, const char* transB, const int* m, const int* n, const int* k,
            const elb200_c32* alpha, const elb200_c32* A, const int* lda, const elb200_c32* B,
            const int* ldb, const elb200_c32* beta, elb200_c32* C, const int* ldc);
void zgemm_(const char* transA, const char* transB, const int* m, const int* n, const int* k,
            const elb200_c64* alpha, const elb200_c64* A, const int* lda, const elb200_c64* B,
            const int* ldb, const elb200_c64* beta, elb200_c64* C, const int* ldc);
void strsm_(const char* side, const char* uplo, const char* trans, const char* diag,
            const int* m, const int* n, const float* alpha, const float* A, const int* lda,
            float* B, const int* ldb);
void dtrsm_(const char* side, const char* uplo, const char* trans, const char* diag,
            const int* m, const int* n, const double* alpha, const double* A, const int* lda,
            double* B, const int* ldb);
void ctrsm_(const char* side, const char* uplo, const char* trans, const char* diag,
            const int* m, const int* n, const elb200_c32* alpha, const elb200_c32* A,
            const int* lda, elb200_c32* B, const int* ldb);
void ztrsm_(const char* side, const char* uplo, const char* trans, const char* diag,
            const int* m, const int* n, const elb200_c64* alpha, const elb200_c64* A,
            const int* lda, elb200_c64* B, const int* ldb);
void ssyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha,
            const float* A, const int* lda, const float* beta, float* C, const int* ldc);
void dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha,
            const double* A, const int* lda, const double* beta, double* C, const int* ldc);
void csyrk_(const char* uplo, const char* trans, const int* n, const int* k,
            const elb200_c32* alpha, const elb200_c32* A, const int* lda,
            const elb200_c32* beta, elb200_c32* C, const int* ldc);
void zsyrk_(const char* uplo, const char* trans, const int* n, const int* k,
            const elb200_c64* alpha, const elb200_c64* A, const int* lda,
            const elb200_c64* beta, elb200_c64* C, const int* ldc);
void cherk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha,
            const elb200_c32* A, const int* lda, const float* beta, elb200_c32* C, const int* ldc);
void zherk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha,
            const elb200_c64* A, const int* lda, const double* beta, elb200_c64* C, const int* ldc);

/* ---- level 1 / 2 leaves of the same boundary (device pointers; kernels/level12_abi.cu) ----
 *   ?syr_ / ?her_  src/core/imports/blas/Syr.hpp:12-30,152-173 (rank-1 updates of the unblocked Cholesky)
 *   ?scal_         Scal.hpp:12-19,118-125      ?axpy_  Axpy.hpp:12-29
 *   ?lacpy_        src/core/imports/lapack.cpp:20-31,381-396 (uplo 'U' / 'L' / anything else = all) */
int elb200_sscal(int64_t n, float alpha, float* x, int64_t incx, elb200_stream_t s);
int elb200_dscal(int64_t n, double alpha, double* x, int64_t incx, elb200_stream_t s);
int elb200_cscal(int64_t n, elb200_c32 alpha, elb200_c32* x, int64_t incx, elb200_stream_t s);
int elb200_zscal(int64_t n, elb200_c64 alpha, elb200_c64* x, int64_t incx, elb200_stream_t s);
int elb200_saxpy(int64_t n, float alpha, const float* x, int64_t incx, float* y, int64_t incy, elb200_stream_t s);
int elb200_daxpy(int64_t n, double alpha, const double* x, int64_t incx, double* y, int64_t incy, elb200_stream_t s);
int elb200_caxpy(int64_t n, elb200_c32 alpha, const elb200_c32* x, int64_t incx, elb200_c32* y, int64_t incy, elb200_stream_t s);
int elb200_zaxpy(int64_t n, elb200_c64 alpha, const elb200_c64* x, int64_t incx, elb200_c64* y, int64_t incy, elb200_stream_t s);
int elb200_slacpy(char uplo, int64_t m, int64_t n, const float* A, int64_t lda, float* B, int64_t ldb, elb200_stream_t s);
int elb200_dlacpy(char uplo, int64_t m, int64_t n, const double* A, int64_t lda, double* B, int64_t ldb, elb200_stream_t s);
int elb200_clacpy(char uplo, int64_t m, int64_t n, const elb200_c32* A, int64_t lda, elb200_c32* B, int64_t ldb, elb200_stream_t s);
int elb200_zlacpy(char uplo, int64_t m, int64_t n, const elb200_c64* A, int64_t lda, elb200_c64* B, int64_t ldb, elb200_stream_t s);
int elb200_ssyr(char uplo, int64_t n, float alpha, const float* x, int64_t incx, float* A, int64_t lda, elb200_stream_t s);
int elb200_dsyr(char uplo, int64_t n, double alpha, const double* x, int64_t incx, double* A, int64_t lda, elb200_stream_t s);
int elb200_cher(char uplo, int64_t n, float alpha, const elb200_c32* x, int64_t incx, elb200_c32* A, int64_t lda, elb200_stream_t s);
int elb200_zher(char uplo, int64_t n, double alpha, const elb200_c64* x, int64_t incx, elb200_c64* A, int64_t lda, elb200_stream_t s);
void sscal_(const int* n, const float* alpha, float* x, const int* incx);
void dscal_(const int* n, const double* alpha, double* x, const int* incx);
void cscal_(const int* n, const elb200_c32* alpha, elb200_c32* x, const int* incx);
void zscal_(const int* n, const elb200_c64* alpha, elb200_c64* x, const int* incx);
void saxpy_(const int* n, const float* alpha, const float* x, const int* incx, float* y, const int* incy);
void daxpy_(const int* n, const double* alpha, const double* x, const int* incx, double* y, const int* incy);
void caxpy_(const int* n, const elb200_c32* alpha, const elb200_c32* x, const int* incx, elb200_c32* y, const int* incy);
void zaxpy_(const int* n, const elb200_c64* alpha, const elb200_c64* x, const int* incx, elb200_c64* y, const int* incy);
void slacpy_(const char* uplo, const int* m, const int* n, const float* A, const int* lda, float* B, const int* ldb);
void dlacpy_(const char* uplo, const int* m, const int* n, const double* A, const int* lda, double* B, const int* ldb);
void clacpy_(const char* uplo, const int* m, const int* n, const elb200_c32* A, const int* lda, elb200_c32* B, const int* ldb);
void zlacpy_(const char* uplo, const int* m, const int* n, const elb200_c64* A, const int* lda, elb200_c64* B, const int* ldb);
void ssyr_(const char* uplo, const int* n, const float* alpha, const float* x, const int* incx, float* A, const int* lda);
void dsyr_(const char* uplo, const int* n, const double* alpha, const double* x, const int* incx, double* A, const int* lda);
void cher_(const char* uplo, const int* n, const float* alpha, const elb200_c32* x, const int* incx, elb200_c32* A, const int* lda);
void zher_(const char* uplo, const int* n, const double* alpha, const elb200_c64* x, const int* incx, elb200_c64* A, const int* lda);

#ifdef __cplusplus
}
#endif
#endif /* ELB200_BLAS_H */
