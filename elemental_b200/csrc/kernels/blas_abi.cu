// Fortran-77 BLAS ABI over the device kernels: exactly the symbols the reference's
// El::blas wrappers bind (src/core/imports/blas/Gemm.hpp:13-40, Trsm.hpp:12-33,
// Syrk.hpp:12-50), taking DEVICE pointers and launching on the layer's current
// stream (elb200_set_stream).  Errors cannot be returned through this ABI (BLAS has
// xerbla); they are reported on stderr and left in elb200_last_error().
#include "device_api.hpp"
#include "elb200_blas.h"

namespace {
void report(int rc, const char* name) {
    if (rc != 0) fprintf(stderr, "elb200 %s: %s\n", name, elb200_last_error());
    elb200::fortran_abi_fence();
}
elb200_stream_t cur() { return (elb200_stream_t)elb200::current_stream(); }
}  // namespace

extern "C" {
void sgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const float* alpha,
            const float* A, const int* lda, const float* B, const int* ldb, const float* beta, float* C,
            const int* ldc) {
    report(elb200_sgemm(*ta, *tb, *m, *n, *k, *alpha, A, *lda, B, *ldb, *beta, C, *ldc, cur()), "sgemm_");
}
void dgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const double* alpha,
            const double* A, const int* lda, const double* B, const int* ldb, const double* beta, double* C,
            const int* ldc) {
    report(elb200_dgemm(*ta, *tb, *m, *n, *k, *alpha, A, *lda, B, *ldb, *beta, C, *ldc, cur()), "dgemm_");
}
void cgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const elb200_c32* alpha,
            const elb200_c32* A, const int* lda, const elb200_c32* B, const int* ldb, const elb200_c32* beta,
            elb200_c32* C, const int* ldc) {
    report(elb200_cgemm(*ta, *tb, *m, *n, *k, *alpha, A, *lda, B, *ldb, *beta, C, *ldc, cur()), "cgemm_");
}
void zgemm_(const char* ta, const char* tb, const int* m, const int* n, const int* k, const elb200_c64* alpha,
            const elb200_c64* A, const int* lda, const elb200_c64* B, const int* ldb, const elb200_c64* beta,
            elb200_c64* C, const int* ldc) {
    report(elb200_zgemm(*ta, *tb, *m, *n, *k, *alpha, A, *lda, B, *ldb, *beta, C, *ldc, cur()), "zgemm_");
}
void strsm_(const char* side, const char* uplo, const char* trans, const char* diag, const int* m, const int* n,
            const float* alpha, const float* A, const int* lda, float* B, const int* ldb) {
    report(elb200_strsm(*side, *uplo, *trans, *diag, *m, *n, *alpha, A, *lda, B, *ldb, cur()), "strsm_");
}
void dtrsm_(const char* side, const char* uplo, const char* trans, const char* diag, const int* m, const int* n,
            const double* alpha, const double* A, const int* lda, double* B, const int* ldb) {
    report(elb200_dtrsm(*side, *uplo, *trans, *diag, *m, *n, *alpha, A, *lda, B, *ldb, cur()), "dtrsm_");
}
void ctrsm_(const char* side, const char* uplo, const char* trans, const char* diag, const int* m, const int* n,
            const elb200_c32* alpha, const elb200_c32* A, const int* lda, elb200_c32* B, const int* ldb) {
    report(elb200_ctrsm(*side, *uplo, *trans, *diag, *m, *n, *alpha, A, *lda, B, *ldb, cur()), "ctrsm_");
}
void ztrsm_(const char* side, const char* uplo, const char* trans, const char* diag, const int* m, const int* n,
            const elb200_c64* alpha, const elb200_c64* A, const int* lda, elb200_c64* B, const int* ldb) {
    report(elb200_ztrsm(*side, *uplo, *trans, *diag, *m, *n, *alpha, A, *lda, B, *ldb, cur()), "ztrsm_");
}
void ssyrk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha, const float* A,
            const int* lda, const float* beta, float* C, const int* ldc) {
    report(elb200_ssyrk(*uplo, *trans, *n, *k, *alpha, A, *lda, *beta, C, *ldc, cur()), "ssyrk_");
}
void dsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha,
            const double* A, const int* lda, const double* beta, double* C, const int* ldc) {
    report(elb200_dsyrk(*uplo, *trans, *n, *k, *alpha, A, *lda, *beta, C, *ldc, cur()), "dsyrk_");
}
void csyrk_(const char* uplo, const char* trans, const int* n, const int* k, const elb200_c32* alpha,
            const elb200_c32* A, const int* lda, const elb200_c32* beta, elb200_c32* C, const int* ldc) {
    report(elb200_csyrk(*uplo, *trans, *n, *k, *alpha, A, *lda, *beta, C, *ldc, cur()), "csyrk_");
}
void zsyrk_(const char* uplo, const char* trans, const int* n, const int* k, const elb200_c64* alpha,
            const elb200_c64* A, const int* lda, const elb200_c64* beta, elb200_c64* C, const int* ldc) {
    report(elb200_zsyrk(*uplo, *trans, *n, *k, *alpha, A, *lda, *beta, C, *ldc, cur()), "zsyrk_");
}
void cherk_(const char* uplo, const char* trans, const int* n, const int* k, const float* alpha,
            const elb200_c32* A, const int* lda, const float* beta, elb200_c32* C, const int* ldc) {
    report(elb200_cherk(*uplo, *trans, *n, *k, *alpha, A, *lda, *beta, C, *ldc, cur()), "cherk_");
}
void zherk_(const char* uplo, const char* trans, const int* n, const int* k, const double* alpha,
            const elb200_c64* A, const int* lda, const double* beta, elb200_c64* C, const int* ldc) {
    report(elb200_zherk(*uplo, *trans, *n, *k, *alpha, A, *lda, *beta, C, *ldc, cur()), "zherk_");
}
}
