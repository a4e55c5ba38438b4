// Complex<double> GEMM / TRRK entry.  TEMPORARY: forwards to the generic SIMT kernel until
// the DMMA (4 real m8n8k4 MMAs per complex tile) kernel lands in this file.
#include "device_api.hpp"
namespace elb200 {
void zgemm_device(int mode, char ta, char tb, i64 m, i64 n, i64 k, c64_t alpha, const c64_t* A, i64 lda,
                  const c64_t* B, i64 ldb, c64_t beta, c64_t* C, i64 ldc, i64 gi0, i64 gis, i64 gj0,
                  i64 gjs, cudaStream_t s) {
    gemm_simt_device<c64_t>(mode, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs, s);
}
}  // namespace elb200
