// Internal typed entry points shared by the kernel translation units and the
// C++ host layer.  T is one of float, double, cplx<float>, cplx<double>.
#pragma once
#include "../common.hpp"
#include "cplx.cuh"

namespace elb200 {

// mode 0 = full GEMM, 1 = lower-triangle TRRK, 2 = upper-triangle TRRK.
// Global index of local C(i,j) is (gi0 + i*gis, gj0 + j*gjs) (TRRK only).
// transA/transB in {'N','T','C'}.
template <class T>
void gemm_device(int mode, char transA, char transB, i64 m, i64 n, i64 k, T alpha, const T* A,
                 i64 lda, const T* B, i64 ldb, T beta, T* C, i64 ldc, i64 gi0, i64 gis, i64 gj0,
                 i64 gjs, cudaStream_t s);

// FP64 DMMA kernel (gemm_f64.cu)
void dgemm_device(int mode, char transA, char transB, i64 m, i64 n, i64 k, double alpha,
                  const double* A, i64 lda, const double* B, i64 ldb, double beta, double* C,
                  i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs, cudaStream_t s);
// Complex<double> DMMA kernel (gemm_c64.cu)
void zgemm_device(int mode, char transA, char transB, i64 m, i64 n, i64 k, c64_t alpha,
                  const c64_t* A, i64 lda, const c64_t* B, i64 ldb, c64_t beta, c64_t* C, i64 ldc,
                  i64 gi0, i64 gis, i64 gj0, i64 gjs, cudaStream_t s);
// same, optionally forcing a real diagonal (HERK semantics: Im(c_ii) := 0)
void zgemm_device_ex(int mode, char transA, char transB, i64 m, i64 n, i64 k, c64_t alpha,
                     const c64_t* A, i64 lda, const c64_t* B, i64 ldb, c64_t beta, c64_t* C, i64 ldc,
                     i64 gi0, i64 gis, i64 gj0, i64 gjs, bool realDiag, cudaStream_t s);
// Generic SIMT kernel, any type (gemm_simt.cu)
template <class T>
void gemm_simt_device(int mode, char transA, char transB, i64 m, i64 n, i64 k, T alpha, const T* A,
                      i64 lda, const T* B, i64 ldb, T beta, T* C, i64 ldc, i64 gi0, i64 gis,
                      i64 gj0, i64 gjs, cudaStream_t s, bool realDiag = false);

// float on the tensor cores: 3xTF32 split, tcgen05.mma kind::tf32, accumulator in TMEM (gemm_tf32.cu).
// Returns false, launching nothing, when A/B are not 16-byte aligned with ld % 4 == 0.
bool sgemm_3xtf32_device(char transA, char transB, i64 m, i64 n, i64 k, float alpha, const float* A, i64 lda,
                         const float* B, i64 ldb, float beta, float* C, i64 ldc, cudaStream_t s);
int sgemm_mode();        // 0 exact FFMA, 1 3xTF32 (elb200_sgemm_set_mode)
void sgemm_note_simt();

template <class T>
void trsm_device(char side, char uplo, char trans, char diag, i64 m, i64 n, T alpha, const T* A,
                 i64 lda, T* B, i64 ldb, cudaStream_t s);

// info_dev: device int, set to col_offset + j + 1 on the first non-positive pivot (if still 0)
template <class T>
void potrf_device(char uplo, i64 n, T* A, i64 lda, int* info_dev, i64 col_offset, cudaStream_t s);

// LU panel and row interchanges (lu.cu).  getrf_panel: in-place LU of an m x n panel (m >= n <= 512) by one cooperative
// kernel; ipiv[j] (device) = row of the panel exchanged with row j; info_dev set to col0 + j + 1 at the first zero pivot.
template <class T>
void getrf_panel_device(i64 m, i64 n, T* A, i64 lda, i64* ipiv, bool pivot, int* info_dev, i64 col0, cudaStream_t s);
// slots of the rows a panel's swaps touch: slotRow[2 nb] (global row or -1), srcSlot[2 nb] (whose original row lands here)
void swap_plan_device(int nb, const i64* ipiv, i64 k, i64* slotRow, int* srcSlot, cudaStream_t s);
template <class T>
void pack_rows_device(int S, const i64* slotRow, const int* srcSlot, const T* A, i64 lda, i64 ncols, int align, int stride,
                      int rank, int shift, T* buf, cudaStream_t s);
template <class T>
void unpack_rows_device(int S, const i64* slotRow, const int* srcSlot, T* A, i64 lda, i64 ncols, int align, int stride,
                        int rank, int shift, const T* all, i64 perRank, cudaStream_t s);
// dst(iLoc, c) := src(perm[shift + iLoc stride], c) (rowwise) or dst(i, cLoc) := src(i, perm[shift + cLoc stride])
template <class T>
void permute_device(bool rowwise, i64 mloc, i64 nloc, const i64* perm, int shift, int stride, const T* src, i64 lds, T* dst,
                    i64 ldd, cudaStream_t s);
// origins[at + j] = offset + j, dests[at + j] = offset + ipiv[j]
void append_swaps_device(i64* origins, i64* dests, i64 at, const i64* ipiv, int count, i64 offset, cudaStream_t s);

// rank-w update (downdate) of the M x nb panel L of a Cholesky factor whose first row is the diagonal row of its
// first column, with V (M x w, same rows) carrying the state between panels (cholmod.cu); *info_dev := 1 when a
// downdate would make the matrix indefinite
template <class T>
void cholmod_panel_device(bool downdate, i64 M, i64 nb, T* L, i64 ldl, T* V, i64 ldv, i64 w, int* info_dev, cudaStream_t s);

// diagonally pivoted Cholesky (cholpiv.cu): replicated panel state, plain local kernels
template <class T>
void diag_extract_device(i64 mloc, i64 nloc, const T* A, i64 lda, int colShift, int colStride, int rowShift, int rowStride,
                         double* d, cudaStream_t s);                       // d[i] = Re A(i,i) for the owned diagonal entries
void argmax_abs_device(const double* d, i64 lo, i64 hi, i64* out, cudaStream_t s);   // first index of max |d[lo:hi)|
template <class T>
void pivot_swap_device(int k, i64 f, double* d, T* X, i64 ldx, i64* pos, i64* ipiv, cudaStream_t s);
template <class T>
void pivot_column_device(i64 M, int k, const T* h, const i64* pos, T* X, i64 ldx, double* d, int* info, i64 col, cudaStream_t s);
template <class T>
void pack_cols_device(int S, const i64* slotCol, const int* srcSlot, const T* A, i64 lda, i64 mloc, int align, int stride,
                      int rank, int shift, T* buf, cudaStream_t s);
template <class T>
void unpack_cols_device(int S, const i64* slotCol, const int* srcSlot, T* A, i64 lda, i64 mloc, int align, int stride,
                        int rank, int shift, const T* all, i64 perRank, cudaStream_t s);

// dst (op)= alpha*op(src) on one strided block (level1.cu)
template <class T>
void lattice_copy_device(const T* src, T* dst, i64 nrows, i64 ncols, i64 s_off, i64 s_rs, i64 s_cs,
                         i64 d_off, i64 d_rs, i64 d_cs, bool conj, const T* alpha, bool accumulate,
                         cudaStream_t s);

// Stream-ordered scratch memory (cudaMallocAsync on the device's default pool,
// release threshold raised so repeated panel-sized requests are served from cache)
void* scratch_alloc(size_t bytes, cudaStream_t s);

// flag operations of the peer-memory redistribution path (kernels/p2p.cu): store `epoch` (release, system
// scope) into every signal[] address, then spin until every wait[] address holds a value >= wait_value
constexpr int P2P_MAX_PEERS = 64;
struct P2PFlagOps {
    unsigned* signal[P2P_MAX_PEERS];
    const unsigned* wait[P2P_MAX_PEERS];
    int nsignal, nwait;
    unsigned epoch, wait_value;
    int* error;  // pinned host memory: set when a wait times out
};
void p2p_flags(const P2PFlagOps& ops, cudaStream_t s);
int p2p_flag_mode();  // 1: stream memory operations (no kernel), 0: one-warp flag kernel
void scratch_free(void* p, cudaStream_t s);

template <class T> struct dtype_code;
template <> struct dtype_code<float> { static constexpr int value = 0; };
template <> struct dtype_code<double> { static constexpr int value = 1; };
template <> struct dtype_code<c32_t> { static constexpr int value = 2; };
template <> struct dtype_code<c64_t> { static constexpr int value = 3; };

}  // namespace elb200
