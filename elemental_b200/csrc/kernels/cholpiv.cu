// Kernels of the diagonally pivoted Cholesky factorisation (host/cholesky_ext.cpp):
//   diag_extract / argmax_abs  <- pivot::Full / pivot::PanelFull (Cholesky/PivotedLowerVariant3.hpp:17-104):
//                                 the largest remaining (lazily updated) diagonal entry, first occurrence
//   pivot_swap                 <- the RowSwap calls on d, X, Y of PivotedLowerPanel (:262-267)
//   pivot_column               <- :269-283: a(k:, k) -= X(k:, 0:k) Y(k, 0:k)^T, scale by 1 / sqrt(alpha11), and the
//                                 diagonal update the reference defers to the next pivot search
//   pack_cols / unpack_cols    <- the column half of the symmetric interchange (HermitianSwap, :257)
// All state of a panel (X, d, the position map) is replicated, so these are plain local kernels.
#include "device_api.hpp"

namespace elb200 {
namespace {

template <class T> struct re_of { typedef T type; };
template <class R> struct re_of<cplx<R>> { typedef R type; };

// d[i] := Re A(i,i) for the diagonal entries this process owns (d must be zeroed: the others add nothing)
template <class T>
__global__ void __launch_bounds__(256) diag_extract_kernel(i64 mloc, i64 nloc, const T* __restrict__ A, i64 lda, int colShift,
                                                           int colStride, int rowShift, int rowStride, double* d) {
    const i64 iLoc = (i64)blockIdx.x * 256 + threadIdx.x;
    if (iLoc >= mloc) return;
    const i64 i = colShift + iLoc * colStride;
    if ((i - rowShift) % rowStride != 0 || i < rowShift) return;
    const i64 jLoc = (i - rowShift) / rowStride;
    if (jLoc >= nloc) return;
    d[i] = (double)scalar_traits<T>::real_part(A[iLoc + jLoc * lda]);
}

// out[0] := lo + argmax_i |d[lo + i]|, first occurrence
__global__ void __launch_bounds__(1024) argmax_abs_kernel(const double* __restrict__ d, i64 lo, i64 hi, i64* out) {
    __shared__ double sv[32];
    __shared__ long long si[32];
    double bv = -1.0;
    long long bi = 0x7fffffffffffffffLL;
    for (i64 i = lo + threadIdx.x; i < hi; i += 1024) {
        const double v = fabs(d[i]);
        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
    }
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, bv, o);
        const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        bv = sv[threadIdx.x]; bi = si[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, bv, o);
            const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (threadIdx.x == 0) out[0] = bi;
    }
}

// positions k and f (relative to the panel's first row) trade places: d, the rows of X computed so far, the map
template <class T>
__global__ void __launch_bounds__(256) pivot_swap_kernel(int k, i64 f, double* d, T* X, i64 ldx, i64* pos, i64* ipiv) {
    const int t = threadIdx.x;
    if (t == 0) ipiv[k] = f;
    if (f == k) return;
    for (int c = t; c < k; c += 256) {
        const T a = X[k + (i64)c * ldx];
        X[k + (i64)c * ldx] = X[f + (i64)c * ldx];
        X[f + (i64)c * ldx] = a;
    }
    if (t == 0) {
        const double a = d[k]; d[k] = d[f]; d[f] = a;
        const i64 p = pos[k]; pos[k] = pos[f]; pos[f] = p;
    }
}

// column k of the panel from the gathered column h of the matrix as it stood at the start of the panel:
//   a_i = h[pos[i]] - sum_{c<k} X(i,c) conj(X(k,c)),  X(k,k) = sqrt(a_k),  X(i,k) = a_i / X(k,k),  d_i -= |X(i,k)|^2
template <class T>
__global__ void __launch_bounds__(256) pivot_column_kernel(i64 M, int k, const T* __restrict__ h, const i64* __restrict__ pos,
                                                           T* X, i64 ldx, double* d, int* info, i64 col) {
    typedef typename re_of<T>::type R;
    typedef scalar_traits<T> st;
    __shared__ T yk[512];     // conj(X(k, 0:k))
    __shared__ R sPiv;
    const int t = threadIdx.x;
    for (int c = t; c < k; c += 256) yk[c] = st::conj(X[k + (i64)c * ldx]);
    __syncthreads();
    if (t == 0) {
        T a = h[pos[k]];
        for (int c = 0; c < k; ++c) a = a - X[k + (i64)c * ldx] * yk[c];
        R ar = st::real_part(a);
        if (!(ar > R(0))) { if (blockIdx.x == 0) atomicCAS(info, 0, (int)(col + 1)); ar = R(1); }
        sPiv = sqrt(ar);
    }
    __syncthreads();
    const R delta = sPiv, inv = R(1) / delta;
    const i64 i = (i64)blockIdx.x * 256 + t;
    if (i >= M) return;
    T* xk = X + (i64)k * ldx;
    if (i < k) { xk[i] = st::zero(); return; }
    if (i == k) { xk[i] = st::from_real(delta); d[i] = 0.0; return; }
    T a = h[pos[i]];
    for (int c = 0; c < k; ++c) a = a - X[i + (i64)c * ldx] * yk[c];
    a = a * inv;
    xk[i] = a;
    d[i] -= (double)st::abs2(a);
}

// column interchange: buf[i + mloc * slot] := A(i, localCol(slotCol[slot])) for the slots whose column this process owns
template <class T>
__global__ void __launch_bounds__(256) pack_cols_kernel(int S, const i64* __restrict__ slotCol, const int* __restrict__ srcSlot,
                                                        const T* __restrict__ A, i64 lda, i64 mloc, int align, int stride,
                                                        int rank, int shift, T* buf) {
    const int slot = blockIdx.y;
    const i64 col = slotCol[slot];
    if (col < 0 || (int)((col + align) % stride) != rank) return;
    const i64 jLoc = (col - shift) / stride;
    for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < mloc; i += (i64)gridDim.x * 256) buf[i + mloc * slot] = A[i + jLoc * lda];
}
template <class T>
__global__ void __launch_bounds__(256) unpack_cols_kernel(int S, const i64* __restrict__ slotCol, const int* __restrict__ srcSlot,
                                                          T* A, i64 lda, i64 mloc, int align, int stride, int rank, int shift,
                                                          const T* __restrict__ all, i64 perRank) {
    const int slot = blockIdx.y;
    const i64 col = slotCol[slot];
    if (col < 0 || (int)((col + align) % stride) != rank) return;
    const int src = srcSlot[slot];
    if (src == slot) return;
    const int owner = (int)((slotCol[src] + align) % stride);
    const T* from = all + (i64)owner * perRank + mloc * src;
    const i64 jLoc = (col - shift) / stride;
    for (i64 i = (i64)blockIdx.x * 256 + threadIdx.x; i < mloc; i += (i64)gridDim.x * 256) A[i + jLoc * lda] = from[i];
}

}  // namespace

template <class T>
void diag_extract_device(i64 mloc, i64 nloc, const T* A, i64 lda, int colShift, int colStride, int rowShift, int rowStride,
                         double* d, cudaStream_t s) {
    if (mloc <= 0 || nloc <= 0) return;
    diag_extract_kernel<T><<<(unsigned)ceil_div(mloc, 256), 256, 0, s>>>(mloc, nloc, A, lda, colShift, colStride, rowShift,
                                                                         rowStride, d);
    ELB_LAUNCH_CHECK();
}
void argmax_abs_device(const double* d, i64 lo, i64 hi, i64* out, cudaStream_t s) {
    argmax_abs_kernel<<<1, 1024, 0, s>>>(d, lo, hi, out);
    ELB_LAUNCH_CHECK();
}
template <class T>
void pivot_swap_device(int k, i64 f, double* d, T* X, i64 ldx, i64* pos, i64* ipiv, cudaStream_t s) {
    pivot_swap_kernel<T><<<1, 256, 0, s>>>(k, f, d, X, ldx, pos, ipiv);
    ELB_LAUNCH_CHECK();
}
template <class T>
void pivot_column_device(i64 M, int k, const T* h, const i64* pos, T* X, i64 ldx, double* d, int* info, i64 col, cudaStream_t s) {
    if (M <= 0) return;
    if (k >= 512) throw std::logic_error("pivot_column: panels wider than 512 columns are not supported");
    pivot_column_kernel<T><<<(unsigned)ceil_div(M, 256), 256, 0, s>>>(M, k, h, pos, X, ldx, d, info, col);
    ELB_LAUNCH_CHECK();
}
template <class T>
void pack_cols_device(int S, const i64* slotCol, const int* srcSlot, const T* A, i64 lda, i64 mloc, int align, int stride,
                      int rank, int shift, T* buf, cudaStream_t s) {
    if (S <= 0 || mloc <= 0) return;
    i64 gx = ceil_div(mloc, 256);
    if (gx > 64) gx = 64;
    pack_cols_kernel<T><<<dim3((unsigned)gx, (unsigned)S), 256, 0, s>>>(S, slotCol, srcSlot, A, lda, mloc, align, stride, rank, shift, buf);
    ELB_LAUNCH_CHECK();
}
template <class T>
void unpack_cols_device(int S, const i64* slotCol, const int* srcSlot, T* A, i64 lda, i64 mloc, int align, int stride,
                        int rank, int shift, const T* all, i64 perRank, cudaStream_t s) {
    if (S <= 0 || mloc <= 0) return;
    i64 gx = ceil_div(mloc, 256);
    if (gx > 64) gx = 64;
    unpack_cols_kernel<T><<<dim3((unsigned)gx, (unsigned)S), 256, 0, s>>>(S, slotCol, srcSlot, A, lda, mloc, align, stride, rank, shift,
                                                                           all, perRank);
    ELB_LAUNCH_CHECK();
}

#define ELB_CP_INST(T)                                                                                                   \
    template void diag_extract_device<T>(i64, i64, const T*, i64, int, int, int, int, double*, cudaStream_t);            \
    template void pivot_swap_device<T>(int, i64, double*, T*, i64, i64*, i64*, cudaStream_t);                            \
    template void pivot_column_device<T>(i64, int, const T*, const i64*, T*, i64, double*, int*, i64, cudaStream_t);     \
    template void pack_cols_device<T>(int, const i64*, const int*, const T*, i64, i64, int, int, int, int, T*, cudaStream_t); \
    template void unpack_cols_device<T>(int, const i64*, const int*, T*, i64, i64, int, int, int, int, const T*, i64, cudaStream_t);
ELB_CP_INST(float)
ELB_CP_INST(double)
ELB_CP_INST(c32_t)
ELB_CP_INST(c64_t)

}  // namespace elb200
