// Complex<double> GEMM / TRRK on the REAL persistent kernel (gemm_f64_ws.cu).
//
// Column-major complex storage of an m x n matrix IS a real (2m) x n matrix whose rows 2i, 2i + 1 hold (re, im) of
// complex row i.  With a = op(A)(i,l), b = op(B)(l,j):
//     [ Re(ab) ]   [ ar  -ai ] [ br ]
//     [ Im(ab) ] = [ ai   ar ] [ bi ]
// so C_real (2m x n) = Ahat (2m x 2k) * Bhat (2k x n), where column 2l of Ahat is column l of op(A) (as re / im
// pairs) and column 2l + 1 is column l of i * op(A), and Bhat is op(B) with its (re, im) pairs along k -- for
// op(B) = B that is B's own storage, no copy.  The product has 2 (2m) n (2k) = 8 mnk real flops: exactly the complex
// product's, none wasted, all of them on the DMMA pipe at the real kernel's rate (the dedicated complex kernel of
// gemm_c64.cu spends four DMMAs per complex MMA on a cp.async pipeline: 24-31 TFLOP/s against 34).
// What it costs: Ahat is built by one O(mk) pass (alpha and any transposition / conjugation folded in), a
// transposed or conjugated B by one O(kn) pass -- HBM-bound, negligible beside O(mnk) for the rank-Blocksize()
// updates of Cholesky / Trsm / SUMMA (m x 128 panel: ~50 us against ~30 ms).  k is processed in chunks so that the
// scratch stays within 512 MB.  The staircase mask of TRRK runs on complex rows (row >> 1) in the kernel.
//   replaces blas::Gemm<Complex<double>> -> zgemm_ (src/core/imports/blas/Gemm.hpp:511) and the LocalTrrk
//   recursion (src/blas_like/level3/Trrk/Local.hpp:782-830) for Complex<double>.
#include "../common.hpp"
#include "cplx.cuh"
#include "device_api.hpp"

namespace elb200 {
bool dgemm_ws_device(int mode, bool ta, bool tb, i64 m, i64 n, i64 k, double alpha, const double* A, i64 lda,
                     const double* B, i64 ldb, double beta, double* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                     double flops, cudaStream_t s, int rowPair);
namespace {

// Ahat(:, 2l) = alpha * X(:, l), Ahat(:, 2l + 1) = i * alpha * X(:, l); X is m x kc complex (ld ldx), Ahat 2m x 2kc real
__global__ void __launch_bounds__(256) expand_left_kernel(i64 m, i64 kc, c64_t alpha, int has_alpha, const c64_t* __restrict__ X,
                                                          i64 ldx, double* __restrict__ out) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i >= m) return;
    for (i64 l = blockIdx.y; l < kc; l += gridDim.y) {
        c64_t a = X[i + l * ldx];
        if (has_alpha) a = alpha * a;
        double2* c0 = reinterpret_cast<double2*>(out + 2 * i + (2 * l) * (2 * m));
        double2* c1 = reinterpret_cast<double2*>(out + 2 * i + (2 * l + 1) * (2 * m));
        *c0 = make_double2(a.re, a.im);
        *c1 = make_double2(-a.im, a.re);
    }
}
// Im(C(i,i)) := 0 on the global diagonal (HERK semantics)
__global__ void __launch_bounds__(256) real_diag_kernel(i64 m, i64 n, c64_t* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs) {
    const i64 j = (i64)blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const i64 gj = gj0 + j * gjs;
    // the local row with gi == gj, if this process owns it
    const i64 d = gj - gi0;
    if (d < 0 || d % gis != 0) return;
    const i64 i = d / gis;
    if (i < m) C[i + j * ldc].im = 0.0;
}

}  // namespace

// False (nothing done) when the product is too small for the extra passes to pay; the caller then uses gemm_c64.cu.
bool zgemm_real_device(int mode, int ta, int tb, i64 m, i64 n, i64 k, c64_t alpha, const c64_t* A, i64 lda,
                       const c64_t* B, i64 ldb, c64_t beta, c64_t* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                       bool realDiag, cudaStream_t s) {
    if (m <= 0 || n <= 0 || k <= 0) return false;
    if ((double)m * (double)n * (double)k < 64.0 * 64.0 * 64.0 * 8.0) return false;
    if (alpha.re == 0.0 && alpha.im == 0.0) return false;
    if (2 * m >= (i64(1) << 31) - 256 || 2 * k >= (i64(1) << 31) - 256) return false;
    // beta: real values go to the kernel as they are (they scale re and im alike); a complex beta scales C first
    double betaR = beta.re;
    if (beta.im != 0.0) {
        if (mode != 0) return false;   // a masked scale would be needed: leave it to the complex kernel
        lattice_copy_device<c64_t>(C, C, m, n, 0, 1, ldc, 0, 1, ldc, false, &beta, false, s);
        betaR = 1.0;
    }
    const bool alphaOne = (alpha.re == 1.0 && alpha.im == 0.0);
    // scratch: Ahat chunk (2m x 2kc doubles), op(A) chunk when A is transposed, Bhat when B is transposed
    i64 kc = k;
    const i64 maxA = (i64(512) << 20) / (i64)(32 * m);   // 32 bytes of Ahat per (row, summation index)
    if (kc > maxA) kc = maxA < 16 ? 16 : (maxA / 16) * 16;
    double* Ahat = (double*)scratch_alloc(sizeof(double) * (size_t)(4 * m * kc), s);
    c64_t* At = ta ? (c64_t*)scratch_alloc(sizeof(c64_t) * (size_t)(m * kc), s) : nullptr;
    c64_t* Bt = tb ? (c64_t*)scratch_alloc(sizeof(c64_t) * (size_t)(kc * n), s) : nullptr;
    for (i64 k0 = 0; k0 < k; k0 += kc) {
        const i64 kw = (k - k0 < kc) ? (k - k0) : kc;
        // op(A)(:, k0 : k0 + kw) as an m x kw column-major block X
        const c64_t* X = A + k0 * lda;
        i64 ldx = lda;
        if (ta) {   // A stored k x m: transpose (and conjugate) the rows k0.. into At
            lattice_copy_device<c64_t>(A, At, m, kw, k0, lda, 1, 0, 1, m, ta == 2, nullptr, false, s);
            X = At; ldx = m;
        }
        dim3 grid((unsigned)ceil_div(m, 256), (unsigned)(kw < 1024 ? kw : 1024));
        expand_left_kernel<<<grid, 256, 0, s>>>(m, kw, alpha, alphaOne ? 0 : 1, X, ldx, Ahat);
        ELB_LAUNCH_CHECK();
        // op(B)(k0 : k0 + kw, :) with its (re, im) pairs along k: B itself, or a transposed copy
        const double* Bhat = reinterpret_cast<const double*>(B + k0);
        i64 ldbh = 2 * ldb;
        if (tb) {   // B stored n x k
            lattice_copy_device<c64_t>(B, Bt, kw, n, k0 * ldb, ldb, 1, 0, 1, kw, tb == 2, nullptr, false, s);
            Bhat = reinterpret_cast<const double*>(Bt); ldbh = 2 * kw;
        }
        const double flops = 8.0 * (double)m * (double)n * (double)kw * (mode == 0 ? 1.0 : 0.5);
        if (!dgemm_ws_device(mode, false, false, 2 * m, n, 2 * kw, 1.0, Ahat, 2 * m, Bhat, ldbh, k0 == 0 ? betaR : 1.0,
                             reinterpret_cast<double*>(C), 2 * ldc, gi0, gis, gj0, gjs, flops, s, 1))
            throw std::logic_error("zgemm: the real kernel rejected operands it was handed aligned");
    }
    if (realDiag) {
        real_diag_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, s>>>(m, n, C, ldc, gi0, gis, gj0, gjs);
        ELB_LAUNCH_CHECK();
    }
    scratch_free(Ahat, s);
    if (At) scratch_free(At, s);
    if (Bt) scratch_free(Bt, s);
    return true;
}

}  // namespace elb200
