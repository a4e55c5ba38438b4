// Exact FP32 GEMM / TRRK on the SIMT FFMA pipe: the default arithmetic of El::Gemm<float> (the 3xTF32 tcgen05
// kernel of gemm_tf32.cu is opt-in because its error model differs).  Replaces blas::Gemm<float> -> sgemm_
// (reference src/core/imports/blas/Gemm.hpp:388) and the float LocalTrrk recursion (Trrk/Local.hpp:782-830).
//
// Bound: FP32 FFMA issue -- 148 SMs x 128 lanes x 2 flop x 1.965 GHz = 74.5 TFLOP/s.  The generic kernel of
// gemm_simt.cu (one code path for four scalar types, 4 x 4 outputs per thread, scalar shared-memory loads) reached
// 16 TFLOP/s = 22 % of that.  This one is the classical register-tiled form:
//   * CTA tile 128 x 128, k-slab 8, 256 threads, two CTAs per SM; each thread owns an 8 x 8 block of C as four
//     4 x 4 quadrants (rows ty*4.. and 64 + ty*4.., columns tx*4.. and 64 + tx*4..): 64 FFMA per k for four
//     LDS.128, i.e. 6 shared-memory wavefronts per warp and k (A: 16 distinct 16-byte chunks, B: 2, the rest
//     is broadcast);
//   * shared tiles are k-major rows As[k][128 + 4], Bs[k][128 + 4] (the pad makes the transposing stores of a
//     k-contiguous operand conflict-free); global loads are one LDG.128 per thread and operand along whichever
//     dimension is contiguous, staged through registers one slab ahead of the FFMAs (double-buffered tiles, one
//     barrier per slab);
//   * C is column-major and the row index is the fastest thread index, so every store instruction writes 256-byte
//     runs of two columns.
// Needs 16-byte aligned A, B with leading dimensions that are multiples of 4; anything else (and Complex<float>)
// stays on gemm_simt.cu.
#include "../common.hpp"
#include "device_api.hpp"

namespace elb200 {
namespace {

constexpr int TM = 128, TN = 128, TK = 8;
constexpr int PITCH = TM + 4;
constexpr int NT = 256;

struct FArgs {
    i64 m, n, k;
    const float* A; i64 lda;
    const float* B; i64 ldb;
    float* C; i64 ldc;
    float alpha, beta;
    i64 gi0, gis, gj0, gjs;
    i64 tilesM, tilesN;
};

// one 128 x 8 (rows x k) slab of op(X) starting at (r0, k0): thread t fetches 4 consecutive elements along the
// contiguous dimension.  ROWS_CONTIG: X stored rows x k (leading dimension ld) -> 4 rows at one k;
// else X stored k x rows -> 4 k at one row.
template <bool ROWS_CONTIG>
__device__ __forceinline__ float4 load_slab(const float* __restrict__ X, i64 ld, i64 r0, i64 k0, i64 rows, i64 kk, int t) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ROWS_CONTIG) {
        const int kq = t >> 5, rq = (t & 31) * 4;          // 8 k x 32 row-quads
        const i64 r = r0 + rq, k = k0 + kq;
        if (k < kk) {
            const float* p = X + r + k * ld;
            if (r + 3 < rows) v = *reinterpret_cast<const float4*>(p);
            else {
                if (r < rows) v.x = p[0];
                if (r + 1 < rows) v.y = p[1];
                if (r + 2 < rows) v.z = p[2];
            }
        }
    } else {
        const int rr = t >> 1, kq = (t & 1) * 4;            // 128 rows x 2 k-quads
        const i64 r = r0 + rr, k = k0 + kq;
        if (r < rows) {
            const float* p = X + k + r * ld;
            if (k + 3 < kk) v = *reinterpret_cast<const float4*>(p);
            else {
                if (k < kk) v.x = p[0];
                if (k + 1 < kk) v.y = p[1];
                if (k + 2 < kk) v.z = p[2];
            }
        }
    }
    return v;
}
template <bool ROWS_CONTIG>
__device__ __forceinline__ void store_slab(float (*S)[PITCH], float4 v, int t) {
    if (ROWS_CONTIG) {
        const int kq = t >> 5, rq = (t & 31) * 4;
        *reinterpret_cast<float4*>(&S[kq][rq]) = v;
    } else {
        const int rr = t >> 1, kq = (t & 1) * 4;
        S[kq][rr] = v.x; S[kq + 1][rr] = v.y; S[kq + 2][rr] = v.z; S[kq + 3][rr] = v.w;
    }
}

// MODE 0 full, 1 lower-triangle, 2 upper-triangle (global indices gi = gi0 + i gis, gj = gj0 + j gjs)
template <bool A_ROWS, bool B_COLS, int MODE>
__global__ void __launch_bounds__(NT, 2) gemm_f32_ffma_kernel(const FArgs p) {
    __shared__ __align__(16) float As[2][TK][PITCH];
    __shared__ __align__(16) float Bs[2][TK][PITCH];
    const int t = threadIdx.x;
    const int ty = t & 15, tx = t >> 4;   // rows fastest
    const i64 total = p.tilesM * p.tilesN;
    for (i64 tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const i64 tm = tile % p.tilesM, tn = tile / p.tilesM;
        const i64 m0 = tm * TM, n0 = tn * TN;
        if (MODE != 0) {
            const i64 mlast = (m0 + TM - 1 < p.m - 1) ? (m0 + TM - 1) : (p.m - 1);
            const i64 nlast = (n0 + TN - 1 < p.n - 1) ? (n0 + TN - 1) : (p.n - 1);
            if (MODE == 1 && !(p.gi0 + mlast * p.gis >= p.gj0 + n0 * p.gjs)) continue;
            if (MODE == 2 && !(p.gi0 + m0 * p.gis <= p.gj0 + nlast * p.gjs)) continue;
        }
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
        const i64 KT = (p.k + TK - 1) / TK;
        float4 ra = load_slab<A_ROWS>(p.A, p.lda, m0, 0, p.m, p.k, t);
        float4 rb = load_slab<B_COLS>(p.B, p.ldb, n0, 0, p.n, p.k, t);
        __syncthreads();   // the previous tile's readers are done with both buffers
        store_slab<A_ROWS>(As[0], ra, t);
        store_slab<B_COLS>(Bs[0], rb, t);
        __syncthreads();
        for (i64 kt = 0; kt < KT; ++kt) {
            const int cur = (int)(kt & 1);
            if (kt + 1 < KT) {
                ra = load_slab<A_ROWS>(p.A, p.lda, m0, (kt + 1) * TK, p.m, p.k, t);
                rb = load_slab<B_COLS>(p.B, p.ldb, n0, (kt + 1) * TK, p.n, p.k, t);
            }
#pragma unroll
            for (int kk = 0; kk < TK; ++kk) {
                const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
                const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
                const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
                const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][64 + tx * 4]);
                const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (kt + 1 < KT) {
                store_slab<A_ROWS>(As[cur ^ 1], ra, t);
                store_slab<B_COLS>(Bs[cur ^ 1], rb, t);
            }
            __syncthreads();
        }
        // ---- epilogue: C = alpha acc + beta C ----
        const bool interior = (m0 + TM <= p.m) && (n0 + TN <= p.n) &&
                              (MODE == 0 ||
                               (MODE == 1 && p.gi0 + m0 * p.gis >= p.gj0 + (n0 + TN - 1) * p.gjs) ||
                               (MODE == 2 && p.gi0 + (m0 + TM - 1) * p.gis <= p.gj0 + n0 * p.gjs));
        const bool vec = interior && ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
        for (int jq = 0; jq < 2; ++jq)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const i64 col = n0 + jq * 64 + tx * 4 + j;
#pragma unroll
                for (int iq = 0; iq < 2; ++iq) {
                    const i64 row = m0 + iq * 64 + ty * 4;
                    float* cp = p.C + row + col * p.ldc;
                    float v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[i] = p.alpha * acc[iq * 4 + i][jq * 4 + j];
                    if (vec) {
                        float4 o = make_float4(v[0], v[1], v[2], v[3]);
                        if (p.beta != 0.f) {
                            const float4 c = *reinterpret_cast<const float4*>(cp);
                            o.x = fmaf(p.beta, c.x, o.x); o.y = fmaf(p.beta, c.y, o.y);
                            o.z = fmaf(p.beta, c.z, o.z); o.w = fmaf(p.beta, c.w, o.w);
                        }
                        *reinterpret_cast<float4*>(cp) = o;
                    } else if (col < p.n) {
                        const i64 gj = p.gj0 + col * p.gjs;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const i64 r = row + i;
                            bool ok = r < p.m;
                            if (MODE == 1) ok = ok && (p.gi0 + r * p.gis >= gj);
                            if (MODE == 2) ok = ok && (p.gi0 + r * p.gis <= gj);
                            if (ok) cp[i] = (p.beta != 0.f) ? fmaf(p.beta, cp[i], v[i]) : v[i];
                        }
                    }
                }
            }
    }
}

template <bool AR, bool BC, int MODE>
void launch(const FArgs& a, cudaStream_t s) {
    const i64 tiles = a.tilesM * a.tilesN;
    i64 grid = tiles;
    const i64 cap = (i64)(sm_limit() > 0 ? sm_limit() : sm_count()) * 2;
    if (grid > cap) grid = cap;
    gemm_f32_ffma_kernel<AR, BC, MODE><<<(unsigned)grid, NT, 0, s>>>(a);
    ELB_LAUNCH_CHECK();
}
template <int MODE>
void dispatch(bool ar, bool bc, const FArgs& a, cudaStream_t s) {
    if (ar) { if (bc) launch<true, true, MODE>(a, s); else launch<true, false, MODE>(a, s); }
    else { if (bc) launch<false, true, MODE>(a, s); else launch<false, false, MODE>(a, s); }
}

}  // namespace

// false (nothing launched) when the operands cannot be read with 16-byte loads
bool sgemm_ffma_device(int mode, bool ta, bool tb, i64 m, i64 n, i64 k, float alpha, const float* A, i64 lda,
                       const float* B, i64 ldb, float beta, float* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                       cudaStream_t s) {
    if (m <= 0 || n <= 0 || k <= 0) return false;
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) || (lda & 3) || (ldb & 3)) return false;
    FArgs a;
    a.m = m; a.n = n; a.k = k;
    a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc;
    a.alpha = alpha; a.beta = beta;
    a.gi0 = gi0; a.gis = gis; a.gj0 = gj0; a.gjs = gjs;
    a.tilesM = ceil_div(m, TM); a.tilesN = ceil_div(n, TN);
    // op(A) rows contiguous <=> A not transposed; op(B) columns contiguous <=> B transposed (stored n x k)
    const bool ar = !ta, bc = tb;
    if (mode == 0) dispatch<0>(ar, bc, a, s);
    else if (mode == 1) dispatch<1>(ar, bc, a, s);
    else dispatch<2>(ar, bc, a, s);
    return true;
}

}  // namespace elb200
