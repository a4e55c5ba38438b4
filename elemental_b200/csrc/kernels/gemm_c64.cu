// Complex<double> GEMM / TRRK / HERK on the DMMA tensor pipe.
//
// Replaces blas::Gemm<Complex<double>> (reference blas.hpp:570-603 -> zgemm_,
// src/core/imports/blas/Gemm.hpp:511), blas::Herk (zherk_, Syrk.hpp:486-519) and the
// LocalTrrk recursion for Complex<double> -- the leaf of the HPDSolve config
// (BASELINE.json configs[3]).
//
// A complex m8n8k4 tile product is four real DMMAs on the interleaved (re,im) operands:
//   Cre += Ar*Br ; Cre += (-Ai)*Bi ; Cim += Ar*Bi ; Cim += Ai*Br
// Each lane fetches its (re,im) fragment entry with ONE LDS.128, so the smem traffic per
// DMMA is half that of the real kernel.  Conjugation ('C') is a sign flip on the loaded
// imaginary part.  CTA tile 128x64x8, 8 warps (4 along M x 2 along N, warp tile 32x32),
// 4-stage cp.async pipeline, accumulators in registers (32 re + 32 im doubles / thread).
// Shared-memory pitches (in 16-byte elements): MN-major BM+2 / BN+2 (== 2 mod 8) and
// K-major 12 (== 4 mod 8): every quarter-warp LDS.128 touches 8 distinct 16-byte banks.
#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int BM = 128, BN = 64, BK = 8;
constexpr int STAGES = 4;
constexpr int NTHREADS = 256;
constexpr int LDK = BK + 4;                 // 12
constexpr int A_TILE = (BM * LDK > BK * (BM + 2)) ? BM * LDK : BK * (BM + 2);  // 1536
constexpr int B_TILE = (BN * LDK > BK * (BN + 2)) ? BN * LDK : BK * (BN + 2);  // 768
constexpr int STAGE_ELEMS = A_TILE + B_TILE;
constexpr int SMEM_BYTES = STAGES * STAGE_ELEMS * 16;  // 147456
constexpr int GROUP_N = 16;

struct ZArgs {
    i64 m, n, k;
    const c64_t* A; i64 lda;
    const c64_t* B; i64 ldb;
    c64_t* C; i64 ldc;
    c64_t alpha, beta;
    i64 gi0, gis, gj0, gjs;
    int conjA, conjB, realDiag;
    i64 tilesM, tilesN;
};

__device__ __forceinline__ void cp16(unsigned dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(bytes));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ROWS x BK tile of op(X): X(r,kk) at X[r + kk*ld] (MN-major) or X[kk + r*ld] (K-major)
template <bool KMAJOR, int ROWS>
__device__ __forceinline__ void load_tile(c64_t* s, const c64_t* __restrict__ X, i64 ld, i64 R, i64 K, i64 r0,
                                          i64 k0, int tid) {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(s);
    constexpr int LDMN = ROWS + 2;
#pragma unroll
    for (int it = 0; it < (ROWS * BK) / NTHREADS; ++it) {
        const int id = tid + it * NTHREADS;
        int rr, kk;
        if (KMAJOR) { kk = id % BK; rr = id / BK; } else { rr = id % ROWS; kk = id / ROWS; }
        const i64 r = r0 + rr, kg = k0 + kk;
        const int valid = (r < R && kg < K) ? 16 : 0;
        const c64_t* src = valid ? (KMAJOR ? X + kg + r * ld : X + r + kg * ld) : X;
        const unsigned off = KMAJOR ? (unsigned)(rr * LDK + kk) : (unsigned)(kk * LDMN + rr);
        cp16(sbase + off * 16u, src, valid);
    }
}
template <bool KMAJOR, int ROWS>
__device__ __forceinline__ double2 frag(const c64_t* s, int r, int kk) {
    constexpr int LDMN = ROWS + 2;
    return *reinterpret_cast<const double2*>(KMAJOR ? s + r * LDK + kk : s + kk * LDMN + r);
}

template <bool A_KMAJOR, bool B_KMAJOR, int MODE>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_c64_kernel(const ZArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    c64_t* smem = reinterpret_cast<c64_t*>(smem_raw);

    const i64 tile = blockIdx.x;
    const i64 group_sz = (i64)GROUP_N * p.tilesM;
    const i64 gid = tile / group_sz;
    const i64 first_n = gid * GROUP_N;
    const i64 gw = (p.tilesN - first_n < GROUP_N) ? (p.tilesN - first_n) : GROUP_N;
    const i64 in_group = tile % group_sz;
    const i64 m0 = (in_group / gw) * BM, n0 = (first_n + in_group % gw) * BN;
    if (MODE != 0) {
        const i64 mlast = (m0 + BM - 1 < p.m - 1) ? (m0 + BM - 1) : (p.m - 1);
        const i64 nlast = (n0 + BN - 1 < p.n - 1) ? (n0 + BN - 1) : (p.n - 1);
        if (MODE == 1 && p.gi0 + mlast * p.gis < p.gj0 + n0 * p.gjs) return;
        if (MODE == 2 && p.gi0 + m0 * p.gis > p.gj0 + nlast * p.gjs) return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp & 3) * 32, wn0 = (warp >> 2) * 32;

    double cre[4][4][2], cim[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) { cre[i][j][0] = cre[i][j][1] = 0.0; cim[i][j][0] = cim[i][j][1] = 0.0; }

    const i64 KT = (p.k + BK - 1) / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            c64_t* sa = smem + s * STAGE_ELEMS;
            load_tile<A_KMAJOR, BM>(sa, p.A, p.lda, p.m, p.k, m0, (i64)s * BK, tid);
            load_tile<B_KMAJOR, BN>(sa + A_TILE, p.B, p.ldb, p.n, p.k, n0, (i64)s * BK, tid);
        }
        cp_commit();
    }
    const double sa_sign = p.conjA ? -1.0 : 1.0, sb_sign = p.conjB ? -1.0 : 1.0;
    for (i64 kt = 0; kt < KT; ++kt) {
        cp_wait<STAGES - 2>();
        __syncthreads();
        {
            const i64 nk = kt + STAGES - 1;
            if (nk < KT) {
                c64_t* sa = smem + (nk % STAGES) * STAGE_ELEMS;
                load_tile<A_KMAJOR, BM>(sa, p.A, p.lda, p.m, p.k, m0, nk * BK, tid);
                load_tile<B_KMAJOR, BN>(sa + A_TILE, p.B, p.ldb, p.n, p.k, n0, nk * BK, tid);
            }
            cp_commit();
        }
        const c64_t* sa = smem + (kt % STAGES) * STAGE_ELEMS;
        const c64_t* sb = sa + A_TILE;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double ar[4], ai[4], nai[4], br[4], bi[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const double2 v = frag<A_KMAJOR, BM>(sa, wm0 + i * 8 + g, ks * 4 + t);
                ar[i] = v.x; ai[i] = sa_sign * v.y; nai[i] = -ai[i];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const double2 v = frag<B_KMAJOR, BN>(sb, wn0 + j * 8 + g, ks * 4 + t);
                br[j] = v.x; bi[j] = sb_sign * v.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    dmma(cre[i][j][0], cre[i][j][1], ar[i], br[j]);
                    dmma(cim[i][j][0], cim[i][j][1], ar[i], bi[j]);
                    dmma(cre[i][j][0], cre[i][j][1], nai[i], bi[j]);
                    dmma(cim[i][j][0], cim[i][j][1], ai[i], br[j]);
                }
        }
    }
    cp_wait<0>();

    const c64_t alpha = p.alpha, beta = p.beta;
    const bool useC = !(beta.re == 0.0 && beta.im == 0.0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const i64 col = n0 + wn0 + j * 8 + 2 * t + e;
            const i64 gj = p.gj0 + col * p.gjs;
            c64_t* cptr = p.C + col * p.ldc;
            c64_t old[4];
            bool ok[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const i64 row = m0 + wm0 + i * 8 + g;
                bool v = (col < p.n) && (row < p.m);
                if (MODE == 1) v = v && (p.gi0 + row * p.gis >= gj);
                if (MODE == 2) v = v && (p.gi0 + row * p.gis <= gj);
                ok[i] = v;
                old[i] = (v && useC) ? cptr[row] : mk(0.0, 0.0);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const i64 row = m0 + wm0 + i * 8 + g;
                if (ok[i]) {
                    c64_t v = alpha * mk(cre[i][j][e], cim[i][j][e]);
                    if (useC) v += beta * old[i];
                    if (p.realDiag && (p.gi0 + row * p.gis == gj)) v.im = 0.0;
                    cptr[row] = v;
                }
            }
        }
    }
}

template <bool AK, bool BK_, int MODE>
void launch(const ZArgs& a, cudaStream_t s) {
    static bool configured = false;
    auto kern = gemm_c64_kernel<AK, BK_, MODE>;
    if (!configured) {
        ELB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    kern<<<(unsigned)(a.tilesM * a.tilesN), NTHREADS, SMEM_BYTES, s>>>(a);
    ELB_LAUNCH_CHECK();
}
template <int MODE>
void dispatch(bool ak, bool bk, const ZArgs& a, cudaStream_t s) {
    if (ak) { if (bk) launch<true, true, MODE>(a, s); else launch<true, false, MODE>(a, s); }
    else { if (bk) launch<false, true, MODE>(a, s); else launch<false, false, MODE>(a, s); }
}
int tcode(char c, const char* what) {
    c = up(c);
    if (c == 'N') return 0;
    if (c == 'T') return 1;
    if (c == 'C') return 2;
    throw std::logic_error(std::string("invalid orientation for ") + what);
}
}  // namespace

bool zgemm_real_device(int mode, int ta, int tb, i64 m, i64 n, i64 k, c64_t alpha, const c64_t* A, i64 lda,
                       const c64_t* B, i64 ldb, c64_t beta, c64_t* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                       bool realDiag, cudaStream_t s);
int g_zgemm_path = 0;   // 0 automatic, 1 always the complex cp.async kernel (elb200_zgemm_set_path)
int g_zgemm_last = 0;   // 1 complex kernel, 2 real persistent kernel

void zgemm_device_ex(int mode, char transA, char transB, i64 m, i64 n, i64 k, c64_t alpha, const c64_t* A, i64 lda,
                     const c64_t* B, i64 ldb, c64_t beta, c64_t* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                     bool realDiag, cudaStream_t s) {
    if (m < 0 || n < 0 || k < 0) throw std::logic_error("zgemm: negative dimension");
    const int ta = tcode(transA, "A"), tb = tcode(transB, "B");
    if (m == 0 || n == 0) return;
    if (lda < ((ta ? k : m) > 1 ? (ta ? k : m) : 1)) throw std::logic_error("zgemm: lda too small");
    if (ldb < ((tb ? n : k) > 1 ? (tb ? n : k) : 1)) throw std::logic_error("zgemm: ldb too small");
    if (ldc < (m > 1 ? m : 1)) throw std::logic_error("zgemm: ldc too small");
    // large products run on the real persistent kernel (gemm_c64_real.cu): same flops, tensor pipe at the real rate
    if (g_zgemm_path != 1 &&
        zgemm_real_device(mode, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs, realDiag, s)) {
        g_zgemm_last = 2;
        return;
    }
    g_zgemm_last = 1;
    ZArgs a;
    a.m = m; a.n = n; a.k = (alpha.re == 0.0 && alpha.im == 0.0) ? 0 : k;
    a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc;
    a.alpha = alpha; a.beta = beta;
    a.gi0 = gi0; a.gis = gis; a.gj0 = gj0; a.gjs = gjs;
    a.conjA = ta == 2; a.conjB = tb == 2; a.realDiag = realDiag ? 1 : 0;
    a.tilesM = ceil_div(m, BM); a.tilesN = ceil_div(n, BN);
    const bool ak = ta != 0, bk = tb == 0;
    if (mode == 0) dispatch<0>(ak, bk, a, s);
    else if (mode == 1) dispatch<1>(ak, bk, a, s);
    else dispatch<2>(ak, bk, a, s);
}

void zgemm_device(int mode, char ta, char tb, i64 m, i64 n, i64 k, c64_t alpha, const c64_t* A, i64 lda,
                  const c64_t* B, i64 ldb, c64_t beta, c64_t* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                  cudaStream_t s) {
    zgemm_device_ex(mode, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs, false, s);
}

}  // namespace elb200

extern "C" {
// 0 automatic (large products on the real persistent kernel), 1 always the dedicated complex kernel
void elb200_zgemm_set_path(int p) { elb200::g_zgemm_path = p; }
// which kernel the last zgemm / ztrrk / zherk / zsyrk call used: 1 complex cp.async kernel, 2 real persistent kernel
int elb200_zgemm_last_kernel(void) { return elb200::g_zgemm_last; }
}
