// FP64 GEMM / TRRK on the DMMA tensor pipe (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4,
// the only f64 MMA shape sm_100a has; tcgen05 has no f64 kind).
//
// Replaces blas::Gemm<double> (reference include/El/core/imports/blas.hpp:570-603 ->
// dgemm_, src/core/imports/blas/Gemm.hpp:431) and, with MODE != 0, the whole
// LocalTrrk recursion (src/blas_like/level3/Trrk/Local.hpp:782-830) as ONE masked
// GEMM: tiles entirely outside the global triangle exit before loading anything,
// tiles that straddle it mask per element in the epilogue.
//
// Layout: column-major operands.  op(A) is m x k, op(B) is k x n.  Each operand
// tile is staged in shared memory in one of two layouts, chosen by which
// dimension is contiguous in global memory:
//   MN-major  s[k][LDMN]   (A 'N', B 'T')   pitch 132 doubles
//   K-major   s[r][LDK]    (A 'T', B 'N')   pitch 20 doubles
// Both pitches are == 4 (mod 16) doubles, which makes every half-warp LDS.64 of
// an m8n8k4 fragment (lane -> row g=lane/4, k t=lane%4) hit 16 distinct bank
// pairs: conflict-free without swizzling.
//
// CTA tile 128x128x16, 8 warps (2 along M x 4 along N, warp tile 64x32),
// 4-stage cp.async pipeline (160 KB smem, one CTA per SM), accumulators in
// registers (64 doubles / thread).
#include "../common.hpp"
#include "elb200_blas.h"

namespace elb200 {
extern int g_dgemm_config;
extern int g_dgemm_last_kernel;
// gemm_f64_ws.cu: persistent TMA-fed kernel; false when the operands are not TMA-eligible
bool dgemm_ws_device(int mode, bool ta, bool tb, i64 m, i64 n, i64 k, double alpha, const double* A, i64 lda,
                     const double* B, i64 ldb, double beta, double* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                     double flops, cudaStream_t s, int rowPair = 0);
namespace {

constexpr int BK = 16;
constexpr int LDK = BK + 4;   // K-major pitch (doubles), == 4 mod 16
constexpr int GROUP_N = 8;    // tile columns per rasterisation group

// Tile configuration.  Every warp owns a 64 x 32 sub-tile (8 x 4 m8n8k4 accumulators,
// 64 doubles per lane); WM x WN warps make the CTA tile.
//   Cfg128x128: 8 warps, 4 stages, 160 KB smem, one CTA per SM  -- best for long k
//   Cfg128x64 : 4 warps, 3 stages,  90 KB smem, TWO CTAs per SM -- the two CTAs are not
//               synchronised with each other, so one CTA's C read-modify-write epilogue and
//               pipeline prologue overlap the other's DMMA main loop.  This is what makes the
//               rank-nb updates (k = Blocksize() = 128..256) of SUMMA-C and Cholesky fast.
template <int WM_, int WN_, int STAGES_, int MINB_>
struct Cfg {
    static constexpr int WM = WM_, WN = WN_, STAGES = STAGES_, MINB = MINB_;
    static constexpr int BM = 64 * WM, BN = 32 * WN;
    static constexpr int NT = 32 * WM * WN;
    static constexpr int LDA_MN = BM + 4, LDB_MN = BN + 4;  // MN-major pitches, == 4 mod 16
    static constexpr int A_TILE = (BM * LDK > BK * LDA_MN) ? BM * LDK : BK * LDA_MN;
    static constexpr int B_TILE = (BN * LDK > BK * LDB_MN) ? BN * LDK : BK * LDB_MN;
    static constexpr int STAGE_DOUBLES = A_TILE + B_TILE;
    static constexpr int SMEM_BYTES = STAGES * STAGE_DOUBLES * 8;
};
typedef Cfg<2, 4, 4, 1> Cfg128x128;
typedef Cfg<2, 2, 3, 2> Cfg128x64;

struct GemmArgs {
    i64 m, n, k;
    const double* A;
    i64 lda;
    const double* B;
    i64 ldb;
    double* C;
    i64 ldc;
    double alpha, beta;
    // global index of local (i,j) is (gi0 + i*gis, gj0 + j*gjs) -- TRRK only
    i64 gi0, gis, gj0, gjs;
    int vecA, vecB;  // 1 when 16-byte cp.async is legal for that operand
    double flops;    // algorithmic flops of this launch (host-side bookkeeping only)
    i64 tilesM, tilesN;
};

__device__ __forceinline__ void cp_async16(unsigned dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async8(unsigned dst, const void* src, int src_bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(dst), "l"(src),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// Stage one ROWS x 16 operand tile.  X(r,kk) is the logical op(X) entry; R, K are
// the logical extents; rows beyond them are zero-filled (cp.async src-size 0).
template <bool KMAJOR, int ROWS, int NT>
__device__ __forceinline__ void load_tile(double* s, const double* __restrict__ X, i64 ld, i64 R,
                                          i64 K, i64 r0, i64 k0, int vec, int tid) {
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(s);
    constexpr int LDMN = ROWS + 4;
    if (!KMAJOR) {
        // X(r,kk) at X[r + kk*ld]; smem s[kk*LDMN + r]
        if (vec) {
#pragma unroll
            for (int it = 0; it < (BK * ROWS / 2) / NT; ++it) {
                int id = tid + it * NT;
                int kk = id / (ROWS / 2);
                int rr = (id % (ROWS / 2)) * 2;
                i64 r = r0 + rr, kg = k0 + kk;
                int valid = 0;
                if (kg < K) {
                    i64 rem = R - r;
                    valid = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
                }
                const double* src = valid ? (X + r + kg * ld) : X;
                cp_async16(sbase + (unsigned)(kk * LDMN + rr) * 8u, src, valid);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (BK * ROWS) / NT; ++it) {
                int id = tid + it * NT;
                int kk = id / ROWS;
                int rr = id % ROWS;
                i64 r = r0 + rr, kg = k0 + kk;
                int valid = (kg < K && r < R) ? 8 : 0;
                const double* src = valid ? (X + r + kg * ld) : X;
                cp_async8(sbase + (unsigned)(kk * LDMN + rr) * 8u, src, valid);
            }
        }
    } else {
        // X(r,kk) at X[kk + r*ld]; smem s[r*LDK + kk]
        if (vec) {
#pragma unroll
            for (int it = 0; it < (ROWS * BK / 2) / NT; ++it) {
                int id = tid + it * NT;
                int rr = id / (BK / 2);
                int kk = (id % (BK / 2)) * 2;
                i64 r = r0 + rr, kg = k0 + kk;
                int valid = 0;
                if (r < R) {
                    i64 rem = K - kg;
                    valid = rem >= 2 ? 16 : (rem == 1 ? 8 : 0);
                }
                const double* src = valid ? (X + kg + r * ld) : X;
                cp_async16(sbase + (unsigned)(rr * LDK + kk) * 8u, src, valid);
            }
        } else {
#pragma unroll
            for (int it = 0; it < (ROWS * BK) / NT; ++it) {
                int id = tid + it * NT;
                int rr = id / BK;
                int kk = id % BK;
                i64 r = r0 + rr, kg = k0 + kk;
                int valid = (kg < K && r < R) ? 8 : 0;
                const double* src = valid ? (X + kg + r * ld) : X;
                cp_async8(sbase + (unsigned)(rr * LDK + kk) * 8u, src, valid);
            }
        }
    }
}

template <bool KMAJOR, int ROWS>
__device__ __forceinline__ double frag(const double* s, int r, int kk) {
    return KMAJOR ? s[r * LDK + kk] : s[kk * (ROWS + 4) + r];
}

// MODE 0: full GEMM; 1: lower-triangle TRRK; 2: upper-triangle TRRK
template <class CF, bool A_KMAJOR, bool B_KMAJOR, int MODE>
__global__ void __launch_bounds__(CF::NT, CF::MINB) gemm_f64_kernel(const GemmArgs p) {
    extern __shared__ __align__(128) double smem[];
    constexpr int BM = CF::BM, BN = CF::BN, STAGES = CF::STAGES, NT = CF::NT;

    // ---- tile coordinates (grouped rasterisation for L2 reuse) ----
    const i64 tile = blockIdx.x;
    const i64 group_sz = (i64)GROUP_N * p.tilesM;
    const i64 gid = tile / group_sz;
    const i64 first_n = gid * GROUP_N;
    const i64 gw = (p.tilesN - first_n < GROUP_N) ? (p.tilesN - first_n) : GROUP_N;
    const i64 in_group = tile % group_sz;
    const i64 tm = in_group / gw;
    const i64 tn = first_n + in_group % gw;
    const i64 m0 = tm * BM, n0 = tn * BN;

    if (MODE != 0) {
        const i64 mlast = (m0 + BM - 1 < p.m - 1) ? (m0 + BM - 1) : (p.m - 1);
        const i64 nlast = (n0 + BN - 1 < p.n - 1) ? (n0 + BN - 1) : (p.n - 1);
        if (MODE == 1) {  // lower: need some gi >= gj
            if (p.gi0 + mlast * p.gis < p.gj0 + n0 * p.gjs) return;
        } else {  // upper: need some gi <= gj
            if (p.gi0 + m0 * p.gis > p.gj0 + nlast * p.gjs) return;
        }
    }

    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (warp % CF::WM) * 64, wn0 = (warp / CF::WM) * 32;

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    const i64 KT = (p.k + BK - 1) / BK;

    // ---- prologue: STAGES-1 tiles in flight ----
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < KT) {
            double* sa = smem + s * CF::STAGE_DOUBLES;
            load_tile<A_KMAJOR, BM, NT>(sa, p.A, p.lda, p.m, p.k, m0, (i64)s * BK, p.vecA, tid);
            load_tile<B_KMAJOR, BN, NT>(sa + CF::A_TILE, p.B, p.ldb, p.n, p.k, n0, (i64)s * BK, p.vecB, tid);
        }
        cp_async_commit();
    }

    for (i64 kt = 0; kt < KT; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const i64 nk = kt + STAGES - 1;
            if (nk < KT) {
                double* sa = smem + (nk % STAGES) * CF::STAGE_DOUBLES;
                load_tile<A_KMAJOR, BM, NT>(sa, p.A, p.lda, p.m, p.k, m0, nk * BK, p.vecA, tid);
                load_tile<B_KMAJOR, BN, NT>(sa + CF::A_TILE, p.B, p.ldb, p.n, p.k, n0, nk * BK, p.vecB, tid);
            }
            cp_async_commit();
        }
        const double* sa = smem + (kt % STAGES) * CF::STAGE_DOUBLES;
        const double* sb = sa + CF::A_TILE;
        // register double-buffered fragments: the LDS of k-step ks+1 are in flight while the
        // 32 DMMAs of k-step ks issue
        double a[2][8], b[2][4];
#pragma unroll
        for (int i = 0; i < 8; ++i) a[0][i] = frag<A_KMAJOR, BM>(sa, wm0 + i * 8 + g, t);
#pragma unroll
        for (int j = 0; j < 4; ++j) b[0][j] = frag<B_KMAJOR, BN>(sb, wn0 + j * 8 + g, t);
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            const int cur = ks & 1, nxt = cur ^ 1;
            if (ks + 1 < BK / 4) {
#pragma unroll
                for (int i = 0; i < 8; ++i) a[nxt][i] = frag<A_KMAJOR, BM>(sa, wm0 + i * 8 + g, (ks + 1) * 4 + t);
#pragma unroll
                for (int j = 0; j < 4; ++j) b[nxt][j] = frag<B_KMAJOR, BN>(sb, wn0 + j * 8 + g, (ks + 1) * 4 + t);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[cur][i], b[cur][j]);
        }
    }
    cp_async_wait<0>();

    // ---- epilogue: C = alpha*acc + beta*C (masked for TRRK) ----
    // All loads of one 8-column group are issued before the first store so that the
    // read-modify-write of C costs one memory round trip per group, not one per element.
    const double alpha = p.alpha, beta = p.beta;
    const bool useC = (beta != 0.0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        double old[2][8];
        bool ok[2][8];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const i64 col = n0 + wn0 + j * 8 + 2 * t + e;
            const i64 gj = p.gj0 + col * p.gjs;
            const double* cptr = p.C + col * p.ldc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const i64 row = m0 + wm0 + i * 8 + g;
                bool v = (col < p.n) && (row < p.m);
                if (MODE == 1) v = v && (p.gi0 + row * p.gis >= gj);
                if (MODE == 2) v = v && (p.gi0 + row * p.gis <= gj);
                ok[e][i] = v;
                old[e][i] = (v && useC) ? cptr[row] : 0.0;
            }
        }
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const i64 col = n0 + wn0 + j * 8 + 2 * t + e;
            double* cptr = p.C + col * p.ldc;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const i64 row = m0 + wm0 + i * 8 + g;
                if (ok[e][i]) {
                    double v = __dmul_rn(alpha, acc[i][j][e]);
                    if (useC) v = __fma_rn(beta, old[e][i], v);
                    cptr[row] = v;
                }
            }
        }
    }
}

template <class CF, bool AK, bool BK_, int MODE>
void launch(GemmArgs a, cudaStream_t s) {
    static bool configured = false;
    auto kern = gemm_f64_kernel<CF, AK, BK_, MODE>;
    if (!configured) {
        ELB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM_BYTES));
        ELB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured = true;
    }
    a.tilesM = ceil_div(a.m, CF::BM);
    a.tilesN = ceil_div(a.n, CF::BN);
    const i64 tiles = a.tilesM * a.tilesN;
    gemm_profile_begin(s);
    kern<<<(unsigned)tiles, CF::NT, CF::SMEM_BYTES, s>>>(a);
    ELB_LAUNCH_CHECK();
    gemm_profile_end(s, a.flops);
}

template <class CF, int MODE>
void dispatch_cfg(bool ak, bool bk, const GemmArgs& a, cudaStream_t s) {
    if (ak) {
        if (bk) launch<CF, true, true, MODE>(a, s);
        else launch<CF, true, false, MODE>(a, s);
    } else {
        if (bk) launch<CF, false, true, MODE>(a, s);
        else launch<CF, false, false, MODE>(a, s);
    }
}


template <int MODE>
void dispatch(bool ak, bool bk, const GemmArgs& a, cudaStream_t s) {
    int cfg = g_dgemm_config;
    if (cfg == 0 || cfg == 3) cfg = 2;  // 2 CTAs/SM measured faster than 128x128 at every k (profiles/r01_probe2_v0.txt)
    if (cfg == 1) dispatch_cfg<Cfg128x128, MODE>(ak, bk, a, s);
    else dispatch_cfg<Cfg128x64, MODE>(ak, bk, a, s);
}

bool is_trans(char c, const char* what) {
    c = up(c);
    if (c == 'N') return false;
    if (c == 'T' || c == 'C') return true;
    throw std::logic_error(std::string("invalid orientation for ") + what);
}

}  // namespace

// kernel selection: 0 = automatic (TMA kernel when eligible), 1 = cp.async 128x128 (1 CTA/SM),
// 2 = cp.async 128x64 (2 CTAs/SM), 3 = TMA kernel when eligible else automatic cp.async
int g_dgemm_config = 0;
int g_dgemm_last_kernel = 0;  // 1 = cp.async kernel, 2 = persistent TMA kernel (tests assert on it)

// mode 0 = gemm, 1 = lower trrk, 2 = upper trrk
void dgemm_device(int mode, char transA, char transB, i64 m, i64 n, i64 k, double alpha,
                  const double* A, i64 lda, const double* B, i64 ldb, double beta, double* C,
                  i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs, cudaStream_t s) {
    if (m < 0 || n < 0 || k < 0) throw std::logic_error("dgemm: negative dimension");
    const bool ta = is_trans(transA, "A"), tb = is_trans(transB, "B");
    if (m == 0 || n == 0) return;
    if (lda < ((ta ? k : m) > 1 ? (ta ? k : m) : 1)) throw std::logic_error("dgemm: lda too small");
    if (ldb < ((tb ? n : k) > 1 ? (tb ? n : k) : 1)) throw std::logic_error("dgemm: ldb too small");
    if (ldc < (m > 1 ? m : 1)) throw std::logic_error("dgemm: ldc too small");
    GemmArgs a;
    a.m = m; a.n = n; a.k = (alpha == 0.0) ? 0 : k;
    a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc;
    a.alpha = alpha; a.beta = beta;
    a.gi0 = gi0; a.gis = gis; a.gj0 = gj0; a.gjs = gjs;
    a.vecA = (((uintptr_t)A & 15) == 0 && (lda % 2) == 0) ? 1 : 0;
    a.vecB = (((uintptr_t)B & 15) == 0 && (ldb % 2) == 0) ? 1 : 0;
    a.tilesM = a.tilesN = 0;  // set per tile configuration at launch
    // algorithmic flops: 2mnk for GEMM; for TRRK 2k per C entry inside the global triangle
    if (mode == 0) a.flops = 2.0 * double(m) * double(n) * double(a.k);
    else {
        double inside = 0;
        for (i64 j = 0; j < n; ++j) {
            const i64 gj = gj0 + j * gjs;
            i64 cnt;
            if (mode == 1) {  // rows with gi0 + i*gis >= gj
                i64 first = gj <= gi0 ? 0 : (gj - gi0 + gis - 1) / gis;
                cnt = first >= m ? 0 : m - first;
            } else {  // rows with gi0 + i*gis <= gj
                cnt = gj < gi0 ? 0 : (gj - gi0) / gis + 1;
                if (cnt > m) cnt = m;
            }
            inside += double(cnt);
        }
        a.flops = 2.0 * inside * double(a.k);
    }
    // default: the persistent TMA kernel (gemm_f64_ws.cu); the cp.async kernel below serves
    // operands TMA cannot address (odd leading dimension / 8-byte-aligned base) and cfg 1 / 2
    if (g_dgemm_config == 0 || g_dgemm_config == 3) {
        if (dgemm_ws_device(mode, ta, tb, m, n, a.k, alpha, A, lda, B, ldb, beta, C, ldc, gi0, gis, gj0, gjs,
                            a.flops, s)) {
            g_dgemm_last_kernel = 2;
            return;
        }
    }
    g_dgemm_last_kernel = 1;
    // A 'T' is K-major; B 'N' is K-major
    const bool ak = ta, bk = !tb;
    if (mode == 0) dispatch<0>(ak, bk, a, s);
    else if (mode == 1) dispatch<1>(ak, bk, a, s);
    else dispatch<2>(ak, bk, a, s);
}

}  // namespace elb200

extern "C" {

void elb200_dgemm_set_config(int cfg) { elb200::g_dgemm_config = cfg; }
int elb200_dgemm_last_kernel(void) { return elb200::g_dgemm_last_kernel; }

int elb200_dgemm(char transA, char transB, int64_t m, int64_t n, int64_t k, double alpha,
                 const double* A, int64_t lda, const double* B, int64_t ldb, double beta,
                 double* C, int64_t ldc, elb200_stream_t s) {
    return elb200::guarded([&] {
        elb200::dgemm_device(0, transA, transB, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, 1,
                             0, 1, (cudaStream_t)s);
    });
}

int elb200_dtrrk(char uplo, char transA, char transB, int64_t m, int64_t n, int64_t k,
                 double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
                 double beta, double* C, int64_t ldc, int64_t rowShift, int64_t rowStride,
                 int64_t colShift, int64_t colStride, elb200_stream_t s) {
    return elb200::guarded([&] {
        const char u = elb200::up(uplo);
        if (u != 'L' && u != 'U') throw std::logic_error("dtrrk: uplo must be 'L' or 'U'");
        elb200::dgemm_device(u == 'L' ? 1 : 2, transA, transB, m, n, k, alpha, A, lda, B, ldb,
                             beta, C, ldc, rowShift, rowStride, colShift, colStride,
                             (cudaStream_t)s);
    });
}

int elb200_dsyrk(char uplo, char trans, int64_t n, int64_t k, double alpha, const double* A,
                 int64_t lda, double beta, double* C, int64_t ldc, elb200_stream_t s) {
    return elb200::guarded([&] {
        const char u = elb200::up(uplo);
        if (u != 'L' && u != 'U') throw std::logic_error("dsyrk: uplo must be 'L' or 'U'");
        const bool tr = elb200::up(trans) != 'N';
        // 'N': C = A A^T ; 'T': C = A^T A
        elb200::dgemm_device(u == 'L' ? 1 : 2, tr ? 'T' : 'N', tr ? 'N' : 'T', n, n, k, alpha, A,
                             lda, A, lda, beta, C, ldc, 0, 1, 0, 1, (cudaStream_t)s);
    });
}

}  // extern "C"
