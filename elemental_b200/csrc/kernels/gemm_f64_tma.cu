// FP64 GEMM / TRRK, second generation: persistent, warp-specialised, TMA-fed DMMA kernel.
//
// Same contract as gemm_f64.cu (replaces blas::Gemm<double> -> dgemm_, reference
// src/core/imports/blas/Gemm.hpp:431, and with MODE != 0 the LocalTrrk recursion of
// src/blas_like/level3/Trrk/Local.hpp:782-830), built for the shape that dominates the
// hot path: the rank-Blocksize() update C += A1 * B1 (k = 128..256) of SUMMA-C and of the
// Cholesky trailing update, where the C read-modify-write and the pipeline fill of every
// tile cost as much as the tile's DMMA work unless they are hidden.
//
// One persistent CTA per SM, 8 warps:
//   * two independent groups of 4 warps; each group owns a stream of 128 x 64 C tiles (warp
//     tile 64 x 32, accumulators in registers) and its own 4-stage operand ring, so while one
//     group runs its epilogue the other keeps the FP64 tensor pipe busy (the DMMA pipe issues
//     one m8n8k4 per 16 clocks per SM sub-partition; two resident warps per sub-partition hide
//     each other's fixed issue latencies);
//   * warp 0 of each group also carries the group's producer cursor: before consuming k-stage i
//     its lane 0 issues the TMA loads (cp.async.bulk.tensor.2d, 128-byte swizzle, completion on
//     an mbarrier) of k-stage i+2 -- across tile boundaries, so the pipeline never drains -- and
//     when the cursor enters a new tile all its lanes prefetch that C tile into L2, which makes
//     the epilogue's read-modify-write loads L2 hits.
//
// Shared-memory operand layouts (what TMA writes) and the conflict-free fragment reads:
//   K-major operand (A 'T' / B 'N'; k contiguous in global memory): one box [rows][16 k],
//     row pitch 128 B, 16-byte chunk index XORed with (row & 7).  An m8n8k4 fragment takes its
//     8 rows in the order {0,2,4,6,1,3,5,7}: each half-warp (4 rows x 4 k) then touches 8
//     distinct chunks x 2 halves = all 32 banks exactly once.
//   MN-major operand (A 'N' / B 'T'; rows contiguous): boxes [16 k][16 rows], pitch 128 B,
//     chunk XORed with (k & 7).  A fragment takes rows {0,1,8,9,2,3,10,11} (+4 for the second
//     fragment of the box): again 16 distinct bank pairs per half-warp.
//   The row permutations are undone in the epilogue (they only relabel rows/columns of C).
// TMA zero-fills out-of-range rows / k, so ragged sizes need no special casing in the main
// loop; the epilogue masks (and applies the global staircase mask for TRRK).
//
// Requirements: A, B 16-byte aligned with even leading dimension (TMA strides are multiples
// of 16 B).  Anything else is served by the cp.async kernel of gemm_f64.cu.
#include <cuda.h>

#include <mutex>

#include "../common.hpp"
#include "elb200_blas.h"

namespace elb200 {
int g_dgemm_tma_flags = 0;
unsigned long long* g_dgemm_tma_prof = nullptr;
namespace {

constexpr int BK = 16;             // doubles per k-stage = one 128-byte swizzle span
constexpr int TM = 128, TN = 64;   // C tile of one consumer group
constexpr int STAGES = 4;
constexpr int GROUPS = 2;
// Warp layout of one group.  A lone warp per SM sub-partition sustains only ~62 % of the DMMA issue
// rate (measured: flags bit 2), two or more saturate it.  Cfg8 = 8 warps per group (warp tile
// 32 x 32): even while the other group is in its epilogue the pipe still sees two warps per
// sub-partition.  Cfg4 = 4 warps per group (warp tile 64 x 32): fewer shared-memory reads per DMMA.
template <int WM_, int WN_>
struct GroupCfg {
    static constexpr int WM = WM_, WN = WN_;
    static constexpr int WARPS = WM * WN;           // warps per group
    static constexpr int FM = 128 / (8 * WM);       // 8-row A fragments per warp
    static constexpr int FN = 64 / (8 * WN);        // 8-column B fragments per warp
    static constexpr int NTHREADS = 32 * 2 * WARPS;
};
typedef GroupCfg<2, 2> Cfg4;
typedef GroupCfg<4, 2> Cfg8;
// 8 warps = 2 per SM sub-partition, so every thread may use up to 255 registers (128 of them hold
// the 64 x 32 warp tile).  A dedicated producer warpgroup with setmaxnreg (168 -> 40 / 232) was
// measured too: upper registers obtained that way were intermittently corrupted while global loads
// into them were in flight (wrong C read-modify-write), so the TMA issue lives in
// warp 0 of each group instead.
constexpr int LOOKAHEAD = 3;  // k-stages the producer cursor runs ahead of the consumers (< STAGES)
constexpr int A_BYTES = TM * BK * 8;  // 16 KB
constexpr int B_BYTES = TN * BK * 8;  //  8 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int RING_BYTES = STAGES * STAGE_BYTES;
constexpr int SMEM_BYTES = GROUPS * RING_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;
constexpr int GO_BAR_OFFSET = GROUPS * 2 * STAGES * 8;  // the stagger barrier follows full/empty[G][S]
constexpr int GROUP_N = 16;  // tile columns per rasterisation band (in 64-column tiles)

struct TmaArgs {
    CUtensorMap mapA, mapB;
    i64 m, n, k;
    double* C;
    i64 ldc;
    double alpha, beta;
    i64 gi0, gis, gj0, gjs;
    i64 tilesM, tilesN;
    unsigned long long* prof;  // MODE 6 only: per-warp clock counters, 8 per warp
    int flags;  // tuning/debug: bit0 = stagger the two groups, bit1 = always use the masked epilogue,
                // bit2 = only group 0 works, bit3 = 4 warps per group (warp tile 64 x 32) instead of 8,
                // bit4 = never use the L2 reduction epilogue, bit7 / bit8 = diagnostic kernels (MODE 3 / 4, NN only), bit9 = fragment-prefetch kernel (MODE 5, NN only), bit5 / bit6 = one-off start offsets per SM / per group
};

// ---- PTX wrappers ------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
// C[i] += v performed by the L2 atomic unit: no load, no register for the old value, nothing to wait
// for.  Each C entry is touched by exactly one thread per launch, so the result is deterministic and
// equal to the load-add-store form.  Measured 2.4 TB/s of C bytes with the epilogue's access pattern
// (scripts/micro/bulk_red_f64.cu), i.e. HBM-bound like a plain read-modify-write stream.
__device__ __forceinline__ void red_add_f64(double* addr, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- row permutations ---------------------------------------------------------------------
// K-major operand: fragment f (8 rows) of a warp takes rows base + 8 f + PK[x], x = MMA row index
__device__ __forceinline__ int permK(int x) { return ((x & 3) << 1) | (x >> 2); }
// MN-major operand: fragment f takes rows base + 16 (f >> 1) + 4 (f & 1) + QM[x]
__device__ __forceinline__ int permM(int x) { return ((x & 2) << 2) | (x & 1) | ((x & 4) >> 1); }  // {0,1,8,9,2,3,10,11}

template <bool KMAJOR>
__device__ __forceinline__ int tile_row(int f, int x) {
    return KMAJOR ? (8 * f + permK(x)) : (16 * (f >> 1) + 4 * (f & 1) + permM(x));
}

// tile index -> (tm, tn): bands of GROUP_N tile columns, walked down the rows (L2 reuse of the
// B band and of the A row panels across the CTAs that run concurrently).  32-bit arithmetic: the
// host guarantees tilesM * tilesN < 2^31.
__device__ __forceinline__ void tile_coords(const TmaArgs& p, i64 tile64, i64& tm, i64& tn) {
    const unsigned tile = (unsigned)tile64;
    const unsigned tilesM = (unsigned)p.tilesM, tilesN = (unsigned)p.tilesN;
    const unsigned band_sz = (unsigned)GROUP_N * tilesM;
    const unsigned band = tile / band_sz;
    const unsigned first_n = band * GROUP_N;
    const unsigned bw = (tilesN - first_n < (unsigned)GROUP_N) ? (tilesN - first_n) : (unsigned)GROUP_N;
    const unsigned in_band = tile - band * band_sz;
    const unsigned q = in_band / bw;
    tm = q;
    tn = first_n + (in_band - q * bw);
}

template <int MODE>
__device__ __forceinline__ bool tile_active(const TmaArgs& p, i64 tm, i64 tn) {
    if (MODE == 0 || MODE >= 3) return true;
    const i64 m0 = tm * TM, n0 = tn * TN;
    const i64 mlast = (m0 + TM - 1 < p.m - 1) ? (m0 + TM - 1) : (p.m - 1);
    const i64 nlast = (n0 + TN - 1 < p.n - 1) ? (n0 + TN - 1) : (p.n - 1);
    if (MODE == 1) return p.gi0 + mlast * p.gis >= p.gj0 + n0 * p.gjs;  // some gi >= gj
    return p.gi0 + m0 * p.gis <= p.gj0 + nlast * p.gjs;                // some gi <= gj
}

// One k-stage of loads for the tile at (m0, n0): expect_tx + TMA boxes, issued by one lane
template <bool A_KMAJOR, bool B_KMAJOR>
__device__ __forceinline__ void issue_stage(const TmaArgs& p, unsigned sa, unsigned full, i64 m0, i64 n0, i64 kt) {
    const unsigned sb = sa + A_BYTES;
    mbar_expect_tx(full, STAGE_BYTES);
    const int k0 = (int)(kt * BK);
    if (A_KMAJOR) {
        tma_load_2d(sa, &p.mapA, k0, (int)m0, full);
    } else {
#pragma unroll
        for (int b = 0; b < TM / 16; ++b) tma_load_2d(sa + b * 2048, &p.mapA, (int)m0 + 16 * b, k0, full);
    }
    if (B_KMAJOR) {
        tma_load_2d(sb, &p.mapB, k0, (int)n0, full);
    } else {
#pragma unroll
        for (int b = 0; b < TN / 16; ++b) tma_load_2d(sb + b * 2048, &p.mapB, (int)n0 + 16 * b, k0, full);
    }
}

// MODE 0: full GEMM; 1: lower-triangle TRRK; 2: upper-triangle TRRK;
// 3 / 4: DIAGNOSTIC full GEMM whose interior tiles skip the C update / use plain stores C = alpha acc (wrong
// results on purpose: they split the rank-nb deficit between the pipeline and the epilogue; separate
// instantiations, so the code generated for modes 0-2 is untouched); 5: EXPERIMENT, full GEMM whose
// fragment loads run one k-step ahead of the DMMAs (bit-identical results)
template <class CF, bool A_KMAJOR, bool B_KMAJOR, int MODE>
__global__ void __launch_bounds__(CF::NTHREADS, 1) gemm_f64_tma_kernel(const __grid_constant__ TmaArgs p) {
    constexpr int CONSUMER_WARPS = CF::WARPS, FM = CF::FM, FN = CF::FN;
    extern __shared__ unsigned char smem_raw[];
    const unsigned raw = smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;       // 1024-byte alignment for the 128 B swizzle
    const unsigned bars = base + GROUPS * RING_BYTES;   // full[G][S], empty[G][S], go

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int group = warp / CONSUMER_WARPS;
    const int cw = warp % CONSUMER_WARPS;
    const unsigned ring = base + group * RING_BYTES;
    const unsigned full0 = bars + (group * 2 * STAGES) * 8;
    const unsigned empty0 = full0 + STAGES * 8;

    if (tid == 0) {
        for (int g2 = 0; g2 < GROUPS; ++g2)
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(bars + ((g2 * 2 * STAGES) + s) * 8, 1);
                mbar_init(bars + ((g2 * 2 * STAGES) + STAGES + s) * 8, CONSUMER_WARPS);
            }
        mbar_init(bars + GO_BAR_OFFSET, CONSUMER_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();

    const i64 KT = (p.k + BK - 1) / BK;
    const i64 total_tiles = p.tilesM * p.tilesN;
    const bool single = (p.flags & 4) != 0;  // experiment: only group 0 works (lone-warp DMMA rate)
    if (single && group == 1) return;
    // experiment (flags bit 5 / bit 6): every SM runs equally long tiles, so the read-modify-write bursts
    // of all epilogues coincide; a one-off start offset per SM (bit 5: blockIdx % 4 quarter tiles) and per
    // group (bit 6: group 1 half a tile later) spreads them for the whole launch.  p.k sets the tile time
    // (~130 ns of DMMA per k-column of a tile pair).
    if (p.flags & 96) {
        unsigned ns = 0;
        const unsigned tile_ns = (unsigned)(p.k > 2048 ? 2048 : p.k) * 130u;
        if (p.flags & 32) ns += (blockIdx.x & 3u) * (tile_ns / 4u);
        if ((p.flags & 64) && group == 1) ns += tile_ns / 2u;
        for (unsigned waited = 0; waited < ns; waited += 1000u) __nanosleep(1000u);
    }
    const i64 first_tile = single ? (i64)blockIdx.x : (i64)blockIdx.x * GROUPS + group;
    const i64 tile_step = single ? (i64)gridDim.x : (i64)gridDim.x * GROUPS;
    const bool useC = (p.beta != 0.0);
    const bool useRed = (p.beta == 1.0) && !(p.flags & 16);  // rank-k update form: C += alpha op(A) op(B)

    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (cw % CF::WM) * (8 * FM), wn0 = (cw / CF::WM) * (8 * FN);

    // per-lane fragment address pieces (bytes, relative to the operand's stage base)
    //   K-major : row*128 + (((2ks + t/2) ^ (row&7)) << 4) + (t&1)*8,  row = wbase + 8f + PK[g]
    //   MN-major: box*2048 + (4ks+t)*128 + (((rho/2) ^ ((4ks+t)&7)) << 4) + (rho&1)*8,
    //             rho = 4(f&1) + QM[g], box = wbase/16 + f/2
    unsigned aoff[4], boff[4];
    if (A_KMAJOR) {
        const int pr = permK(g);
        const unsigned rowoff = (unsigned)(wm0 + pr) * 128u + (unsigned)(t & 1) * 8u;
        const int L = pr ^ (t >> 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) aoff[ks] = rowoff + (unsigned)((L ^ (2 * ks)) << 4);
    } else {
        const int q = permM(g);
        const unsigned lo = (unsigned)(wm0 / 16) * 2048u + (unsigned)t * 128u + (unsigned)(q & 1) * 8u;
        const int L = (q >> 1) ^ t;
#pragma unroll
        for (int x = 0; x < 4; ++x) aoff[x] = lo + (unsigned)((L ^ (2 * x)) << 4);  // x = (f&1) + 2*(ks&1)
    }
    if (B_KMAJOR) {
        const int pr = permK(g);
        const unsigned rowoff = (unsigned)(wn0 + pr) * 128u + (unsigned)(t & 1) * 8u;
        const int L = pr ^ (t >> 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) boff[ks] = rowoff + (unsigned)((L ^ (2 * ks)) << 4);
    } else {
        const int q = permM(g);
        const unsigned lo = (unsigned)(wn0 / 16) * 2048u + (unsigned)t * 128u + (unsigned)(q & 1) * 8u;
        const int L = (q >> 1) ^ t;
#pragma unroll
        for (int x = 0; x < 4; ++x) boff[x] = lo + (unsigned)((L ^ (2 * x)) << 4);
    }

    // ---- producer cursor (warp 0 of the group): runs LOOKAHEAD k-stages ahead of the consumers,
    //      across tile boundaries ----
    i64 ptile = first_tile, pkt = 0, pm0 = 0, pn0 = 0;
    int pstage = 0;
    unsigned pphase = 0;
    bool pvalid = false;
    // move the producer cursor to the next active tile at or after `from`; prefetch its C tile into L2
    auto producer_seek = [&](i64 from) {
        pvalid = false;
        for (i64 tile = from; tile < total_tiles; tile += tile_step) {
            i64 tm, tn;
            tile_coords(p, tile, tm, tn);
            if (!tile_active<MODE>(p, tm, tn)) continue;
            ptile = tile; pm0 = tm * TM; pn0 = tn * TN; pkt = 0; pvalid = true;
            break;
        }
        if (pvalid && useC && !useRed) {
            // 64 columns x 128 rows x 8 B: lane handles columns lane and lane + 32
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const i64 col = pn0 + lane + 32 * cc;
                if (col < p.n) {
                    const char* cp = (const char*)(p.C + pm0 + col * p.ldc);
                    const i64 rows = (p.m - pm0 < TM) ? (p.m - pm0) : TM;
                    const char* end = cp + rows * 8;
                    for (const char* q = (const char*)((uintptr_t)cp & ~(uintptr_t)127); q < end; q += 128) prefetch_l2(q);
                }
            }
        }
    };
    // issue the loads of one k-stage at the producer cursor and advance it
    auto producer_step = [&]() {
        if (!pvalid) return;
        mbar_wait(empty0 + pstage * 8, pphase ^ 1u);
        if (lane == 0)
            issue_stage<A_KMAJOR, B_KMAJOR>(p, ring + pstage * STAGE_BYTES, full0 + pstage * 8, pm0, pn0, pkt);
        __syncwarp();
        if (++pstage == STAGES) { pstage = 0; pphase ^= 1u; }
        if (++pkt == KT) producer_seek(ptile + tile_step);
    };
    if (cw == 0) {
        producer_seek(first_tile);
        const int lookahead = (p.flags & 2048) ? LOOKAHEAD - 1 : LOOKAHEAD;  // bit 11: producer 2 stages ahead
#pragma unroll 1
        for (int i = 0; i < lookahead; ++i) producer_step();
    }

    int stage = 0;
    unsigned phase = 0;
    const double alpha = p.alpha, beta = p.beta;
    // MODE 6 (diagnostic instantiation): SM clocks this warp spent in {0 producer_step, 1 waiting for a full
    // stage, 2 fragment loads + DMMAs, 3 epilogue, 4 whole kernel, 5 tiles}
    long long pc[6] = {0, 0, 0, 0, 0, 0};
    long long pt0 = 0, ptk = 0;
    if constexpr (MODE == 6) pt0 = clock64();
    // Optional stagger (flags bit 0): group 1 starts its first tile when group 0 is half-way
    // through its own, so that the epilogue of one group runs under the main loop of the other.
    const unsigned go_bar = bars + GO_BAR_OFFSET;
    const bool stagger = (p.flags & 1) != 0;
    bool released = false;  // group 0: has signalled; group 1: has waited

    for (i64 tile = first_tile; tile < total_tiles; tile += tile_step) {
        i64 tm, tn;
        tile_coords(p, tile, tm, tn);
        if (!tile_active<MODE>(p, tm, tn)) continue;
        const i64 m0 = tm * TM, n0 = tn * TN;

        double acc[FM][FN][2];
#pragma unroll
        for (int i = 0; i < FM; ++i)
#pragma unroll
            for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

        if (stagger && group == 1 && !released) {
            mbar_wait(go_bar, 0);
            released = true;
        }
        if constexpr (MODE == 5) {
            // EXPERIMENT (separate instantiation, elb200_dgemm_set_debug_flags(512), NN only): the fragments of
            // k-step ks + 1 are loaded before the DMMAs of k-step ks are issued -- across k-stage boundaries
            // too -- so a fragment load may sit behind a burst of the other group's epilogue REDs in the LSU
            // queue for a whole k-step (~1000 clocks) without starving the tensor pipe.  Same DMMA order as the
            // default loop, hence bit-identical results.
            double fa[2][FM], fb[2][FN];
            auto load_frags = [&](double (&a)[FM], double (&b)[FN], unsigned sa, unsigned sb, int ks) {
#pragma unroll
                for (int i = 0; i < FM; ++i) {
                    if (A_KMAJOR) a[i] = lds64(sa + aoff[ks] + (unsigned)i * 1024u);
                    else a[i] = lds64(sa + aoff[(i & 1) + 2 * (ks & 1)] + (unsigned)(i >> 1) * 2048u + (unsigned)ks * 512u);
                }
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    if (B_KMAJOR) b[j] = lds64(sb + boff[ks] + (unsigned)j * 1024u);
                    else b[j] = lds64(sb + boff[(j & 1) + 2 * (ks & 1)] + (unsigned)(j >> 1) * 2048u + (unsigned)ks * 512u);
                }
            };
#pragma unroll 1
            for (i64 kt = 0; kt < KT; ++kt) {
                if (cw == 0) producer_step();
                const unsigned sa = ring + stage * STAGE_BYTES;
                const unsigned sb = sa + A_BYTES;
                if (kt == 0) {
                    mbar_wait(full0 + stage * 8, phase);
                    load_frags(fa[0], fb[0], sa, sb, 0);
                }
                const int nstage = (stage + 1 == STAGES) ? 0 : stage + 1;
                const unsigned nphase = (stage + 1 == STAGES) ? (phase ^ 1u) : phase;
#pragma unroll
                for (int ks = 0; ks < BK / 4; ++ks) {
                    constexpr int dummy = 0; (void)dummy;
                    const int cur = ks & 1;
                    if (ks + 1 < BK / 4) {
                        load_frags(fa[cur ^ 1], fb[cur ^ 1], sa, sb, ks + 1);
                    } else if (kt + 1 < KT) {
                        mbar_wait(full0 + nstage * 8, nphase);
                        const unsigned na = ring + nstage * STAGE_BYTES;
                        load_frags(fa[cur ^ 1], fb[cur ^ 1], na, na + A_BYTES, 0);
                    }
#pragma unroll
                    for (int i = 0; i < FM; ++i)
#pragma unroll
                        for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[cur][i], fb[cur][j]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty0 + stage * 8);
                if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            }
        } else
#pragma unroll 1
        for (i64 kt = 0; kt < KT; ++kt) {
            if (stagger && group == 0 && !released && 2 * kt + 1 >= KT) {
                __syncwarp();
                if (lane == 0) mbar_arrive(go_bar);
                released = true;
            }
            if constexpr (MODE == 6) ptk = clock64();
            if (cw == 0) producer_step();
            if constexpr (MODE == 6) { const long long t = clock64(); pc[0] += t - ptk; ptk = t; }
            mbar_wait(full0 + stage * 8, phase);
            if constexpr (MODE == 6) { const long long t = clock64(); pc[1] += t - ptk; ptk = t; }
            const unsigned sa = ring + stage * STAGE_BYTES;
            const unsigned sb = sa + A_BYTES;
#pragma unroll
            for (int ks = 0; ks < BK / 4; ++ks) {
                double a[FM], b[FN];
#pragma unroll
                for (int i = 0; i < FM; ++i) {
                    if (A_KMAJOR) a[i] = lds64(sa + aoff[ks] + (unsigned)i * 1024u);
                    else a[i] = lds64(sa + aoff[(i & 1) + 2 * (ks & 1)] + (unsigned)(i >> 1) * 2048u + (unsigned)ks * 512u);
                }
#pragma unroll
                for (int j = 0; j < FN; ++j) {
                    if (B_KMAJOR) b[j] = lds64(sb + boff[ks] + (unsigned)j * 1024u);
                    else b[j] = lds64(sb + boff[(j & 1) + 2 * (ks & 1)] + (unsigned)(j >> 1) * 2048u + (unsigned)ks * 512u);
                }
#pragma unroll
                for (int i = 0; i < FM; ++i)
#pragma unroll
                    for (int j = 0; j < FN; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + stage * 8);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            if constexpr (MODE == 6) pc[2] += clock64() - ptk;
        }
        if constexpr (MODE == 6) { ptk = clock64(); pc[5] += 1; }

        // ---- epilogue: C = alpha*acc + beta*C; C was prefetched into L2 by the producer ----
        bool interior = (m0 + TM <= p.m) && (n0 + TN <= p.n);
        if (MODE == 1) interior = interior && (p.gi0 + m0 * p.gis >= p.gj0 + (n0 + TN - 1) * p.gjs);
        if (MODE == 2) interior = interior && (p.gi0 + (m0 + TM - 1) * p.gis <= p.gj0 + n0 * p.gjs);
        if (interior && !(p.flags & 2)) {
            // fast path (all but the edge / diagonal tiles): no masks, one base pointer per column
            // and compile-time row offsets; two memory round trips per tile
            double* cbase = p.C + (m0 + wm0 + tile_row<A_KMAJOR>(0, g)) + (n0 + wn0) * p.ldc;
            if constexpr (MODE == 3) {
                double sink = 0.0;   // keeps the accumulators alive
#pragma unroll
                for (int i = 0; i < FM; ++i)
#pragma unroll
                    for (int j = 0; j < FN; ++j) sink += acc[i][j][0] + acc[i][j][1];
                if (sink == 1.2345e300) cbase[0] = sink;
            } else if constexpr (MODE == 4) {
#pragma unroll
                for (int j = 0; j < FN; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double* cptr = cbase + (i64)tile_row<B_KMAJOR>(j, 2 * t + e) * p.ldc;
#pragma unroll
                        for (int i = 0; i < FM; ++i) cptr[tile_row<A_KMAJOR>(i, 0)] = __dmul_rn(alpha, acc[i][j][e]);
                    }
            } else if (useRed) {
#pragma unroll
                for (int j = 0; j < FN; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double* cptr = cbase + (i64)tile_row<B_KMAJOR>(j, 2 * t + e) * p.ldc;
#pragma unroll
                        for (int i = 0; i < FM; ++i) red_add_f64(cptr + tile_row<A_KMAJOR>(i, 0), __dmul_rn(alpha, acc[i][j][e]));
                    }
            } else
#pragma unroll
            for (int jj = 0; jj < FN; jj += 2) {
                double old[2][2][FM];
                if (useC) {
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2)
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const double* cptr = cbase + (i64)tile_row<B_KMAJOR>(jj + j2, 2 * t + e) * p.ldc;
#pragma unroll
                            for (int i = 0; i < FM; ++i) old[j2][e][i] = __ldcg(cptr + tile_row<A_KMAJOR>(i, 0));
                        }
                }
#pragma unroll
                for (int j2 = 0; j2 < 2; ++j2)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double* cptr = cbase + (i64)tile_row<B_KMAJOR>(jj + j2, 2 * t + e) * p.ldc;
#pragma unroll
                        for (int i = 0; i < FM; ++i) {
                            double v = __dmul_rn(alpha, acc[i][jj + j2][e]);
                            if (useC) v = __fma_rn(beta, old[j2][e][i], v);
                            cptr[tile_row<A_KMAJOR>(i, 0)] = v;
                        }
                    }
            }
        } else {
#pragma unroll
            for (int j = 0; j < FN; ++j) {
                double old[2][FM];
                bool ok[2][FM];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const i64 col = n0 + wn0 + tile_row<B_KMAJOR>(j, 2 * t + e);
                    const i64 gj = p.gj0 + col * p.gjs;
                    const double* cptr = p.C + col * p.ldc;
#pragma unroll
                    for (int i = 0; i < FM; ++i) {
                        const i64 row = m0 + wm0 + tile_row<A_KMAJOR>(i, g);
                        bool v = (col < p.n) && (row < p.m);
                        if (MODE == 1) v = v && (p.gi0 + row * p.gis >= gj);
                        if (MODE == 2) v = v && (p.gi0 + row * p.gis <= gj);
                        ok[e][i] = v;
                        old[e][i] = (v && useC && !useRed) ? __ldcg(cptr + row) : 0.0;
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const i64 col = n0 + wn0 + tile_row<B_KMAJOR>(j, 2 * t + e);
                    double* cptr = p.C + col * p.ldc;
#pragma unroll
                    for (int i = 0; i < FM; ++i) {
                        const i64 row = m0 + wm0 + tile_row<A_KMAJOR>(i, g);
                        if (ok[e][i]) {
                            double v = __dmul_rn(alpha, acc[i][j][e]);
                            if (useRed) { red_add_f64(cptr + row, v); continue; }
                            if (useC) v = __fma_rn(beta, old[e][i], v);
                            cptr[row] = v;
                        }
                    }
                }
            }
        }
        if constexpr (MODE == 6) pc[3] += clock64() - ptk;
    }
    if constexpr (MODE == 6) {
        pc[4] = clock64() - pt0;
        if (lane == 0 && p.prof) {
            unsigned long long* out = p.prof + ((size_t)blockIdx.x * (GROUPS * CONSUMER_WARPS) + warp) * 8;
            for (int i = 0; i < 6; ++i) out[i] = (unsigned long long)pc[i];
        }
    }
    // a group that never ran a tile must still release the other one
    if (stagger && group == 0 && !released) {
        __syncwarp();
        if (lane == 0) mbar_arrive(go_bar);
    }
}

// ---- host side ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    });
    return fn;
}

// 2-D f64 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, outer stride ld
void make_map(CUtensorMap* map, const double* ptr, i64 inner, i64 outer, i64 ld, int boxInner, int boxOuter) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8u};
    cuuint32_t box[2] = {(cuuint32_t)boxInner, (cuuint32_t)boxOuter};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
}

template <class CF, bool AK, bool BKM, int MODE>
void launch(const TmaArgs& a, double flops, cudaStream_t s) {
    static bool configured = false;
    auto kern = gemm_f64_tma_kernel<CF, AK, BKM, MODE>;
    if (!configured) {
        ELB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const i64 tiles = a.tilesM * a.tilesN;
    i64 grid = (tiles + GROUPS - 1) / GROUPS;
    if (grid > sm_count()) grid = sm_count();
    if (sm_limit() > 0 && grid > sm_limit()) grid = sm_limit();  // leave SMs to a concurrent panel stream
    gemm_profile_begin(s);
    kern<<<(unsigned)grid, CF::NTHREADS, SMEM_BYTES, s>>>(a);
    ELB_LAUNCH_CHECK();
    gemm_profile_end(s, flops);
}

template <class CF, int MODE>
void dispatch_cf(bool ak, bool bk, const TmaArgs& a, double flops, cudaStream_t s) {
    if (ak) {
        if (bk) launch<CF, true, true, MODE>(a, flops, s);
        else launch<CF, true, false, MODE>(a, flops, s);
    } else {
        if (bk) launch<CF, false, true, MODE>(a, flops, s);
        else launch<CF, false, false, MODE>(a, flops, s);
    }
}
template <int MODE>
void dispatch(bool ak, bool bk, const TmaArgs& a, double flops, cudaStream_t s) {
    if (a.flags & 8) dispatch_cf<Cfg4, MODE>(ak, bk, a, flops, s);
    else dispatch_cf<Cfg8, MODE>(ak, bk, a, flops, s);
}

}  // namespace

// Returns false (nothing launched) when the operands do not meet TMA's alignment rules.
// ta / tb: op(A) / op(B) is the transpose of the stored matrix.
bool dgemm_tma_device(int mode, bool ta, bool tb, i64 m, i64 n, i64 k, double alpha, const double* A, i64 lda,
                      const double* B, i64 ldb, double beta, double* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                      double flops, cudaStream_t s) {
    if (k <= 0 || m <= 0 || n <= 0) return false;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1)) return false;
    if (m >= (i64(1) << 31) || n >= (i64(1) << 31) || k >= (i64(1) << 31)) return false;
    if (ceil_div(m, TM) * ceil_div(n, TN) >= (i64(1) << 31) / GROUP_N) return false;
    if (!encode_fn()) return false;
    TmaArgs a;
    // A 'T' is K-major: stored k x m, k contiguous.  A 'N' is MN-major: stored m x k, m contiguous.
    const bool ak = ta, bk = !tb;
    if (ak) make_map(&a.mapA, A, k, m, lda, BK, TM);
    else make_map(&a.mapA, A, m, k, lda, 16, BK);
    if (bk) make_map(&a.mapB, B, k, n, ldb, BK, TN);
    else make_map(&a.mapB, B, n, k, ldb, 16, BK);
    a.m = m; a.n = n; a.k = k;
    a.C = C; a.ldc = ldc;
    a.alpha = alpha; a.beta = beta;
    a.gi0 = gi0; a.gis = gis; a.gj0 = gj0; a.gjs = gjs;
    a.tilesM = ceil_div(m, TM);
    a.tilesN = ceil_div(n, TN);
    a.flags = g_dgemm_tma_flags;
    a.prof = g_dgemm_tma_prof;
    if (mode == 0 && (a.flags & 1024) && !ak && bk) {  // diagnostic: per-warp phase clocks, NN only
        launch<Cfg8, false, true, 6>(a, flops, s);
        return true;
    }
    if (mode == 0 && (a.flags & 384) && !ak && bk) {   // diagnostic epilogues, NN only
        if (a.flags & 128) launch<Cfg8, false, true, 3>(a, flops, s);
        else launch<Cfg8, false, true, 4>(a, flops, s);
        return true;
    }
    if (mode == 0 && (a.flags & 512) && !ak && bk) {   // experiment: fragment prefetch one k-step ahead, NN only
        launch<Cfg8, false, true, 5>(a, flops, s);
        return true;
    }
    if (mode == 0) dispatch<0>(ak, bk, a, flops, s);
    else if (mode == 1) dispatch<1>(ak, bk, a, flops, s);
    else dispatch<2>(ak, bk, a, flops, s);
    return true;
}

}  // namespace elb200

namespace elb200 { extern int g_dgemm_ws_flags; extern unsigned long long* g_dgemm_ws_prof; }
extern "C" void elb200_dgemm_set_debug_flags(int f) { elb200::g_dgemm_tma_flags = f; elb200::g_dgemm_ws_flags = f; }
// device buffer of (grid * 16 warps * 8) u64 for the phase-clock diagnostic (flags bit 10)
extern "C" void elb200_dgemm_set_profile_buffer(void* p) {
    elb200::g_dgemm_tma_prof = (unsigned long long*)p;
    elb200::g_dgemm_ws_prof = (unsigned long long*)p;
}
