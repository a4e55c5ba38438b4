// FP64 GEMM / TRRK, third generation: persistent TMA-fed DMMA kernel with a non-blocking producer and a tile-info ring.
//
// Same contract as gemm_f64.cu (replaces blas::Gemm<double> -> dgemm_, reference
// src/core/imports/blas/Gemm.hpp:431, and with MODE != 0 the LocalTrrk recursion of
// src/blas_like/level3/Trrk/Local.hpp:782-830), built from what the per-warp phase clocks and the ncu
// source page of the second generation (round 1's gemm_f64_tma.cu) showed on the rank-Blocksize() update
// (profiles/r02_dgemm_phase_clocks_gen2.txt, r02_dmma_loop_variants.txt):
//   * the warp that issued the TMA loads spent 20 % of its time doing so (9 UTMALDG per k-stage, each behind
//     an ELECT / R2UR sequence) and, being a consumer too, was the slowest warp of its group: every other
//     warp then waited 8-18 % of its time for stages that were issued late;
//   * every FP64 ALU instruction (DMUL by alpha, DSETP on beta) queues behind the DMMAs of the same SM
//     sub-partition: the 64 DMULs of an epilogue cost as much as the 64 REDs;
//   * ~3200 clocks per tile went to the tile bookkeeping every consumer warp repeated (two integer divisions,
//     spilled 64-bit cursor state, the staircase test).
// Hence (a 17th, dedicated producer warp does not fit: the fifth warp of a sub-partition caps every thread at 96
// registers, and the 64-register accumulator tile then spills inside the k loop):
//   * 16 warps per CTA, one CTA per SM: two consumer groups of 8 warps (group tile 128 x 64, warp tile 32 x 32,
//     64 accumulator registers), each with its own 4-stage x 24 KB operand ring.  Warp 0 of a group is still the
//     group's producer, but production is now cheap and never blocks unless the warp itself would starve:
//     before every k-step it TESTS (mbarrier.test_wait) whether the next ring slot has been released and, if so,
//     issues the loads; it only waits for a slot when it has consumed everything it issued.
//   * the producer is the only warp that knows about tiles: it walks the (masked) tile sequence and publishes
//     {row0, col0, interior?} of every tile in a small shared-memory ring that the consumers read after the
//     tile's first full-barrier wait.  A slot with valid == 0 ends the group.  The other 7 warps of a group run
//     no tile arithmetic at all.
//   * MN-major operands (A 'N', B 'T') arrive as ONE 3-D box [16-row chunk][16 k][16 rows] per operand
//     and stage (tensor map dims {16 rows, k, rows / 16} with strides {8 B, ld, 128 B}) instead of 8 / 4
//     2-D boxes: 2 UTMALDG per stage for NN instead of 9.  Needs rows % 16 == 0; otherwise 2-D boxes.
//   * no FP64 ALU work in the rank-k epilogue when alpha = +-1 (sign flip on the integer pipe), beta is
//     classified on the host.
// Shared-memory layouts, the conflict-free fragment reads with their row permutations and the L2 reduction
// epilogue (red.global.add.f64: C += alpha acc performed by the L2 atomic unit, one writer per entry, hence
// deterministic) are those of the second generation:
//   K-major operand (A 'T' / B 'N'): one box [rows][16 k], row pitch 128 B, 16-byte chunk index XORed with
//     (row & 7); an m8n8k4 fragment takes its 8 rows in the order {0,2,4,6,1,3,5,7}.
//   MN-major operand (A 'N' / B 'T'): boxes [16 k][16 rows], pitch 128 B, chunk XORed with (k & 7); a fragment
//     takes rows {0,1,8,9,2,3,10,11} (+4 for the second fragment of the box).
// TMA zero-fills out-of-range rows / k; the epilogue masks (and applies the global staircase mask for TRRK).
// Requirements: A, B 16-byte aligned with even leading dimension, else gemm_f64.cu serves the call.
#include <cuda.h>

#include <mutex>
#include <type_traits>

#include "../common.hpp"
#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
int g_dgemm_ws_flags = 0;                       // bit 4: never use the L2 reduction epilogue (tests); bit 10: phase clocks
unsigned long long* g_dgemm_ws_prof = nullptr;  // device buffer for the phase clocks
namespace {

constexpr int BK = 16;             // doubles per k-stage = one 128-byte swizzle span
constexpr int TM = 128, TN = 64;   // C tile of one consumer group
constexpr int STAGES = 4;
constexpr int GROUPS = 2;
constexpr int CW = 8;              // consumer warps per group: 4 (rows) x 2 (columns), warp tile 32 x 32
constexpr int WM = 4;
constexpr int FM = 4, FN = 4;      // 8-row A fragments / 8-column B fragments per warp
constexpr int NTHREADS = 32 * GROUPS * CW;
constexpr int A_BYTES = TM * BK * 8;  // 16 KB
constexpr int B_BYTES = TN * BK * 8;  //  8 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int RING_BYTES = STAGES * STAGE_BYTES;
constexpr int INFO_SLOTS = 8;      // > STAGES + 1: a slot is rewritten only after its tile's epilogue started
constexpr int BAR_BYTES = GROUPS * 2 * STAGES * 8;
constexpr int INFO_BYTES = GROUPS * INFO_SLOTS * 16;
constexpr int PST_BYTES = GROUPS * 64;  // producer cursors
constexpr unsigned TILE_CHUNK = 4;     // tiles per atomicAdd on the tile counter
constexpr int SMEM_BYTES = GROUPS * RING_BYTES + 1024 /*alignment slack*/ + BAR_BYTES + INFO_BYTES + PST_BYTES;
constexpr int GROUP_N = 16;        // tile columns per rasterisation band (in 64-column tiles)

struct WsArgs {
    CUtensorMap mapA, mapB;
    i64 m, n, k;
    double* C;
    i64 ldc;
    double alpha, beta;
    i64 gi0, gis, gj0, gjs;
    i64 tilesM, tilesN;
    unsigned long long* prof;
    int a3d, b3d;     // MN-major operand described by the 3-D single-box tensor map
    int epi;          // 0: C += alpha acc (L2 reduction), 1: C = alpha acc, 2: C = alpha acc + beta C
    int alphaMode;    // 0: general, 1: alpha == 1, 2: alpha == -1
    int bandw;        // tile columns per rasterisation band
    unsigned staggerNs;  // group 1 starts this much later than group 0
    int rowPair;      // 1: rows 2i, 2i + 1 of C are (re, im) of complex row i: the staircase mask uses row >> 1
    unsigned* tileCounter;  // zeroed before the launch: tiles are handed out in raster order, TILE_CHUNK at a time
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
        "WS_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WS_WAIT_DONE;\n"
        "bra WS_WAIT_LOOP;\n"
        "WS_WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// non-blocking test / one bounded (hardware time limit) wait
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_try_once(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
        : "memory");
}
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
// x with its sign bit XORed by `flip` (0 or 0x80000000): negation on the integer pipe
__device__ __forceinline__ double flip_sign(double x, unsigned flip) {
    return __hiloint2double(__double2hiint(x) ^ (int)flip, __double2loint(x));
}

// ---- row permutations ----
// K-major operand: fragment f (8 rows) of a warp takes rows base + 8 f + PK[x], x = MMA row index
// MN-major operand: fragment f takes rows base + 16 (f >> 1) + 4 (f & 1) + QM[x]
__device__ __forceinline__ int permK(int x) { return ((x & 3) << 1) | (x >> 2); }
__device__ __forceinline__ int permM(int x) { return ((x & 2) << 2) | (x & 1) | ((x & 4) >> 1); }  // {0,1,8,9,2,3,10,11}
template <bool KMAJOR>
__device__ __forceinline__ int tile_row(int f, int x) {
    return KMAJOR ? (8 * f + permK(x)) : (16 * (f >> 1) + 4 * (f & 1) + permM(x));
}

// tile index -> (tm, tn): bands of GROUP_N tile columns, walked down the rows.  Producer only.
__device__ __forceinline__ void tile_coords(const WsArgs& p, unsigned tile, unsigned& tm, unsigned& tn) {
    const unsigned tilesM = (unsigned)p.tilesM, tilesN = (unsigned)p.tilesN;
    const unsigned gn = (unsigned)p.bandw;
    const unsigned band_sz = gn * tilesM;
    const unsigned band = tile / band_sz;
    const unsigned first_n = band * gn;
    const unsigned bw = (tilesN - first_n < gn) ? (tilesN - first_n) : gn;
    const unsigned in_band = tile - band * band_sz;
    const unsigned q = in_band / bw;
    tm = q;
    tn = first_n + (in_band - q * bw);
}

// 0: no entry of the tile lies in the triangle (skipped); 1: some do (masked epilogue); 2: the whole tile is
// inside the matrix and inside the triangle (unmasked epilogue)
template <int MODE>
__device__ __forceinline__ int tile_class(const WsArgs& p, i64 m0, i64 n0) {
    const bool full = (m0 + TM <= p.m) && (n0 + TN <= p.n);
    if (MODE == 0) return full ? 2 : 1;
    const i64 mlast = (m0 + TM - 1 < p.m - 1) ? (m0 + TM - 1) : (p.m - 1);
    const i64 nlast = (n0 + TN - 1 < p.n - 1) ? (n0 + TN - 1) : (p.n - 1);
    const int rp = p.rowPair;   // complex rows stored as (re, im) row pairs
    const i64 rfirst = m0 >> rp, rlast = mlast >> rp, rend = (m0 + TM - 1) >> rp;
    if (MODE == 1) {
        if (!(p.gi0 + rlast * p.gis >= p.gj0 + n0 * p.gjs)) return 0;                     // no gi >= gj
        return (full && p.gi0 + rfirst * p.gis >= p.gj0 + (n0 + TN - 1) * p.gjs) ? 2 : 1;  // all gi >= gj
    }
    if (!(p.gi0 + rfirst * p.gis <= p.gj0 + nlast * p.gjs)) return 0;
    return (full && p.gi0 + rend * p.gis <= p.gj0 + n0 * p.gjs) ? 2 : 1;
}

// ---- epilogue of one finished group tile (all 8 warps of the group) ----
template <bool A_KMAJOR, bool B_KMAJOR, int MODE>
__device__ __forceinline__ void epilogue(const WsArgs& p, double (&acc)[FM][FN][2], int im0, int in0, int iflags,
                                         int wm0, int wn0, int g, int t) {
    const i64 m0 = im0, n0 = in0;
    const int epi = p.epi;
    const double alpha = p.alpha;
    if (iflags & 2) {
        // unmasked: one base pointer per column and compile-time row offsets
        double* cbase = p.C + (m0 + wm0 + tile_row<A_KMAJOR>(0, g)) + (n0 + wn0) * p.ldc;
        if (epi == 0) {
            // three separately compiled loops (a runtime select would keep the DMUL in all of them, and every
            // FP64 ALU instruction queues behind the DMMAs of its sub-partition)
            auto red_tile = [&](auto AM) {
                constexpr int am = decltype(AM)::value;
#pragma unroll
                for (int j = 0; j < FN; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double* cptr = cbase + (i64)tile_row<B_KMAJOR>(j, 2 * t + e) * p.ldc;
#pragma unroll
                        for (int i = 0; i < FM; ++i) {
                            double v = acc[i][j][e];
                            if (am == 2) v = flip_sign(v, 0x80000000u);
                            if (am == 0) v = __dmul_rn(alpha, v);
                            red_add_f64(cptr + tile_row<A_KMAJOR>(i, 0), v);
                        }
                    }
            };
            if (p.alphaMode == 1) red_tile(std::integral_constant<int, 1>());
            else if (p.alphaMode == 2) red_tile(std::integral_constant<int, 2>());
            else red_tile(std::integral_constant<int, 0>());
        } else if (epi == 1) {
            // C = alpha acc: plain stores, again without FP64 ALU work for alpha = +-1
            auto store_tile = [&](auto AM) {
                constexpr int am = decltype(AM)::value;
#pragma unroll
                for (int j = 0; j < FN; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double* cptr = cbase + (i64)tile_row<B_KMAJOR>(j, 2 * t + e) * p.ldc;
#pragma unroll
                        for (int i = 0; i < FM; ++i) {
                            double v = acc[i][j][e];
                            if (am == 2) v = flip_sign(v, 0x80000000u);
                            if (am == 0) v = __dmul_rn(alpha, v);
                            cptr[tile_row<A_KMAJOR>(i, 0)] = v;
                        }
                    }
            };
            if (p.alphaMode == 1) store_tile(std::integral_constant<int, 1>());
            else if (p.alphaMode == 2) store_tile(std::integral_constant<int, 2>());
            else store_tile(std::integral_constant<int, 0>());
        } else {
            const double beta = p.beta;
#pragma unroll
            for (int jj = 0; jj < FN; jj += 2) {
                double old[2][2][FM];
#pragma unroll
                for (int j2 = 0; j2 < 2; ++j2)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const double* cptr = cbase + (i64)tile_row<B_KMAJOR>(jj + j2, 2 * t + e) * p.ldc;
#pragma unroll
                        for (int i = 0; i < FM; ++i) old[j2][e][i] = __ldcg(cptr + tile_row<A_KMAJOR>(i, 0));
                    }
#pragma unroll
                for (int j2 = 0; j2 < 2; ++j2)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        double* cptr = cbase + (i64)tile_row<B_KMAJOR>(jj + j2, 2 * t + e) * p.ldc;
#pragma unroll
                        for (int i = 0; i < FM; ++i)
                            cptr[tile_row<A_KMAJOR>(i, 0)] = __fma_rn(beta, old[j2][e][i], __dmul_rn(alpha, acc[i][jj + j2][e]));
                    }
            }
        }
    } else {
        const double beta = p.beta;
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
                    if (MODE == 1) v = v && (p.gi0 + (row >> p.rowPair) * p.gis >= gj);
                    if (MODE == 2) v = v && (p.gi0 + (row >> p.rowPair) * p.gis <= gj);
                    ok[e][i] = v;
                    old[e][i] = (v && epi == 2) ? __ldcg(cptr + row) : 0.0;
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
                        if (epi == 0) { red_add_f64(cptr + row, v); continue; }
                        if (epi == 2) v = __fma_rn(beta, old[e][i], v);
                        cptr[row] = v;
                    }
                }
            }
        }
    }
}

// MODE 0: full GEMM; 1: lower-triangle TRRK; 2: upper-triangle TRRK.  PROF: phase-clock diagnostic build.
template <bool A_KMAJOR, bool B_KMAJOR, int MODE, bool PROF>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_f64_ws_kernel(const __grid_constant__ WsArgs p) {
    extern __shared__ unsigned char smem_raw[];
    const unsigned raw = smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;       // 1024-byte alignment for the 128 B swizzle
    const unsigned bars = base + GROUPS * RING_BYTES;   // full[G][S], empty[G][S]
    const unsigned infos = bars + BAR_BYTES;            // info[G][INFO_SLOTS]: {row0, col0, flags, -}
    const unsigned psts = infos + INFO_BYTES;           // producer cursor of each group (32 B)

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    const int group = warp / CW;
    const int cw = warp % CW;
    const unsigned ring = base + group * RING_BYTES;
    const unsigned full0 = bars + (group * 2 * STAGES) * 8;
    const unsigned empty0 = full0 + STAGES * 8;
    const unsigned info0 = infos + group * INFO_SLOTS * 16;
    const unsigned pst0 = psts + group * 64;
    const int KT = (int)((p.k + BK - 1) / BK);

    // ---- producer (warp 0 of the group).  Its cursor {tile, row0, col0, class | k-stage, ring slot, phase, tile
    //      sequence number} lives in SHARED memory: kept in registers it would be live across the k loop of
    //      every warp, and the compiler then recomputes the fragment addresses at each k-step instead ----
    // Next tile with work, drawn from the launch-wide tile counter (all CTAs and both groups share it, so the tile
    // order is the raster order and nobody is left with more than a few tiles more than anybody else: with the
    // static round-robin of the previous versions the staircase of a TRRK left some CTAs up to 12 % more tiles).
    // `cur` / `end` are the unconsumed part of the chunk this warp holds.  Returns total when nothing is left.
    // Runs on all lanes of warp 0 (lane 0 does the atomic).
    auto seek = [&](unsigned& cur, unsigned& end, int& m0, int& n0, int& cls) -> unsigned {
        const unsigned total = (unsigned)(p.tilesM * p.tilesN);
        for (;;) {
            if (cur >= end) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(p.tileCounter, TILE_CHUNK);
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base >= total) { cur = end = total; return total; }
                cur = base;
                end = (base + TILE_CHUNK < total) ? base + TILE_CHUNK : total;
            }
            const unsigned tl = cur++;
            unsigned tm, tn;
            tile_coords(p, tl, tm, tn);
            const int c = tile_class<MODE>(p, (i64)tm * TM, (i64)tn * TN);
            if (c != 0) { m0 = (int)(tm * TM); n0 = (int)(tn * TN); cls = c; return tl; }
        }
    };
    int ahead = 0;        // stages issued and not yet consumed by this warp (warp 0 only)
    bool pdone = false;   // the end-of-sequence slot has been issued
    long long pbusy = 0;
    // issue one k-stage (or the end-of-sequence slot) into the next ring slot.  blocking: wait for the slot to be
    // released; else give up when it has not been yet.  All lanes of warp 0 run it, lane 0 issues.
    auto produce = [&](bool blocking) {
        unsigned ptile; int pm0, pn0, pcls, pkt, pstage, pphase, ptseq;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(ptile), "=r"(pm0), "=r"(pn0), "=r"(pcls) : "r"(pst0) : "memory");
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(pkt), "=r"(pstage), "=r"(pphase), "=r"(ptseq) : "r"(pst0 + 16) : "memory");
        const unsigned full = full0 + pstage * 8;
        const unsigned empty = empty0 + pstage * 8;
        if (blocking) mbar_wait(empty, (unsigned)pphase ^ 1u);
        else if (!mbar_test(empty, (unsigned)pphase ^ 1u)) return;
        long long tb = 0;
        if (PROF) tb = clock64();
        const unsigned slot = info0 + (ptseq & (INFO_SLOTS - 1)) * 16;
        const unsigned total = (unsigned)(p.tilesM * p.tilesN);
        ++ahead;
        if (ptile >= total) {
            // end of the tile sequence: a slot with valid == 0 behind a plain arrive
            if (lane == 0) {
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
                mbar_arrive(full);
            }
            __syncwarp();
            pdone = true;
            return;
        }
        if (lane == 0) {
            if (pkt == 0)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(pm0), "r"(pn0),
                             "r"(1 | (pcls == 2 ? 2 : 0)), "r"(0) : "memory");
            const unsigned sa = ring + pstage * STAGE_BYTES;
            const unsigned sb = sa + A_BYTES;
            mbar_expect_tx(full, STAGE_BYTES);
            const int k0 = pkt * BK;
            if (A_KMAJOR) {
                tma_load_2d(sa, &p.mapA, k0, pm0, full);
            } else if (p.a3d) {
                tma_load_3d(sa, &p.mapA, 0, k0, pm0 >> 4, full);
            } else {
#pragma unroll
                for (int b = 0; b < TM / 16; ++b) tma_load_2d(sa + b * 2048, &p.mapA, pm0 + 16 * b, k0, full);
            }
            if (B_KMAJOR) {
                tma_load_2d(sb, &p.mapB, k0, pn0, full);
            } else if (p.b3d) {
                tma_load_3d(sb, &p.mapB, 0, k0, pn0 >> 4, full);
            } else {
#pragma unroll
                for (int b = 0; b < TN / 16; ++b) tma_load_2d(sb + b * 2048, &p.mapB, pn0 + 16 * b, k0, full);
            }
        }
        if (++pstage == STAGES) { pstage = 0; pphase ^= 1; }
        if (++pkt == KT) {
            pkt = 0;
            ++ptseq;
            unsigned cur, end;
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(cur), "=r"(end) : "r"(pst0 + 32) : "memory");
            ptile = seek(cur, end, pm0, pn0, pcls);
            if (lane == 0) asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(pst0 + 32), "r"(cur), "r"(end) : "memory");
        }
        if (lane == 0) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pst0), "r"(ptile), "r"(pm0), "r"(pn0), "r"(pcls) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pst0 + 16), "r"(pkt), "r"(pstage), "r"(pphase), "r"(ptseq) : "memory");
        }
        __syncwarp();
        if (PROF) pbusy += clock64() - tb;
    };

    if (tid == 0) {
        for (int g2 = 0; g2 < GROUPS; ++g2)
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(bars + ((g2 * 2 * STAGES) + s) * 8, 1);
                mbar_init(bars + ((g2 * 2 * STAGES) + STAGES + s) * 8, CW);
            }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
    long long pt0 = 0;
    if (PROF) pt0 = clock64();
    // The two groups run the same code on equally long tiles: started together they stay in step and run their
    // epilogues at the same time, leaving the tensor pipe idle.  Group 1 therefore starts half a tile late.
    if (group == 1 && p.staggerNs) {
        for (unsigned waited = 0; waited < p.staggerNs; waited += 500u) __nanosleep(500u);
    }
    if (cw == 0) {
        // the group's first tile (after the stagger: a late group simply draws later tiles), then fill the ring
        int m0 = 0, n0 = 0, cls = 0;
        unsigned cur = 0, end = 0;
        const unsigned tl = seek(cur, end, m0, n0, cls);
        if (lane == 0) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pst0), "r"(tl), "r"(m0), "r"(n0), "r"(cls) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(pst0 + 16), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
            asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(pst0 + 32), "r"(cur), "r"(end) : "memory");
        }
        __syncwarp();
#pragma unroll 1
        for (int i = 0; i < STAGES && !pdone; ++i) produce(false);  // the ring starts empty: fill it
    }

    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (cw % WM) * (8 * FM), wn0 = (cw / WM) * (8 * FN);

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

    // opaque to the compiler: it must keep the eight offsets in registers instead of recomputing them from the
    // lane id (a dependent chain of ~25 integer instructions) in front of every stage's first fragment loads
#pragma unroll
    for (int x = 0; x < 4; ++x) {
        asm volatile("" : "+r"(aoff[x]));
        asm volatile("" : "+r"(boff[x]));
    }

    // ---- one loop over k-stages; a tile is KT consecutive stages ----
    int stage = 0, kt = 0, tseq = 0;
    unsigned phase = 0;
    // PROF: SM clocks of this warp in {0 issuing loads, 1 waiting for a full stage, 2 fragment loads + DMMAs,
    // 3 epilogue, 4 whole kernel}, 5 = tiles
    long long pc1 = 0, pc2 = 0, pc3 = 0, ptiles = 0, ptk = 0;
    double acc[FM][FN][2];

#pragma unroll 1
    for (;;) {
        if (PROF) ptk = clock64();
        if (cw == 0 && !pdone) produce(ahead == 0);   // blocks only if this warp consumed all it issued
        mbar_wait(full0 + stage * 8, phase);
        if (PROF) { const long long tt = clock64(); pc1 += tt - ptk; ptk = tt; }
        if (kt == 0) {
            int iflags;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(iflags) : "r"(info0 + (tseq & (INFO_SLOTS - 1)) * 16 + 8) : "memory");
            if (!(iflags & 1)) break;
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < FN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        }
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
        --ahead;
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if (PROF) pc2 += clock64() - ptk;
        if (++kt == KT) {
            kt = 0;
            if (PROF) { ptk = clock64(); ++ptiles; }
            // the tile's slot is still intact: the producer is at most STAGES stages (<= STAGES tiles) ahead
            int im0, in0, iflags, ipad;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(im0), "=r"(in0), "=r"(iflags), "=r"(ipad)
                         : "r"(info0 + (tseq & (INFO_SLOTS - 1)) * 16) : "memory");
            ++tseq;
            epilogue<A_KMAJOR, B_KMAJOR, MODE>(p, acc, im0, in0, iflags, wm0, wn0, g, t);
            if (PROF) pc3 += clock64() - ptk;
        }
    }
    if (PROF && lane == 0 && p.prof) {
        unsigned long long* out = p.prof + ((size_t)blockIdx.x * (GROUPS * CW) + warp) * 8;
        out[0] = (unsigned long long)pbusy;
        out[1] = (unsigned long long)pc1; out[2] = (unsigned long long)pc2; out[3] = (unsigned long long)pc3;
        out[4] = (unsigned long long)(clock64() - pt0); out[5] = (unsigned long long)ptiles;
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
void make_map2(CUtensorMap* map, const double* ptr, i64 inner, i64 outer, i64 ld, int boxInner, int boxOuter) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8u};
    cuuint32_t box[2] = {(cuuint32_t)boxInner, (cuuint32_t)boxOuter};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)ptr, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled failed (code " + std::to_string((int)r) + ")");
}
// MN-major operand stored rows x k (rows contiguous, rows % 16 == 0) as {16 rows, k, rows / 16} with strides
// {8 B, ld, 128 B}: one box {16, BK, chunks} lands as [chunk][k][16 rows], i.e. `chunks` of the 2-D boxes above
// back to back.  False when the driver rejects the (non-monotonic) strides: the caller then uses 2-D boxes.
bool make_map3(CUtensorMap* map, const double* ptr, i64 rows, i64 k, i64 ld, int chunks) {
    cuuint64_t dims[3] = {16, (cuuint64_t)k, (cuuint64_t)(rows / 16)};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 8u, 128u};
    cuuint32_t box[3] = {16, (cuuint32_t)BK, (cuuint32_t)chunks};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void*)ptr, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

int g_last_3d = 0;  // bit 0: A used the 3-D map, bit 1: B did (elb200_dgemm_ws_last_maps)

// one zeroed tile counter per launch, from a ring long enough that a counter is never reused while an earlier
// launch that used it can still be running (launches on different streams overlap at most a few deep)
unsigned* next_tile_counter(cudaStream_t s) {
    constexpr int RING = 4096;
    static unsigned* ring = nullptr;
    static unsigned long long used = 0;
    if (!ring) ELB_CUDA(cudaMalloc((void**)&ring, RING * sizeof(unsigned)));
    unsigned* c = ring + (used++ % RING);
    ELB_CUDA(cudaMemsetAsync(c, 0, sizeof(unsigned), s));
    return c;
}

template <bool AK, bool BKM, int MODE, bool PROF>
void launch(const WsArgs& a0, double flops, cudaStream_t s) {
    WsArgs a = a0;
    a.tileCounter = next_tile_counter(s);
    static bool configured = false;
    auto kern = gemm_f64_ws_kernel<AK, BKM, MODE, PROF>;
    if (!configured) {
        ELB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const i64 tiles = a.tilesM * a.tilesN;
    i64 grid = (tiles + GROUPS - 1) / GROUPS;
    if (grid > sm_count()) grid = sm_count();
    if (sm_limit() > 0 && grid > sm_limit()) grid = sm_limit();  // leave SMs to a concurrent panel stream
    gemm_profile_begin(s);
    kern<<<(unsigned)grid, NTHREADS, SMEM_BYTES, s>>>(a);
    ELB_LAUNCH_CHECK();
    gemm_profile_end(s, flops);
}

template <int MODE>
void dispatch(bool ak, bool bk, const WsArgs& a, double flops, cudaStream_t s) {
    if (ak) {
        if (bk) launch<true, true, MODE, false>(a, flops, s);
        else launch<true, false, MODE, false>(a, flops, s);
    } else {
        if (bk) launch<false, true, MODE, false>(a, flops, s);
        else launch<false, false, MODE, false>(a, flops, s);
    }
}

}  // namespace

// Returns false (nothing launched) when the operands do not meet TMA's alignment rules.
// ta / tb: op(A) / op(B) is the transpose of the stored matrix.
bool dgemm_ws_device(int mode, bool ta, bool tb, i64 m, i64 n, i64 k, double alpha, const double* A, i64 lda,
                     const double* B, i64 ldb, double beta, double* C, i64 ldc, i64 gi0, i64 gis, i64 gj0, i64 gjs,
                     double flops, cudaStream_t s, int rowPair) {
    if (k <= 0 || m <= 0 || n <= 0) return false;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 1) || (ldb & 1)) return false;
    if (m >= (i64(1) << 31) - TM || n >= (i64(1) << 31) - TN || k >= (i64(1) << 31) - BK) return false;
    if (ceil_div(m, TM) * ceil_div(n, TN) >= (i64(1) << 31) / GROUP_N) return false;
    if (!encode_fn()) return false;
    // General alpha on a rank-k shaped update: every DMUL of the epilogue queues behind the DMMAs of its
    // sub-partition (27.5 instead of 34 TFLOP/s on the rank-128 update).  When one stored operand is small next to
    // C, scale a copy of it once -- O(k max(m, n)) on HBM -- and run the alpha = 1 epilogue: C += A (alpha B).
    double* scaled = nullptr;
    if (alpha != 0.0 && alpha != 1.0 && alpha != -1.0 && (beta == 1.0 || beta == 0.0)) {
        const bool useA = m <= n;   // the smaller stored operand
        const i64 rows = useA ? (ta ? k : m) : (tb ? n : k), cols = useA ? (ta ? m : k) : (tb ? k : n);
        if (8 * rows * cols <= m * n) {
            const i64 ld = rows + (rows & 1);
            scaled = (double*)scratch_alloc(sizeof(double) * (size_t)(ld * cols), s);
            lattice_copy_device<double>(useA ? A : B, scaled, rows, cols, 0, 1, useA ? lda : ldb, 0, 1, ld, false, &alpha, false, s);
            if (useA) { A = scaled; lda = ld; } else { B = scaled; ldb = ld; }
            alpha = 1.0;
        }
    }
    struct Release { double* p; cudaStream_t s; ~Release() { if (p) scratch_free(p, s); } } release{scaled, s};
    WsArgs a;
    // A 'T' is K-major: stored k x m, k contiguous.  A 'N' is MN-major: stored m x k, m contiguous.
    const bool ak = ta, bk = !tb;
    a.a3d = a.b3d = 0;
    if (ak) make_map2(&a.mapA, A, k, m, lda, BK, TM);
    else if (m % 16 == 0 && make_map3(&a.mapA, A, m, k, lda, TM / 16)) a.a3d = 1;
    else make_map2(&a.mapA, A, m, k, lda, 16, BK);
    if (bk) make_map2(&a.mapB, B, k, n, ldb, BK, TN);
    else if (n % 16 == 0 && make_map3(&a.mapB, B, n, k, ldb, TN / 16)) a.b3d = 1;
    else make_map2(&a.mapB, B, n, k, ldb, 16, BK);
    g_last_3d = a.a3d | (a.b3d << 1);
    a.m = m; a.n = n; a.k = k;
    a.C = C; a.ldc = ldc;
    a.alpha = alpha; a.beta = beta;
    a.gi0 = gi0; a.gis = gis; a.gj0 = gj0; a.gjs = gjs;
    a.rowPair = rowPair ? 1 : 0;
    a.tilesM = ceil_div(m, TM);
    a.tilesN = ceil_div(n, TN);
    a.prof = g_dgemm_ws_prof;
    a.epi = (beta == 1.0 && !(g_dgemm_ws_flags & 16)) ? 0 : (beta == 0.0 ? 1 : 2);
    a.alphaMode = (alpha == 1.0) ? 1 : (alpha == -1.0 ? 2 : 0);
    a.bandw = GROUP_N;
    // half-way (a little less measured best) into group 0's first tile: a tile pair takes ~2.1 us per k-stage
    {
        const i64 kt = ceil_div(k, BK);
        a.staggerNs = (unsigned)((kt < 16 ? kt : 16) * 750);
        if (a.tilesM * a.tilesN < 4 * (i64)sm_count()) a.staggerNs = 0;   // too few tiles for it to pay
        if ((g_dgemm_ws_flags >> 26) & 15) a.staggerNs = (unsigned)(((g_dgemm_ws_flags >> 26) & 15) - 1) * 1000u;  // tests: 1 = off
    }
    if ((g_dgemm_ws_flags >> 20) & 63) a.bandw = (g_dgemm_ws_flags >> 20) & 63;
    if (mode == 0 && (g_dgemm_ws_flags & 1024) && !ak && bk) {  // phase-clock diagnostic build, NN only
        launch<false, true, 0, true>(a, flops, s);
        return true;
    }
    if (mode == 0) dispatch<0>(ak, bk, a, flops, s);
    else if (mode == 1) dispatch<1>(ak, bk, a, flops, s);
    else dispatch<2>(ak, bk, a, flops, s);
    return true;
}

}  // namespace elb200

extern "C" {
// bit 0: A of the last launch used the single-box 3-D tensor map, bit 1: B did
int elb200_dgemm_ws_last_maps(void) { return elb200::g_last_3d; }
// bit 4: never use the L2 reduction epilogue; bit 10: phase-clock diagnostic build (NN only); bits 20..25: rasterisation
// band width; bits 26..29: 1 = no group stagger, v > 1 = stagger of (v - 1) microseconds
void elb200_dgemm_set_debug_flags(int f) { elb200::g_dgemm_ws_flags = f; }
// device buffer of (grid * 16 warps * 8) u64 for the phase-clock diagnostic
void elb200_dgemm_set_profile_buffer(void* p) { elb200::g_dgemm_ws_prof = (unsigned long long*)p; }
}
