// FP32 GEMM on the 5th-generation tensor cores: 3xTF32 split, tcgen05.mma kind::tf32 with the
// accumulator in tensor memory.
//
// Replaces blas::Gemm<float> -> sgemm_ (reference src/core/imports/blas/Gemm.hpp:388-400) for the
// float configuration of the hot path (BASELINE.json configs[4]: El::Gemm float, SUMMA_Dot).  The
// exact-FFMA alternative is gemm_simt.cu; which one El::Gemm<float> uses is a mode switch
// (elb200_sgemm_set_mode), never a silent substitution.
//
// Arithmetic.  Every fp32 operand x is split as x = hi + lo + r with hi = rn_tf32(x),
// lo = x - hi (exact in fp32; the tensor core ignores its low 13 bits), |r| <= 2^-22 |x|.  The product uses three tensor
// passes per k-step, small terms first:
//     D += A_lo B_hi;   D += A_hi B_lo;   D += A_hi B_hi          (A_lo B_lo ~ 2^-22 is dropped)
// so each a_ik b_kj carries a relative error of about 3 * 2^-22 before the fp32 accumulation in
// TMEM -- the same order as an FFMA chain of length k (tests state the tolerance against an FP64
// product).
//
// Structure (one persistent CTA per SM, 448 threads, 128 x 128 C tiles, k-blocks of 32):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (128-byte swizzle) of the raw fp32 A and B
//               k-blocks into a 3-stage ring, completion on an mbarrier;
//   warps 2-5   splitters: read the landed fp32 block, write hi in place and lo into a second
//               buffer of the same layout (the split is elementwise, so it is oblivious of the
//               swizzle), fence.proxy.async, arrive;
//   warp 1      one thread issues the tcgen05.mma's (M=128, N=128, K=8) on shared-memory matrix
//               descriptors; tcgen05.commit releases the ring stage and, after the last k-block,
//               publishes the accumulator; the accumulator is double-buffered in TMEM (2 x 128
//               columns) so the epilogue of tile i runs under the main loop of tile i+1;
//   warps 6-13  epilogue: tcgen05.ld 32x32b.x16 (a warp owns the 32-lane TMEM quadrant warp % 4 and
//               one half of the columns), C = alpha * acc + beta * C with rows on lanes (coalesced
//               in column-major C).
// The tensor core adds into the fp32 accumulator with truncation (measured: the error of a plain
// TMEM accumulation grows linearly with k).  The k loop is therefore cut into chunks of 1024: each
// chunk accumulates in TMEM, the epilogue warps promote it into registers with round-to-nearest
// adds (the two TMEM accumulators alternate between chunks), so the drift is bounded by the chunk
// length whatever k is.
// Operands may be K-major (A 'T', B 'N': k contiguous; plain 128-byte swizzle) or MN-major (A 'N',
// B 'T'; the 32-byte-atom flavour of the 128-byte swizzle, the only layout from which the tensor
// core transposes 32-bit operands -- see make_desc): TMA boxes, swizzle mode and descriptor differ.
// TMA zero-fills ragged m / n / k edges.
//
// Requirements: A, B 16-byte aligned, lda, ldb multiples of 4 (TMA strides are multiples of 16 B).
#include <cuda.h>

#include <mutex>

#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int BM = 128, BN = 128, BK = 32;   // fp32 elements; BK * 4 B = one 128-byte swizzle span
constexpr int STAGES = 3;
constexpr int OP_BYTES = 128 * BK * 4;       // one 128 x 32 fp32 operand block: 16 KB
constexpr int STAGE_BYTES = 4 * OP_BYTES;    // A hi | B hi | A lo | B lo
constexpr int SPLIT_WARPS = 4, EPI_WARPS = 8;
constexpr int KCHUNK = 32;                   // k-blocks (1024 k) accumulated in TMEM between two promotions
constexpr int NUM_THREADS = 32 * (2 + SPLIT_WARPS + EPI_WARPS);
constexpr int TMEM_COLS = 2 * BN;            // two fp32 accumulators of 128 lanes x 128 columns
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 256 /*barriers*/;

struct TcArgs {
    CUtensorMap mapA, mapB;
    i64 m, n, k;
    float* C;
    i64 ldc;
    float alpha, beta;
    i64 tilesM, tilesN;
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
        "TC_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TC_WAIT_DONE;\n"
        "bra TC_WAIT_LOOP;\n"
        "TC_WAIT_DONE:\n"
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
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc], issued by one thread on behalf of the CTA
__device__ __forceinline__ void umma_tf32(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// round to nearest (ties away, what cvt.rna.tf32.f32 does) with two full-rate integer operations
__device__ __forceinline__ unsigned rn_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xFFFFE000u; }

// Shared-memory matrix descriptor (tcgen05 / "UMMA"):
//   bits [0,14)  start address >> 4        bits [16,30) leading-dimension byte offset >> 4
//   bits [32,46) stride byte offset >> 4   bits [46,48) descriptor version (1 on sm_100)
//   bits [61,64) layout type (2 = SWIZZLE_128B, 1 = SWIZZLE_128B with 32-byte atoms)
// K-major block [rows][32 k]: rows are 128 B apart, 8-row swizzle atoms 1024 B apart (stride
// offset); the leading offset is not used by swizzled K-major layouts (1 by convention).
// MN-major block: boxes [32 k][32 rows], k rows 128 B apart.  32-bit operands can only be
// transposed by the tensor core from the 32-byte-atom flavour of the 128-byte swizzle (the
// plain SWIZZLE_128B MN-major descriptor silently multiplies by zero for kind::tf32 -- measured
// with scripts/micro/umma_probe.cu): 32-byte chunks XOR (k mod 4), atoms of 4 k = 512 B apart
// (stride offset), 32-row groups one box = 4096 B apart (leading offset).  TMA writes the same
// pattern with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
template <bool KMAJOR>
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr) {
    const unsigned long long lbo = KMAJOR ? 1ull : (4096ull >> 4);
    const unsigned long long sbo = KMAJOR ? (1024ull >> 4) : (512ull >> 4);
    const unsigned long long layout = KMAJOR ? 2ull : 1ull;
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}
// advance of the start-address field per K=8 step: 32 B inside the swizzle span (K-major) or two
// 4-k atoms = 1024 B (MN-major)
template <bool KMAJOR>
__device__ __forceinline__ unsigned kstep_units() { return KMAJOR ? 2u : 64u; }

// Instruction descriptor (kind::tf32): D format f32 (bits 4-5 = 1), A/B format TF32 (bits 7-9 /
// 10-12 = 2), A/B major (bits 15 / 16: 0 = K, 1 = MN), N >> 3 (bits 17-22), M >> 4 (bits 24-28)
template <bool A_KMAJOR, bool B_KMAJOR>
__device__ __forceinline__ unsigned make_idesc() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((A_KMAJOR ? 0u : 1u) << 15) | ((B_KMAJOR ? 0u : 1u) << 16) |
           ((unsigned)(BN >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
}

// tile index -> (tile row, tile column): bands of RASTER_N tile columns walked down the rows, so the ~148
// tiles in flight cover about 9 x 16 tiles and share 25 operand panels instead of the 66 of a plain
// column-major walk (ncu of the first version: 49 GB of DRAM reads for 2.4 GB of operands).  Used by the
// producer and the epilogue alike; exported for the CPU bijection test (elb200_tf32_tile_coords).
constexpr int RASTER_N = 16;
__host__ __device__ __forceinline__ void tf32_tile_coords(i64 t, i64 tilesM, i64 tilesN, i64& tm, i64& tn) {
    const i64 band_sz = (i64)RASTER_N * tilesM;
    const i64 band = t / band_sz;
    const i64 first_n = band * RASTER_N;
    const i64 bw = (tilesN - first_n < RASTER_N) ? (tilesN - first_n) : (i64)RASTER_N;
    const i64 in_band = t - band * band_sz;
    tm = in_band / bw;
    tn = first_n + (in_band - tm * bw);
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(NUM_THREADS, 1) sgemm_3xtf32_kernel(const __grid_constant__ TcArgs p) {
    extern __shared__ unsigned char smem_raw[];
    const unsigned raw = smem_u32(smem_raw);
    const unsigned base = (raw + 1023u) & ~1023u;  // the 128-byte swizzle pattern repeats every 1024 B
    const unsigned bars = base + STAGES * STAGE_BYTES;
    const unsigned full0 = bars, split0 = bars + 8 * STAGES, empty0 = bars + 16 * STAGES;
    const unsigned tfull0 = bars + 24 * STAGES, tempty0 = tfull0 + 16, tmem_slot = tempty0 + 16;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(split0 + 8 * s, 32 * SPLIT_WARPS);
            mbar_init(empty0 + 8 * s, 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull0 + 8 * a, 1);
            mbar_init(tempty0 + 8 * a, 32 * EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    if (warp == 1) {
        // TMEM allocation is warp-collective; the base address lands in shared memory
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    unsigned tmem_base;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    const i64 KB = (p.k + BK - 1) / BK;
    const i64 total = p.tilesM * p.tilesN;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            unsigned ph = 0;
            for (i64 t = blockIdx.x; t < total; t += gridDim.x) {
                i64 tm, tn;
                tf32_tile_coords(t, p.tilesM, p.tilesN, tm, tn);
                const int m0 = (int)(tm * BM), n0 = (int)(tn * BN);
                for (i64 kb = 0; kb < KB; ++kb) {
                    mbar_wait(empty0 + 8 * stage, ph ^ 1u);
                    const unsigned sa = base + stage * STAGE_BYTES, sb = sa + OP_BYTES;
                    const unsigned fb = full0 + 8 * stage;
                    mbar_expect_tx(fb, 2 * OP_BYTES);
                    const int k0 = (int)(kb * BK);
                    if (A_KMAJOR) {
                        tma_load_2d(sa, &p.mapA, k0, m0, fb);
                    } else {
#pragma unroll
                        for (int b = 0; b < BM / 32; ++b) tma_load_2d(sa + b * 4096, &p.mapA, m0 + 32 * b, k0, fb);
                    }
                    if (B_KMAJOR) {
                        tma_load_2d(sb, &p.mapB, k0, n0, fb);
                    } else {
#pragma unroll
                        for (int b = 0; b < BN / 32; ++b) tma_load_2d(sb + b * 4096, &p.mapB, n0 + 32 * b, k0, fb);
                    }
                    if (++stage == STAGES) { stage = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            const unsigned idesc = make_idesc<A_KMAJOR, B_KMAJOR>();
            int stage = 0;
            unsigned ph = 0;
            unsigned ci = 0;  // chunk counter: accumulator ci & 1
            for (i64 t = blockIdx.x; t < total; t += gridDim.x) {
                for (i64 kb0 = 0; kb0 < KB; kb0 += KCHUNK, ++ci) {
                    const unsigned acc = ci & 1u, aph = (ci >> 1) & 1u;
                    mbar_wait(tempty0 + 8 * acc, aph ^ 1u);  // the epilogue has drained this accumulator
                    tc_fence_after();
                    const unsigned d = tmem_base + acc * BN;
                    const i64 kb1 = kb0 + KCHUNK < KB ? kb0 + KCHUNK : KB;
                    for (i64 kb = kb0; kb < kb1; ++kb) {
                        mbar_wait(split0 + 8 * stage, ph);
                        tc_fence_after();
                        const unsigned sa = base + stage * STAGE_BYTES;
                        const unsigned long long ahi = make_desc<A_KMAJOR>(sa), bhi = make_desc<B_KMAJOR>(sa + OP_BYTES);
                        const unsigned long long alo = make_desc<A_KMAJOR>(sa + 2 * OP_BYTES),
                                                 blo = make_desc<B_KMAJOR>(sa + 3 * OP_BYTES);
#pragma unroll
                        for (int ks = 0; ks < BK / 8; ++ks) {
                            const unsigned long long ka = (unsigned long long)(ks * kstep_units<A_KMAJOR>());
                            const unsigned long long kbb = (unsigned long long)(ks * kstep_units<B_KMAJOR>());
                            umma_tf32(d, alo + ka, bhi + kbb, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
                            umma_tf32(d, ahi + ka, blo + kbb, idesc, 1u);
                            umma_tf32(d, ahi + ka, bhi + kbb, idesc, 1u);
                        }
                        umma_commit(empty0 + 8 * stage);  // ring stage free once these MMAs have read it
                        if (++stage == STAGES) { stage = 0; ph ^= 1u; }
                    }
                    umma_commit(tfull0 + 8 * acc);  // chunk accumulator complete
                }
            }
        }
    } else if (warp < 2 + SPLIT_WARPS) {
        // ===== splitters: fp32 -> (hi, lo) TF32 pairs, elementwise at identical offsets =====
        const int st = tid - 64;
        int stage = 0;
        unsigned ph = 0;
        for (i64 t = blockIdx.x; t < total; t += gridDim.x) {
            for (i64 kb = 0; kb < KB; ++kb) {
                mbar_wait(full0 + 8 * stage, ph);
                const unsigned hi_base = base + stage * STAGE_BYTES + (unsigned)st * 16u;
                const unsigned lo_base = hi_base + 2 * OP_BYTES;
#pragma unroll 4
                for (int i = 0; i < (2 * OP_BYTES) / (16 * 32 * SPLIT_WARPS); ++i) {
                    const unsigned off = (unsigned)i * (16u * 32u * SPLIT_WARPS);
                    float x0, x1, x2, x3;
                    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                 : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3)
                                 : "r"(hi_base + off));
                    const unsigned h0 = rn_tf32(x0), h1 = rn_tf32(x1), h2 = rn_tf32(x2), h3 = rn_tf32(x3);
                    const unsigned l0 = __float_as_uint(x0 - __uint_as_float(h0)), l1 = __float_as_uint(x1 - __uint_as_float(h1));
                    const unsigned l2 = __float_as_uint(x2 - __uint_as_float(h2)), l3 = __float_as_uint(x3 - __uint_as_float(h3));
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(hi_base + off), "r"(h0), "r"(h1), "r"(h2),
                                 "r"(h3)
                                 : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(lo_base + off), "r"(l0), "r"(l1), "r"(l2),
                                 "r"(l3)
                                 : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic writes -> tensor-core reads
                mbar_arrive(split0 + 8 * stage);
                if (++stage == STAGES) { stage = 0; ph ^= 1u; }
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers (promotion of every k-chunk) -> C =====
        const int q = warp & 3;                    // a warp may only touch TMEM lanes [32 (warp % 4), +32)
        const int half = (warp - 2 - SPLIT_WARPS) >> 2;  // columns [64 half, +64) of the tile
        const float alpha = p.alpha, beta = p.beta;
        unsigned ci = 0;
        for (i64 t = blockIdx.x; t < total; t += gridDim.x) {
            i64 tm, tn;
            tf32_tile_coords(t, p.tilesM, p.tilesN, tm, tn);
            const i64 m0 = tm * BM, n0 = tn * BN + half * 64;
            float sum[64];
            for (i64 kb0 = 0; kb0 < KB; kb0 += KCHUNK, ++ci) {
                const unsigned acc = ci & 1u, aph = (ci >> 1) & 1u;
                mbar_wait(tfull0 + 8 * acc, aph);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    unsigned v[16];
                    const unsigned taddr = tmem_base + ((unsigned)(q * 32) << 16) + acc * BN + (unsigned)(half * 64 + c * 16);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                          "=r"(v[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (kb0 == 0) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[c * 16 + j] = __uint_as_float(v[j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) sum[c * 16 + j] += __uint_as_float(v[j]);
                    }
                }
                // this chunk is in registers: hand the accumulator back to the MMA warp
                tc_fence_before();
                mbar_arrive(tempty0 + 8 * acc);
            }
            const i64 row = m0 + q * 32 + lane;
            if (row < p.m) {
                float* cbase = p.C + row + n0 * p.ldc;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float old[16];
                    if (beta != 0.f) {
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            old[j] = (n0 + c * 16 + j < p.n) ? __ldcs(cbase + (i64)(c * 16 + j) * p.ldc) : 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        if (n0 + c * 16 + j < p.n) {
                            float r = alpha * sum[c * 16 + j];
                            if (beta != 0.f) r = fmaf(beta, old[j], r);
                            cbase[(i64)(c * 16 + j) * p.ldc] = r;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
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

// 2-D f32 tensor map: inner (contiguous) extent `inner`, outer extent `outer`, outer stride ld
void make_map(CUtensorMap* map, const float* ptr, i64 inner, i64 outer, i64 ld, int boxInner, int boxOuter,
              CUtensorMapSwizzle swizzle) {
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4u};
    cuuint32_t box[2] = {(cuuint32_t)boxInner, (cuuint32_t)boxOuter};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) throw std::runtime_error("cuTensorMapEncodeTiled (f32) failed (code " + std::to_string((int)r) + ")");
}

template <bool AK, bool BKM>
void launch(const TcArgs& a, double flops, cudaStream_t s) {
    static bool configured = false;
    auto kern = sgemm_3xtf32_kernel<AK, BKM>;
    if (!configured) {
        ELB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    i64 grid = a.tilesM * a.tilesN;
    if (grid > sm_count()) grid = sm_count();
    if (sm_limit() > 0 && grid > sm_limit()) grid = sm_limit();
    gemm_profile_begin(s);
    kern<<<(unsigned)grid, NUM_THREADS, SMEM_BYTES, s>>>(a);
    ELB_LAUNCH_CHECK();
    gemm_profile_end(s, flops);
}

int g_sgemm_mode = 0;  // 0 exact FFMA (gemm_simt.cu), 1 3xTF32 on tcgen05
int g_sgemm_last = 0;  // which kernel served the last float GEMM: 1 SIMT, 2 tcgen05

}  // namespace

bool sgemm_3xtf32_eligible(i64 m, i64 n, i64 k, const float* A, i64 lda, const float* B, i64 ldb) {
    if (m <= 0 || n <= 0 || k <= 0) return false;
    if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || (lda & 3) || (ldb & 3)) return false;
    if (m >= (i64(1) << 31) - 256 || n >= (i64(1) << 31) - 256 || k >= (i64(1) << 31) - 256) return false;
    return encode_fn() != nullptr;
}

// C := alpha op(A) op(B) + beta C in 3xTF32 arithmetic; returns false (nothing launched) when the
// operands do not meet TMA's alignment rules
bool sgemm_3xtf32_device(char ta_, char tb_, i64 m, i64 n, i64 k, float alpha, const float* A, i64 lda,
                         const float* B, i64 ldb, float beta, float* C, i64 ldc, cudaStream_t s) {
    const bool ta = up(ta_) != 'N', tb = up(tb_) != 'N';
    if (!sgemm_3xtf32_eligible(m, n, k, A, lda, B, ldb)) return false;
    TcArgs a;
    // A 'T'/'C' is stored k x m (k contiguous: K-major); A 'N' is stored m x k (MN-major)
    const bool ak = ta, bk = !tb;
    const CUtensorMapSwizzle swK = CU_TENSOR_MAP_SWIZZLE_128B, swMN = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    if (ak) make_map(&a.mapA, A, k, m, lda, BK, BM, swK);
    else make_map(&a.mapA, A, m, k, lda, 32, BK, swMN);
    if (bk) make_map(&a.mapB, B, k, n, ldb, BK, BN, swK);
    else make_map(&a.mapB, B, n, k, ldb, 32, BK, swMN);
    a.m = m; a.n = n; a.k = k;
    a.C = C; a.ldc = ldc;
    a.alpha = alpha; a.beta = beta;
    a.tilesM = ceil_div(m, BM);
    a.tilesN = ceil_div(n, BN);
    const double flops = 2.0 * double(m) * double(n) * double(k);
    if (ak) { if (bk) launch<true, true>(a, flops, s); else launch<true, false>(a, flops, s); }
    else { if (bk) launch<false, true>(a, flops, s); else launch<false, false>(a, flops, s); }
    g_sgemm_last = 2;
    return true;
}

int sgemm_mode() { return g_sgemm_mode; }
void sgemm_note_simt() { g_sgemm_last = 1; }

}  // namespace elb200

extern "C" {
using namespace elb200;

int elb200_sgemm_3xtf32(char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha, const float* A, int64_t lda,
                        const float* B, int64_t ldb, float beta, float* C, int64_t ldc, elb200_stream_t s) {
    return guarded([&] {
        const char a = up(ta), b = up(tb);
        if ((a != 'N' && a != 'T' && a != 'C') || (b != 'N' && b != 'T' && b != 'C'))
            throw std::logic_error("sgemm_3xtf32: invalid trans");
        if (m < 0 || n < 0 || k < 0) throw std::logic_error("sgemm_3xtf32: negative dimension");
        if (m == 0 || n == 0) return;
        if (k == 0) {  // C := beta C
            if (beta == 0.f)
                ELB_CUDA(cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)m, (size_t)n,
                                           (cudaStream_t)s));
            else if (beta != 1.f)
                lattice_copy_device<float>(C, C, m, n, 0, 1, ldc, 0, 1, ldc, false, &beta, false, (cudaStream_t)s);
            return;
        }
        if (!sgemm_3xtf32_device(a, b, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, (cudaStream_t)s))
            throw std::runtime_error(
                "sgemm_3xtf32: operands must be 16-byte aligned with leading dimensions that are multiples of 4 "
                "(TMA); use elb200_sgemm for arbitrary layouts");
    });
}
void elb200_tf32_tile_coords(int64_t t, int64_t tilesM, int64_t tilesN, int64_t* tm, int64_t* tn) {
    i64 a = 0, b = 0;
    elb200::tf32_tile_coords(t, tilesM, tilesN, a, b);
    *tm = a; *tn = b;
}
void elb200_sgemm_set_mode(int mode) { elb200::g_sgemm_mode = mode == 1 ? 1 : 0; }
int elb200_sgemm_get_mode(void) { return elb200::g_sgemm_mode; }
int elb200_sgemm_last_kernel(void) { return elb200::g_sgemm_last; }
}
