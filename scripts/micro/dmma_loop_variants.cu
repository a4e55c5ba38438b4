// What does each ingredient of the DGEMM main loop cost the FP64 tensor pipe?  One persistent CTA of 16 warps per
// SM (128 registers per thread, as the GEMM kernels), every warp owns a 4 x 4 grid of m8n8k4 accumulators
// (64 registers) and runs `stages` x 64 DMMAs.  Variants add the ingredients one by one:
//   0  DMMAs only (the ceiling)
//   1  + __syncwarp() every 64 DMMAs
//   2  + mbarrier.try_wait on an already completed phase every 64 DMMAs
//   3  + lane-0 mbarrier.arrive every 64 DMMAs on a barrier nobody waits for
//   4  + 8 LDS.64 fragment loads per 16 DMMAs (conflict-free pattern, single-buffered fragments)
//   5  like 4 with the fragments of k-step i+1 loaded before the DMMAs of k-step i (double-buffered)
//   6  4 + accumulator reset and an integer keep-alive every 8 stages (a k = 128 tile)
//   7  4 + a REAL full/empty handshake: warp 0 of each 8-warp group arrives on `full` after testing `empty`
//   8  7 + 64 red.global.add.f64 per thread every 8 stages (the rank-128 epilogue)
//   9  8 + REAL TMA loads: the producer issues the A (3-D box) and B (2-D box) loads of the tile's k-stage from the
//      32768 x 128 / 128 x 32768 panels with expect_tx instead of the plain arrive
//  10  9 without the epilogue (TMA + handshake + LDS + DMMA)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/dmma_loop_variants.bin scripts/micro/dmma_loop_variants.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra LD;\nbra LW;\nLD:\n}\n"
                 ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void red_add_f64(double* addr, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

constexpr int STAGES = 4;
struct Maps { CUtensorMap a, b; };
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

template <int V>
__global__ void __launch_bounds__(512, 1) loop_kernel(int stages, double* sink, double* cbuf, long long ldc, const __grid_constant__ Maps maps) {
    extern __shared__ unsigned char smem_raw[];
    const unsigned base = ((unsigned)__cvta_generic_to_shared(smem_raw) + 1023u) & ~1023u;
    const unsigned bars = base + 2 * STAGES * 24576;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = warp >> 3, cw = warp & 7;
    const unsigned full0 = bars + group * 2 * STAGES * 8, empty0 = full0 + STAGES * 8;
    if (tid == 0) {
        for (int g = 0; g < 2; ++g)
            for (int s = 0; s < STAGES; ++s) {
                mbar_init(bars + (g * 2 * STAGES + s) * 8, 1);
                mbar_init(bars + (g * 2 * STAGES + STAGES + s) * 8, 8);
            }
        mbar_init(bars + 256, 1);   // "done" barrier: completed once below
        mbar_init(bars + 264, 1);   // nobody waits for this one
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // shared-memory operands: any finite values
    for (int i = tid; i < 2 * STAGES * 24576 / 8; i += 512) ((double*)smem_raw)[i + ((base - (unsigned)__cvta_generic_to_shared(smem_raw)) >> 3)] = 1.0 + 1e-9 * i;
    __syncthreads();
    if (tid == 0) mbar_arrive(bars + 256);
    __syncthreads();

    const int g = lane >> 2, t = lane & 3;
    const int wm0 = (cw & 3) * 32, wn0 = (cw >> 2) * 32;
    // the MN-major A / K-major B fragment addressing of the GEMM kernel (NN)
    unsigned aoff[4], boff[4];
    {
        const int q = ((g & 2) << 2) | (g & 1) | ((g & 4) >> 1);
        const unsigned lo = (unsigned)(wm0 / 16) * 2048u + (unsigned)t * 128u + (unsigned)(q & 1) * 8u;
        const int L = (q >> 1) ^ t;
        for (int x = 0; x < 4; ++x) aoff[x] = lo + (unsigned)((L ^ (2 * x)) << 4);
        const int pr = ((g & 3) << 1) | (g >> 2);
        const unsigned rowoff = (unsigned)(wn0 + pr) * 128u + (unsigned)(t & 1) * 8u;
        const int L2 = pr ^ (t >> 1);
        for (int ks = 0; ks < 4; ++ks) boff[ks] = rowoff + (unsigned)((L2 ^ (2 * ks)) << 4);
    }
    double acc[4][4][2];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double a[4], b[4];
    for (int i = 0; i < 4; ++i) { a[i] = 1.0 + lane * 1e-9 + i; b[i] = 1.0 - lane * 1e-9 - i; }

    int stage = 0; unsigned phase = 0;
    int pstage = 0; unsigned pphase = 0; int ahead = 0, issued = 0;
    auto produce = [&](bool blocking) {
        if (issued >= stages) return;
        if (blocking) mbar_wait(empty0 + pstage * 8, pphase ^ 1u);
        else if (!mbar_test(empty0 + pstage * 8, pphase ^ 1u)) return;
        if (V >= 9) {
            if (lane == 0) {
                const long long tile = (long long)(blockIdx.x * 2 + group) + (long long)(issued >> 3) * gridDim.x * 2;
                const int m0 = (int)(tile % 256) * 128, n0 = (int)((tile / 256) % 512) * 64, k0 = (issued & 7) * 16;
                const unsigned sa = base + group * STAGES * 24576 + pstage * 24576;
                mbar_expect_tx(full0 + pstage * 8, 24576);
                tma_load_3d(sa, &maps.a, 0, k0, m0 >> 4, full0 + pstage * 8);
                tma_load_2d(sa + 16384, &maps.b, k0, n0, full0 + pstage * 8);
            }
        } else if (lane == 0) mbar_arrive(full0 + pstage * 8);
        __syncwarp();
        ++ahead; ++issued;
        if (++pstage == STAGES) { pstage = 0; pphase ^= 1u; }
    };
    if (V >= 7 && cw == 0) for (int i = 0; i < STAGES; ++i) produce(false);

#pragma unroll 1
    for (int s = 0; s < stages; ++s) {
        if (V == 2 || V == 3) mbar_wait(bars + 256, 0);
        if (V >= 7) {
            if (cw == 0 && ahead == 0) produce(true);
            mbar_wait(full0 + stage * 8, phase);
        }
        const unsigned sa = base + group * STAGES * 24576 + stage * 24576, sb = sa + 16384;
        if (V == 5) {
            double fa[2][4], fb[2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                fa[0][i] = lds64(sa + aoff[(i & 1)] + (unsigned)(i >> 1) * 2048u);
                fb[0][i] = lds64(sb + boff[0] + (unsigned)i * 1024u);
            }
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int cur = ks & 1;
                if (ks + 1 < 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        fa[cur ^ 1][i] = lds64(sa + aoff[(i & 1) + 2 * ((ks + 1) & 1)] + (unsigned)(i >> 1) * 2048u + (unsigned)(ks + 1) * 512u);
                        fb[cur ^ 1][i] = lds64(sb + boff[ks + 1] + (unsigned)i * 1024u);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], fa[cur][i], fb[cur][j]);
            }
        } else {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                if (V >= 7 && cw == 0) produce(false);
                if (V >= 4) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        a[i] = lds64(sa + aoff[(i & 1) + 2 * (ks & 1)] + (unsigned)(i >> 1) * 2048u + (unsigned)ks * 512u);
                        b[i] = lds64(sb + boff[ks] + (unsigned)i * 1024u);
                    }
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
            }
        }
        if (V >= 1) __syncwarp();
        if (V == 3 && lane == 0) mbar_arrive(bars + 264);
        if (V >= 7) {
            if (lane == 0) mbar_arrive(empty0 + stage * 8);
            --ahead;
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        if ((V == 6 || V >= 8) && (s & 7) == 7) {
            if (V == 8 || V == 9) {
                // the epilogue's access pattern: lane (g, t) owns rows {..}, columns 2t, 2t+1 of each 8 x 8 fragment
                const long long tile = (long long)(blockIdx.x * 2 + group) + (long long)(s >> 3) * gridDim.x * 2;
                double* cb = cbuf + (tile % 256) * 128 + wm0 + g + ((long long)wn0 + ((tile / 256) % 512) * 64) * ldc;
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int i = 0; i < 4; ++i) red_add_f64(cb + (8 * j + 2 * t + e) * ldc + 8 * i, acc[i][j][e]);
            } else {
                int x = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) x ^= __double2hiint(acc[i][j][0]) ^ __double2loint(acc[i][j][1]);
                if (x == 0x12345678) sink[1] = 1.0;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        }
    }
    double sum = 0.0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) sum += acc[i][j][0] + acc[i][j][1];
    if (sum == 123.456) sink[0] = sum;
}

Maps g_maps;
template <int V>
void run(int nsm, int stages, double* sink, double* cbuf, long long ldc) {
    const int smem = 2 * STAGES * 24576 + 1024 + 512;
    cudaFuncSetAttribute(loop_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    loop_kernel<V><<<nsm, 512, smem>>>(stages / 8, sink, cbuf, ldc, g_maps);
    float best = 1e9f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        loop_kernel<V><<<nsm, 512, smem>>>(stages, sink, cbuf, ldc, g_maps);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const double flops = double(nsm) * 16.0 * double(stages) * 64.0 * 512.0;
    printf("variant %d: %8.3f ms  %7.2f TFLOP/s  (%s)\n", V, best, flops / (best * 1e-3) / 1e12, cudaGetErrorString(err));
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    double *sink, *cbuf;
    const long long ldc = 32768;            // C of the benchmark: 32768 x 32768 doubles, 256 x 512 tiles of 128 x 64
    cudaMalloc(&sink, 64);
    cudaMalloc(&cbuf, ldc * 32768 * 8);
    cudaMemset(cbuf, 0, ldc * 32768 * 8);
    if (cudaGetLastError() != cudaSuccess) { printf("cbuf alloc failed\n"); return 1; }
    // operand panels of the rank-128 update and their tensor maps (as gemm_f64_ws.cu builds them for NN)
    double *A, *B;
    cudaMalloc(&A, 32768LL * 128 * 8); cudaMalloc(&B, 128LL * 32768 * 8);
    cudaMemset(A, 0, 32768LL * 128 * 8); cudaMemset(B, 0, 128LL * 32768 * 8);
    {
        typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* ptr = nullptr; cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
        Fn enc = (Fn)ptr;
        cuuint64_t d3[3] = {16, 128, 32768 / 16}; cuuint64_t s3[2] = {32768 * 8ull, 128}; cuuint32_t b3[3] = {16, 16, 8}; cuuint32_t e3[3] = {1, 1, 1};
        CUresult r1 = enc(&g_maps.a, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, A, d3, s3, b3, e3, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        cuuint64_t d2[2] = {128, 32768}; cuuint64_t s2[1] = {128 * 8ull}; cuuint32_t b2[2] = {16, 64}; cuuint32_t e2[2] = {1, 1};
        CUresult r2 = enc(&g_maps.b, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, B, d2, s2, b2, e2, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("tensor maps: %d %d\n", (int)r1, (int)r2);
    }
    const int stages = 8 * 443;  // the rank-128 update: 443 tiles of 8 stages per group
    if (only < 0 || only == 0) run<0>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 1) run<1>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 2) run<2>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 3) run<3>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 4) run<4>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 5) run<5>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 6) run<6>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 7) run<7>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 8) run<8>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 9) run<9>(nsm, stages, sink, cbuf, ldc);
    if (only < 0 || only == 10) run<10>(nsm, stages, sink, cbuf, ldc);
    return 0;
}
