// UMMA shared-memory descriptor probe (kind::tf32): which shared-memory word does the tensor core
// read for logical A(m, k), given a descriptor?  Shared memory is filled with tags (word index,
// 7 bits per run so that every tag is exact in tf32), B is a K-major selection matrix
// (B[n][k] = (n == k)), so D[m][n] = tag of the word read for A(m, k = n).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o scripts/micro/umma_probe.bin scripts/micro/umma_probe.cu
//   ./umma_probe.bin                -> sweeps a few (major, LBO, SBO) candidates
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int REGION_WORDS = 16384;           // 64 KB of tagged words for the A operand
constexpr int B_OFF = REGION_WORDS * 4;       // B tile (16 rows x 128 B, K-major SW128) after it
constexpr int SMEM = B_OFF + 128 * 128 + 1024 + 64;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(unsigned long long adesc_hi, unsigned a_start_off, unsigned idesc, int shift,
                                                float* out, int f16, int probeB) {
    extern __shared__ unsigned char raw[];
    const unsigned rawa = smem_u32(raw);
    const unsigned base = (rawa + 1023u) & ~1023u;
    unsigned char* al = raw + (base - rawa);
    float* words = reinterpret_cast<float*>(al);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (f16) {
        __half* hw = reinterpret_cast<__half*>(al);
        for (int i = tid; i < 2 * REGION_WORDS; i += 128) hw[i] = __float2half((float)((i >> shift) & 0x7F));
    } else {
        for (int i = tid; i < REGION_WORDS; i += 128) words[i] = (float)((i >> shift) & 0x7F);
    }
    // B: K-major, 128-byte swizzle, rows n = 0..15, k = 0..7 (32 B of each 128-byte row); B[n][k] = (n == k)
    float* bw = reinterpret_cast<float*>(al + B_OFF);
    for (int i = tid; i < 128 * 32; i += 128) bw[i] = 0.f;
    __syncthreads();
    if (!f16 && tid < 8) {
        const int n = tid, k = tid;
        const int chunk = (k >> 2) ^ (n & 7);
        bw[n * 32 + chunk * 4 + (k & 3)] = 1.f;
    }
    if (f16 && tid < 16) {
        const int n = tid, k = tid;
        const int chunk = (k >> 3) ^ (n & 7);
        reinterpret_cast<__half*>(bw)[n * 64 + chunk * 8 + (k & 7)] = __float2half(1.f);
    }
    const unsigned bar = base + B_OFF + 128 * 128, slot = bar + 8;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(slot) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned tmem;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(slot));
    if (tid == 0) {
        const unsigned long long adesc = adesc_hi | (unsigned long long)(((base + a_start_off) >> 4) & 0x3FFFu);
        const unsigned long long bdesc = (unsigned long long)(((base + B_OFF) >> 4) & 0x3FFFu) | (1ull << 16) |
                                         ((1024ull >> 4) << 32) | (1ull << 46) | (2ull << 61);
        const unsigned long long da = probeB ? bdesc : adesc, db = probeB ? adesc : bdesc;
        if (f16)
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(0u)
                : "memory");
        else
            asm volatile(
                "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                "l"(da), "l"(db), "r"(idesc), "r"(0u)
                : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra W_DONE;\nbra W_LOOP;\nW_DONE:\n}\n" ::"r"(bar),
        "r"(0u)
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned v[16];
    const unsigned taddr = tmem + ((unsigned)(warp * 32) << 16);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * 16 + j] = __uint_as_float(v[j]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

static void run(const char* name, int f16, int probeB, int mn_major, unsigned lbo, unsigned sbo, unsigned layout, unsigned start_off,
                unsigned extra_idesc = 0) {
    // probe A: M=128 (tagged), N=16 (selection).  probe B: the selection matrix is A (rows m<K), tagged operand is B (N=16)
    const unsigned fmt = f16 ? 0u : 2u;
    const unsigned idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((unsigned)(mn_major && !probeB) << 15) |
                           ((unsigned)(mn_major && probeB) << 16) | ((16u >> 3) << 17) | ((128u >> 4) << 24) | extra_idesc;
    const unsigned long long hi = ((unsigned long long)((lbo >> 4) & 0x3FFF) << 16) | ((unsigned long long)((sbo >> 4) & 0x3FFF) << 32) |
                                  (1ull << 46) | ((unsigned long long)layout << 61);
    float* d;
    CK(cudaMalloc(&d, 128 * 16 * 4));
    std::vector<float> h[3];
    for (int r = 0; r < 3; ++r) {
        h[r].resize(128 * 16);
        CK(cudaMemset(d, 0xFF, 128 * 16 * 4));
        probe<<<1, 128, SMEM>>>(hi, start_off, idesc, 7 * r, d, f16, probeB);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h[r].data(), d, 128 * 16 * 4, cudaMemcpyDeviceToHost));
    }
    const int K = f16 ? 16 : 8, es = f16 ? 2 : 4;
    printf("== %s: %s probe %s %s-major layout_type=%u LBO=%u SBO=%u start+%u idesc=%08x (byte offsets; raw h0 of first entry %g)\n", name,
           f16 ? "f16" : "tf32", probeB ? "B" : "A", mn_major ? "MN" : "K", layout, lbo, sbo, start_off, idesc,
           h[0][0]);
    if (!probeB) {
        const int ms[] = {0, 1, 2, 7, 8, 9, 31, 32, 33, 64, 127};
        for (int m : ms) {
            printf("  m=%3d:", m);
            for (int k = 0; k < K; ++k)
                printf(" %6d", es * ((int)h[0][m * 16 + k] + 128 * (int)h[1][m * 16 + k] + 16384 * (int)h[2][m * 16 + k]));
            printf("\n");
        }
    } else {
        for (int k = 0; k < K; ++k) {   // D[m = k][n] = B(k, n)
            printf("  k=%3d:", k);
            for (int n = 0; n < 16; ++n)
                printf(" %6d", es * ((int)h[0][k * 16 + n] + 128 * (int)h[1][k * 16 + n] + 16384 * (int)h[2][k * 16 + n]));
            printf("\n");
        }
    }
    CK(cudaFree(d));
}

int main() {
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    run("sanity", 0, 0, 0, 16, 1024, 2, 0);
    run("tf32 A MN-major", 0, 0, 1, 4096, 1024, 2, 0);
    run("tf32 A MN-major 128B_BASE32B", 0, 0, 1, 4096, 512, 1, 0);
    run("tf32 A MN-major 128B_BASE32B +1024", 0, 0, 1, 4096, 512, 1, 1024);
    run("tf32 B MN-major 128B_BASE32B", 0, 1, 1, 4096, 512, 1, 0);
    run("f16 A MN-major", 1, 0, 1, 4096, 1024, 2, 0);
    return 0;
}
