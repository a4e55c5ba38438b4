// Microbenchmark: throughput of cp.reduce.async.bulk (.add.f64, smem -> global, performed at L2)
// vs plain st.global of the same bytes.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_red bulk_red_f64.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
// each CTA owns `tiles` tiles of 64 columns x 128 rows (column = 1 KB contiguous), column pitch ld doubles
__global__ void __launch_bounds__(256, 1) red_kernel(double* C, long long ld, long long tilesM, long long tilesTotal, int mode) {
    extern __shared__ __align__(128) double stage[];  // 64 x 128 doubles = 64 KB
    const int tid = threadIdx.x;
    for (long long tile = blockIdx.x; tile < tilesTotal; tile += gridDim.x) {
        const long long tm = tile % tilesM, tn = tile / tilesM;
        double* base = C + tm * 128 + tn * 64 * ld;
        if (mode == 0) {
            // wait until the previous bulk reads of the staging buffer are done
            if (tid < 64) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
            for (int i = tid; i < 64 * 128; i += 256) stage[i] = 1.0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (tid < 64) {
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(base + tid * ld),
                             "r"(smem_u32(stage + tid * 128)), "r"(1024)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        } else if (mode == 1) {
            for (int i = tid; i < 64 * 128; i += 256) base[(i & 127) + (long long)(i >> 7) * ld] = 1.0;
        } else if (mode == 2) {
            for (int i = tid; i < 64 * 128; i += 256) {
                double* q = base + (i & 127) + (long long)(i >> 7) * ld;
                *q = *q + 1.0;
            }
        } else {
            // the DMMA epilogue's access pattern: warp w owns rows 64*(w&1).., cols 32*(w>>1)..; one RED per
            // (i, j, e): lanes g -> 8 consecutive rows, t -> columns 2t+e
            const int w = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
            if (w < 4) {
                double* wb = base + 64 * (w & 1) + g + (long long)(32 * (w >> 1)) * ld;
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e)
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            asm volatile("red.global.add.f64 [%0], %1;" ::"l"(wb + 8 * i + (long long)(8 * j + 2 * t + e) * ld), "d"(1.0) : "memory");
            }
        }
    }
    if (mode == 0 && tid < 64) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
int main() {
    const long long m = 32768, n = 32768, ld = m;
    double* C;
    cudaMalloc(&C, sizeof(double) * ld * n);
    cudaMemset(C, 0, sizeof(double) * ld * n);
    cudaFuncSetAttribute(red_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const long long tilesM = m / 128, tiles = tilesM * (n / 64);
    // per-SM rates: one CTA (mode 3 uses 4 warps) over 4096 tiles -- is the SM's RED.64 rate near 8 B/clk?
    for (int mode = 0; mode < 4; ++mode) {
        const long long t1 = 4096;
        red_kernel<<<1, 256, 65536>>>(C, ld, tilesM, t1, mode);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        red_kernel<<<1, 256, 65536>>>(C, ld, tilesM, t1, mode);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("ONE CTA mode %d: %.3f ms for %lld tiles of 64 KB: %.2f GB/s per SM = %.2f B/clk at 1.965 GHz\n", mode, ms, t1,
               65536.0 * t1 / ms / 1e6, 65536.0 * t1 / ms / 1e6 / 1.965);
    }
    for (int mode = 0; mode < 4; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            red_kernel<<<148, 256, 65536>>>(C, ld, tilesM, tiles, mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("mode %d (%s): %.3f ms, %.1f GB/s of C bytes, err=%s\n", mode, mode == 0 ? "bulk red.add.f64" : mode == 1 ? "st.global" : mode == 2 ? "ld+st" : "red.global.add.f64 per thread", ms,
                   8.0 * m * n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    double h[4];
    cudaMemcpy(h, C + 12345, 32, cudaMemcpyDeviceToHost);
    printf("sample values (expect 3 red + 1 store... ) %g %g\n", h[0], h[1]);
    return 0;
}
