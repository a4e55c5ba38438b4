// Shared host-side plumbing for the elb200 layer: error capture, the layer's
// current CUDA stream, and small helpers used by every launcher.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <stdexcept>

namespace elb200 {

typedef long long i64;

// Thread-local last error text; the C-ABI returns a nonzero code and the
// caller fetches the text with elb200_last_error().
void set_last_error(const std::string& s);
const char* last_error();

// The stream every Fortran-ABI leaf (dgemm_, dtrsm_, ...) launches on.
cudaStream_t current_stream();
void set_current_stream(cudaStream_t s);
void fortran_abi_fence();  // ELB200_BLAS_SYNC=1: the Fortran-ABI leaves synchronise before returning

struct CudaError : std::runtime_error {
    explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        char buf[512];
        snprintf(buf, sizeof buf, "CUDA error %s at %s:%d: %s", what, file, line,
                 cudaGetErrorString(e));
        throw CudaError(buf);
    }
}
#define ELB_CUDA(x) ::elb200::cuda_check((x), #x, __FILE__, __LINE__)
// every kernel launch of the layer goes through this macro: it also feeds the launch counter
// that bench.py reports as "gpu_launches"
extern unsigned long long g_kernel_launches;
#define ELB_LAUNCH_CHECK()                                                                  \
    do {                                                                                    \
        ++::elb200::g_kernel_launches;                                                      \
        ::elb200::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__);      \
    } while (0)

// optional per-launch timing of the GEMM kernels (CUDA events on the launching stream)
void gemm_profile_begin(cudaStream_t s);
void gemm_profile_end(cudaStream_t s, double flops);

inline i64 ceil_div(i64 a, i64 b) { return (a + b - 1) / b; }

// Wraps a C-ABI body: exceptions become error codes + last_error text.
template <class F>
int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        set_last_error(e.what());
        return 1;
    } catch (...) {
        set_last_error("unknown exception");
        return 1;
    }
}

inline char up(char c) { return (c >= 'a' && c <= 'z') ? char(c - 32) : c; }

int sm_count();

// Concurrency plumbing for the overlapped panel loops (SUMMA-C prefetch, Cholesky look-ahead):
//   * aux_stream(i): lazily created non-blocking streams of the highest priority; the
//     redistributions / panel factorisations that run beside a trailing update are enqueued there;
//   * sm_limit: the persistent GEMM kernels launch at most this many CTAs (0 = one per SM).  A
//     persistent CTA owns its SM's whole register file, so a concurrent NCCL / pack / potrf kernel
//     can only run beside it on SMs the GEMM leaves free.
cudaStream_t aux_stream(int idx);
int sm_limit();
void set_sm_limit(int n);

}  // namespace elb200
