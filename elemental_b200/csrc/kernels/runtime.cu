// Layer runtime: last-error text, current stream, device check and the FP64
// tensor-pipe (DMMA) ceiling probe that bench.py uses as roofline denominator.
#include "../common.hpp"
#include "elb200_blas.h"
#include <cstdlib>
#include <vector>

namespace elb200 {

static thread_local std::string g_last_error;
static thread_local cudaStream_t g_stream = nullptr;

void set_last_error(const std::string& s) { g_last_error = s; }
const char* last_error() { return g_last_error.c_str(); }
cudaStream_t current_stream() { return g_stream; }
void set_current_stream(cudaStream_t s) { g_stream = s; }

unsigned long long g_kernel_launches = 0;

// ELB200_BLAS_SYNC=1: every Fortran-ABI entry point (dgemm_, dtrsm_, dscal_, ...) synchronises its stream before
// returning, i.e. behaves like a host BLAS.  For callers that read the arrays from the host right after the call
// without a fence of their own -- the unmodified reference linked against this library.
void fortran_abi_fence() {
    static int mode = -1;
    if (mode < 0) {
        const char* e = std::getenv("ELB200_BLAS_SYNC");
        mode = (e && std::atoi(e) != 0) ? 1 : 0;
    }
    if (mode == 1) cudaStreamSynchronize(g_stream);
}

static int g_sm_limit = 0;
int sm_limit() { return g_sm_limit; }
void set_sm_limit(int n) { g_sm_limit = n < 0 ? 0 : n; }

cudaStream_t aux_stream(int idx) {
    static cudaStream_t streams[4] = {nullptr, nullptr, nullptr, nullptr};
    if (idx < 0 || idx >= 4) throw std::logic_error("aux_stream: index out of range");
    if (!streams[idx]) {
        int lo = 0, hi = 0;
        ELB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = numerically smallest = highest priority
        ELB_CUDA(cudaStreamCreateWithPriority(&streams[idx], cudaStreamNonBlocking, hi));
    }
    return streams[idx];
}

// ---- GEMM launch profiling (off by default) ----
namespace {
struct ProfPair { cudaEvent_t a, b; double flops; };
bool g_prof_on = false;
std::vector<ProfPair> g_prof;
size_t g_prof_used = 0;
}  // namespace
void gemm_profile_begin(cudaStream_t s) {
    if (!g_prof_on) return;
    if (g_prof_used == g_prof.size()) {
        ProfPair p;
        ELB_CUDA(cudaEventCreate(&p.a));
        ELB_CUDA(cudaEventCreate(&p.b));
        p.flops = 0;
        g_prof.push_back(p);
    }
    ELB_CUDA(cudaEventRecord(g_prof[g_prof_used].a, s));
}
void gemm_profile_end(cudaStream_t s, double flops) {
    if (!g_prof_on) return;
    g_prof[g_prof_used].flops = flops;
    ELB_CUDA(cudaEventRecord(g_prof[g_prof_used].b, s));
    ++g_prof_used;
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        ELB_CUDA(cudaGetDevice(&dev));
        ELB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    }
    return n;
}

namespace {
// 8 warps per CTA, each with 16 independent m8n8k4 accumulator tiles.
__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* sink) {
    double c[16][2];
#pragma unroll
    for (int i = 0; i < 16; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            asm volatile(
                "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                : "+d"(c[i][0]), "+d"(c[i][1])
                : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) sink[0] = s;  // keep the loop alive
}
}  // namespace
}  // namespace elb200

extern "C" {

const char* elb200_last_error(void) { return elb200::last_error(); }
int elb200_version(void) { return 100; }

int elb200_device_check(void) {
    return elb200::guarded([] {
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0)
            throw std::runtime_error(
                "elb200: no CUDA device visible -- this layer has no CPU fallback");
        int dev = 0, major = 0;
        ELB_CUDA(cudaGetDevice(&dev));
        ELB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        if (major != 10)
            throw std::runtime_error("elb200: kernels are built for sm_100a only");
    });
}

unsigned long long elb200_launch_count(int reset) {
    unsigned long long v = elb200::g_kernel_launches;
    if (reset) elb200::g_kernel_launches = 0;
    return v;
}
void elb200_gemm_profile(int enable) {
    elb200::g_prof_on = enable != 0;
    elb200::g_prof_used = 0;
}
int elb200_gemm_profile_read(double* total_ms, long long* launches, double* flops) {
    return elb200::guarded([&] {
        double ms = 0, fl = 0;
        for (size_t i = 0; i < elb200::g_prof_used; ++i) {
            ELB_CUDA(cudaEventSynchronize(elb200::g_prof[i].b));
            float t = 0;
            ELB_CUDA(cudaEventElapsedTime(&t, elb200::g_prof[i].a, elb200::g_prof[i].b));
            ms += t;
            fl += elb200::g_prof[i].flops;
        }
        if (total_ms) *total_ms = ms;
        if (launches) *launches = (long long)elb200::g_prof_used;
        if (flops) *flops = fl;
        elb200::g_prof_used = 0;
    });
}

void elb200_set_sm_limit(int n) { elb200::set_sm_limit(n); }
int elb200_get_sm_limit(void) { return elb200::sm_limit(); }
void elb200_set_stream(elb200_stream_t s) { elb200::set_current_stream((cudaStream_t)s); }
elb200_stream_t elb200_get_stream(void) { return (elb200_stream_t)elb200::current_stream(); }

int elb200_dmma_peak(int iters, double* flops_per_s, float* ms_out) {
    return elb200::guarded([&] {
        using namespace elb200;
        const int ctas = sm_count() * 2;
        double* sink = nullptr;
        ELB_CUDA(cudaMalloc(&sink, 8));
        cudaEvent_t e0, e1;
        ELB_CUDA(cudaEventCreate(&e0));
        ELB_CUDA(cudaEventCreate(&e1));
        dmma_peak_kernel<<<ctas, 256>>>(iters / 10 + 1, sink);  // warm-up
        ELB_CUDA(cudaEventRecord(e0));
        dmma_peak_kernel<<<ctas, 256>>>(iters, sink);
        ELB_CUDA(cudaEventRecord(e1));
        ELB_CUDA(cudaEventSynchronize(e1));
        ELB_LAUNCH_CHECK();
        float ms = 0.f;
        ELB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        // per warp per iteration: 16 MMAs x (8*8*4) FMAs x 2 flops
        const double flops = double(ctas) * 8.0 * double(iters) * 16.0 * 512.0;
        if (flops_per_s) *flops_per_s = flops / (ms * 1e-3);
        if (ms_out) *ms_out = ms;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        cudaFree(sink);
    });
}

}  // extern "C"
