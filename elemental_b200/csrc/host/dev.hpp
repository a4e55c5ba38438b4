// Host-layer internals: mapping from El scalar types to the kernels' POD types,
// NCCL error checking and the typed leaf wrappers the algorithms call.
#pragma once
#include <nccl.h>

#include <string>
#include <utility>
#include <vector>

#include "../kernels/device_api.hpp"
#include "elb200/core.hpp"
#include "elb200_level1.h"

namespace El {
namespace dev {

template <typename T> struct DevType;
template <> struct DevType<float> { typedef float type; };
template <> struct DevType<double> { typedef double type; };
template <> struct DevType<Complex<float>> { typedef elb200::c32_t type; };
template <> struct DevType<Complex<double>> { typedef elb200::c64_t type; };
template <typename T> using D = typename DevType<T>::type;

template <typename T> inline D<T> val(T x);
template <> inline float val<float>(float x) { return x; }
template <> inline double val<double>(double x) { return x; }
template <> inline elb200::c32_t val<Complex<float>>(Complex<float> x) { return elb200::mk(x.real(), x.imag()); }
template <> inline elb200::c64_t val<Complex<double>>(Complex<double> x) { return elb200::mk(x.real(), x.imag()); }

template <typename T> inline D<T>* ptr(T* p) { return reinterpret_cast<D<T>*>(p); }
template <typename T> inline const D<T>* ptr(const T* p) { return reinterpret_cast<const D<T>*>(p); }

template <typename T> constexpr int Code() { return elb200::dtype_code<D<T>>::value; }

inline cudaStream_t stream() { return (cudaStream_t)CurrentStream(); }

// ---- helpers of the overlapped panel loops ----
// RAII: make `s` the layer's current stream (everything the host layer enqueues goes there)
struct StreamScope {
    cudaStream_t prev;
    explicit StreamScope(cudaStream_t s) : prev(stream()) { elb200::set_current_stream(s); }
    ~StreamScope() { elb200::set_current_stream(prev); }
};
// RAII: cap the persistent GEMM kernels at n CTAs (0 = all SMs)
struct SmLimitScope {
    int prev;
    explicit SmLimitScope(int n) : prev(elb200::sm_limit()) { elb200::set_sm_limit(n); }
    ~SmLimitScope() { elb200::set_sm_limit(prev); }
};
struct Event {
    cudaEvent_t e = nullptr;
    Event() { ELB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); }
    ~Event() { if (e) cudaEventDestroy(e); }
    Event(const Event&) = delete;
    Event& operator=(const Event&) = delete;
    void Record(cudaStream_t s) { ELB_CUDA(cudaEventRecord(e, s)); }
    void Wait(cudaStream_t s) { ELB_CUDA(cudaStreamWaitEvent(s, e, 0)); }
};
// A device int, zeroed on construction on the current stream; Read() synchronises that stream.  The
// kernels that can fail numerically (potrf pivots, the zero-diagonal scan of Trsm) write it only when it
// still holds 0, so one flag serves a whole blocked sweep and the host raises the exception once.
struct DeviceFlag {
    int* dev_ = nullptr;
    DeviceFlag() {
        dev_ = (int*)elb200::scratch_alloc(sizeof(int), stream());
        ELB_CUDA(cudaMemsetAsync(dev_, 0, sizeof(int), stream()));
    }
    ~DeviceFlag() { if (dev_) cudaFreeAsync(dev_, stream()); }
    DeviceFlag(const DeviceFlag&) = delete;
    DeviceFlag& operator=(const DeviceFlag&) = delete;
    int Read() {
        int h = 0;
        ELB_CUDA(cudaMemcpyAsync(&h, dev_, sizeof(int), cudaMemcpyDeviceToHost, stream()));
        ELB_CUDA(cudaStreamSynchronize(stream()));
        return h;
    }
};
// ELB200_TRACE=1: the factorisation drivers synchronise around each phase and print the summed
// device time per phase at the end (a diagnostic; it serialises the look-ahead)
struct PhaseTimer {
    static bool Enabled();
    std::vector<std::pair<std::string, double>> acc;
    cudaEvent_t a = nullptr, b = nullptr;
    PhaseTimer();
    ~PhaseTimer();
    void Begin(cudaStream_t s);
    void End(cudaStream_t s, const char* name);
    void Report(const char* title);
};
// SMs the panel stream gets beside a trailing update; ELB200_PANEL_SMS overrides the default
int PanelSms(int dflt);
// 0 disables the overlapped loops (ELB200_OVERLAP=0): everything runs on one stream
bool OverlapEnabled();

inline void nccl_check(ncclResult_t r, const char* what, const char* file, int line) {
    if (r != ncclSuccess) {
        RuntimeError(std::string("NCCL error in ") + what + " at " + file + ":" + std::to_string(line) + ": " +
                     ncclGetErrorString(r));
    }
}
#define ELB_NCCL(x) ::El::dev::nccl_check((x), #x, __FILE__, __LINE__)

inline void c_check(int rc, const char* what) {
    if (rc != 0) RuntimeError(std::string(what) + ": " + elb200::last_error());
}

}  // namespace dev
}  // namespace El
