// Flag kernels of the NVLink peer-memory redistribution path (host side: host/redist.cpp,
// host/core.cpp Grid::P2P).
//
// Every rank owns an exchange window in device memory that all its peers have mapped (CUDA IPC).
// A redistribution PUSHES its outgoing pieces straight into the destination's window (lattice
// kernel or copy engine writing over NVLink), so the "wire" step of the reference's
// pack -> MPI -> unpack (include/El/blas_like/level1/Copy/*.hpp) needs no communication kernel at
// all; what remains are two one-warp kernels per redistribution that move 4-byte epoch numbers:
//   p2p_exchange : tell every destination "my pieces of epoch e have landed" (ready flag in ITS
//                  window), then wait until every source has said so in MINE;
//   p2p_ack      : tell every peer "I have consumed epoch e" and wait until all of them consumed
//                  epoch e - 1, which frees the window half the next epoch writes into.
// Flags only ever grow, waits are `>=`, and every rank runs both kernels for every epoch of a
// channel (one channel per stream that issues redistributions), so the protocol cannot deadlock as
// long as all ranks issue the same sequence of redistributions -- the SPMD contract of the
// reference's collectives.  A wait that is not satisfied within ~10 s raises a sticky error flag in
// pinned host memory and returns instead of hanging the device.
#include <cuda.h>

#include <cstdlib>
#include <cstring>
#include <mutex>

#include "../common.hpp"
#include "device_api.hpp"

namespace elb200 {
namespace {

constexpr long long SPIN_LIMIT_CLOCKS = 20000000000LL;  // ~10 s at 2 GHz

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(64) p2p_flags_kernel(P2PFlagOps ops) {
    const int t = threadIdx.x;
    if (t < ops.nsignal) {
        __threadfence_system();
        st_release_sys(ops.signal[t], ops.epoch);
    }
    if (t < ops.nwait) {
        const long long t0 = clock64();
        unsigned ns = 32;
        while ((int)(ld_acquire_sys(ops.wait[t]) - ops.wait_value) < 0) {
            __nanosleep(ns);
            if (ns < 1024) ns <<= 1;
            if (clock64() - t0 > SPIN_LIMIT_CLOCKS) {
                *ops.error = 1;  // pinned host memory, sticky
                break;
            }
        }
        __threadfence_system();
    }
}

}  // namespace

namespace {
// Stream memory operations (cuStreamBatchMemOp: write-value / wait-value executed by the stream's front end): the
// epoch numbers move without a kernel, i.e. without an SM -- which matters beside a persistent GEMM that owns every
// SM -- and without a launch on the dependency chain of the panel stream.  Same protocol, same `>=` waits (cyclic
// compare).  No time-out in this mode: a rank that leaves the SPMD sequence hangs its peers as it would inside an
// NCCL collective.  ELB200_P2P_MEMOPS=0 (or a driver that rejects the call) selects the flag kernel above.
typedef CUresult (*BatchMemOpFn)(CUstream, unsigned int, CUstreamBatchMemOpParams*, unsigned int);
BatchMemOpFn g_batch = nullptr;
int g_memops = -1;   // -1 unknown, 0 kernel, 1 stream memory operations
unsigned long long g_memop_batches = 0;

bool memops_enabled() {
    static std::once_flag once;
    std::call_once(once, [] {
        g_memops = 0;
        const char* e = std::getenv("ELB200_P2P_MEMOPS");
        if (e && std::atoi(e) == 0) return;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamBatchMemOp", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess && ptr) {
            g_batch = (BatchMemOpFn)ptr;
            g_memops = 1;
        }
    });
    return g_memops == 1;
}
}  // namespace

int p2p_flag_mode() { return memops_enabled() ? 1 : 0; }

void p2p_flags(const P2PFlagOps& ops, cudaStream_t s) {
    if (ops.nsignal == 0 && ops.nwait == 0) return;
    if (ops.nsignal > P2P_MAX_PEERS || ops.nwait > P2P_MAX_PEERS) throw std::logic_error("p2p_flags: too many peers");
    if (memops_enabled()) {
        CUstreamBatchMemOpParams params[2 * P2P_MAX_PEERS];
        std::memset(params, 0, sizeof(CUstreamBatchMemOpParams) * (size_t)(ops.nsignal + ops.nwait));
        int n = 0;
        for (int i = 0; i < ops.nsignal; ++i, ++n) {
            params[n].operation = CU_STREAM_MEM_OP_WRITE_VALUE_32;
            params[n].writeValue.address = (CUdeviceptr)(uintptr_t)ops.signal[i];
            params[n].writeValue.value = ops.epoch;
            params[n].writeValue.flags = CU_STREAM_WRITE_VALUE_DEFAULT;   // prior writes of the stream are visible first
        }
        for (int i = 0; i < ops.nwait; ++i, ++n) {
            params[n].operation = CU_STREAM_MEM_OP_WAIT_VALUE_32;
            params[n].waitValue.address = (CUdeviceptr)(uintptr_t)ops.wait[i];
            params[n].waitValue.value = ops.wait_value;
            params[n].waitValue.flags = CU_STREAM_WAIT_VALUE_GEQ;
        }
        const CUresult r = g_batch((CUstream)s, (unsigned)n, params, 0);
        if (r == CUDA_SUCCESS) { ++g_memop_batches; return; }
        g_memops = 0;   // not supported here: the kernel from now on (nothing was enqueued by the failed call)
    }
    p2p_flags_kernel<<<1, 64, 0, s>>>(ops);
    ELB_LAUNCH_CHECK();
}

}  // namespace elb200
