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

void p2p_flags(const P2PFlagOps& ops, cudaStream_t s) {
    if (ops.nsignal == 0 && ops.nwait == 0) return;
    if (ops.nsignal > P2P_MAX_PEERS || ops.nwait > P2P_MAX_PEERS) throw std::logic_error("p2p_flags: too many peers");
    p2p_flags_kernel<<<1, 64, 0, s>>>(ops);
    ELB_LAUNCH_CHECK();
}

}  // namespace elb200
