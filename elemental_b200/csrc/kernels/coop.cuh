// Helpers of the cooperative (grid-synchronous) panel kernels: lu.cu, cholmod.cu, cholpiv.cu.
#pragma once
#include "../common.hpp"
#include "cplx.cuh"

namespace elb200 {
namespace {

// loads that bypass L1: data another CTA of the same launch has just written
template <class T> __device__ inline T ldcg(const T* p) { return __ldcg(p); }
template <> __device__ inline c32_t ldcg<c32_t>(const c32_t* p) {
    const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
    return mk(v.x, v.y);
}
template <> __device__ inline c64_t ldcg<c64_t>(const c64_t* p) {
    const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
    return mk(v.x, v.y);
}

// all CTAs of the (cooperatively launched, hence co-resident) grid meet; `bar` only grows
__device__ inline void grid_barrier(unsigned* bar, unsigned nblk, unsigned& epoch) {
    ++epoch;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
        const unsigned target = epoch * nblk;
        while (*(volatile unsigned*)bar < target) {}
        __threadfence();
    }
    __syncthreads();
}

}  // namespace
}  // namespace elb200
