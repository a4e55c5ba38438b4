// TEST INFRASTRUCTURE (depth-1 link test): the reference allocates every Matrix /
// DistMatrix buffer with `new G[size]` (include/El/core/Memory/impl.hpp:14-26).  Linked into libElRefDev.so with
// -Bsymbolic-functions, this replaces operator new[] / delete[] FOR THAT LIBRARY ONLY by page-locked, device-mapped
// host memory, which makes every buffer the reference hands to dgemm_ / dtrsm_ / ... addressable by the GPU without
// touching a line of the reference -- the effect of the Memory<G> allocator patch a maintainer would apply.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <new>
#include <unordered_set>

namespace {
std::mutex g_mu;
std::unordered_set<void*>& Owned() { static std::unordered_set<void*>* s = new std::unordered_set<void*>(); return *s; }
unsigned long long g_allocs = 0;
}

void* operator new[](std::size_t n) {
    void* p = nullptr;
    if (n == 0) n = 1;
    const cudaError_t e = cudaHostAlloc(&p, n, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e != cudaSuccess || !p) {
        std::fprintf(stderr, "[devalloc] cudaHostAlloc(%zu bytes) failed: %s\n", n, cudaGetErrorString(e));
        cudaGetLastError();
        throw std::bad_alloc();
    }
    std::lock_guard<std::mutex> l(g_mu);
    Owned().insert(p);
    ++g_allocs;
    return p;
}
void operator delete[](void* p) noexcept {
    if (!p) return;
    bool mine;
    { std::lock_guard<std::mutex> l(g_mu); mine = Owned().erase(p) != 0; }
    if (mine) { cudaDeviceSynchronize(); cudaFreeHost(p); }
    else std::free(p);
}
void operator delete[](void* p, std::size_t) noexcept { operator delete[](p); }
extern "C" unsigned long long elref_device_allocs(void) { return g_allocs; }
