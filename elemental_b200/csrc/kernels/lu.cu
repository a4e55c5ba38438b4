// LU panel factorisation with partial pivoting, and the row-interchange kernels around it.
//   getrf_panel_device   <- lu::Panel (src/lapack_like/factor/LU/Panel.hpp:14-55 sequential, :62-157 distributed)
//                           and lu::Unb (LU/Local.hpp:44-61) when pivoting is off
//   swap_plan / pack / unpack <- DistPermutation::PermuteRows of the panel's swap sequence applied to the columns
//                           outside the panel (LU.cpp:203 `PB.PermuteRows( AB )`)
//   permute_rows / cols  <- Permutation::PermuteRows / PermuteCols with an explicit preimage vector
//                           (src/lapack_like/perm/Permutation.cpp:545-600)
//
// B200-first: the reference pivots one column at a time with a MaxLoc all-reduce and a row broadcast per column
// across the process column (2 nb latency-bound collectives per panel).  Here the whole (m - k) x nb panel is
// REPLICATED (one gather over NVLink; at n = 65536, nb = 128 it is 67 MB and stays in the 126 MB L2) and factored
// redundantly by ONE cooperative kernel per panel: every CTA owns a slab of rows, pivot candidates meet in global
// memory, two grid barriers per column, no host round trip -- the pivots stay on the device and feed the
// interchange kernels directly.  The pivot rule is the reference's (i?amax: largest |x|, |re| + |im| for complex,
// first occurrence), so the permutation is the one the reference (and LAPACK getrf) produces.
#include <cstdlib>

#include "coop.cuh"
#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int LU_THREADS = 512;
constexpr int LU_MAX_N = 512;   // widest panel (Blocksize()) the kernel's shared rows hold

struct Cand {
    double val;
    long long idx;
};

template <class T> __device__ inline double abs1(T x) { return fabs((double)x); }
template <> __device__ inline double abs1<c32_t>(c32_t x) { return (double)(fabsf(x.re) + fabsf(x.im)); }
template <> __device__ inline double abs1<c64_t>(c64_t x) { return fabs(x.re) + fabs(x.im); }

template <class T> __device__ inline T recip(T x) { return T(1) / x; }
template <class R> __device__ inline cplx<R> recip_c(cplx<R> x) {
    // Smith's formula: no overflow in |x|^2
    if (fabs((double)x.re) >= fabs((double)x.im)) {
        const R r = x.im / x.re, d = x.re + x.im * r;
        return mk(R(1) / d, -r / d);
    }
    const R r = x.re / x.im, d = x.re * r + x.im;
    return mk(r / d, R(-1) / d);
}
template <> __device__ inline c32_t recip<c32_t>(c32_t x) { return recip_c(x); }
template <> __device__ inline c64_t recip<c64_t>(c64_t x) { return recip_c(x); }

__device__ inline bool better(double v, long long i, double bv, long long bi) { return v > bv || (v == bv && i < bi); }

__device__ inline void block_best(double& v, long long& i, Cand* red) {
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const long long oi = __shfl_down_sync(0xffffffffu, i, o);
        if (better(ov, oi, v, i)) { v = ov; i = oi; }
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[w].val = v; red[w].idx = i; }
    __syncthreads();
    if (w == 0) {
        v = l < (LU_THREADS >> 5) ? red[l].val : -1.0;
        i = l < (LU_THREADS >> 5) ? red[l].idx : 0x7fffffffffffffffLL;
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, v, o);
            const long long oi = __shfl_down_sync(0xffffffffu, i, o);
            if (better(ov, oi, v, i)) { v = ov; i = oi; }
        }
    }
    __syncthreads();
}

// A: M x n panel (M >= n), column-major.  CTA b owns rows [b * rows, (b + 1) * rows).
template <class T, bool PIVOT>
__global__ void __launch_bounds__(LU_THREADS) lu_panel_kernel(i64 M, int n, T* A, i64 lda, i64* ipiv, int* info, i64 col0,
                                                              unsigned* bar, Cand* cand, i64 rows) {
    __shared__ T u[LU_MAX_N];   // the pivot row
    __shared__ T w[LU_MAX_N];   // row j before the interchange
    __shared__ Cand red[LU_THREADS / 32];
    __shared__ long long sPiv;
    __shared__ int sSingular;
    const unsigned nblk = gridDim.x;
    unsigned epoch = 0;
    const i64 r0 = (i64)blockIdx.x * rows, r1 = (r0 + rows < M) ? r0 + rows : M;
    const int t = threadIdx.x;

    if (PIVOT) {   // candidates of column 0
        double bv = -1.0;
        long long bi = 0x7fffffffffffffffLL;
        for (i64 i = r0 + t; i < r1; i += LU_THREADS) {
            const double v = abs1(A[i]);
            if (better(v, i, bv, bi)) { bv = v; bi = i; }
        }
        block_best(bv, bi, red);
        if (t == 0) { cand[blockIdx.x].val = bv; cand[blockIdx.x].idx = bi; }
    }
    for (int j = 0; j < n; ++j) {
        grid_barrier(bar, nblk, epoch);   // row j final, candidates of column j published
        if (PIVOT) {
            if (t < 32) {
                double bv = -1.0;
                long long bi = 0x7fffffffffffffffLL;
                for (unsigned b = t; b < nblk; b += 32) {
                    const double v = __ldcg(&cand[b].val);
                    const long long i = __ldcg(&cand[b].idx);
                    if (better(v, i, bv, bi)) { bv = v; bi = i; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_down_sync(0xffffffffu, bv, o);
                    const long long oi = __shfl_down_sync(0xffffffffu, bi, o);
                    if (better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                }
                if (t == 0) {
                    sSingular = !(bv > 0.0);
                    sPiv = (bv > 0.0) ? bi : (long long)j;
                }
            }
        } else if (t == 0) {
            sPiv = j;
            sSingular = scalar_traits<T>::is_zero(ldcg(&A[j + (i64)j * lda]));
        }
        __syncthreads();
        const i64 piv = sPiv;
        const bool singular = sSingular != 0;
        for (int c = t; c < n; c += LU_THREADS) {
            u[c] = ldcg(&A[piv + (i64)c * lda]);
            if (PIVOT) w[c] = ldcg(&A[j + (i64)c * lda]);
        }
        if (blockIdx.x == 0 && t == 0) {
            if (ipiv) ipiv[j] = piv;
            if (singular) atomicCAS(info, 0, (int)(col0 + j + 1));
        }
        if (PIVOT) {
            grid_barrier(bar, nblk, epoch);   // everyone holds both rows: they may now be overwritten
            if (piv != j) {
                if (j >= r0 && j < r1)
                    for (int c = t; c < n; c += LU_THREADS) A[j + (i64)c * lda] = u[c];
                if (piv >= r0 && piv < r1)
                    for (int c = t; c < n; c += LU_THREADS) A[piv + (i64)c * lda] = w[c];
            }
        }
        __syncthreads();
        // l := a(:, j) / u_j below the diagonal, then the rank-1 update of the columns to the right
        const i64 ib = (r0 > (i64)j + 1) ? r0 : (i64)j + 1;
        const i64 nr = r1 - ib;
        double bv = -1.0;
        long long bi = 0x7fffffffffffffffLL;
        if (nr > 0) {
            const T inv = singular ? scalar_traits<T>::zero() : recip(u[j]);
            T* colj = A + (i64)j * lda;
            for (i64 i = ib + t; i < r1; i += LU_THREADS) colj[i] = colj[i] * inv;
            __syncthreads();
            const int nc = n - j - 1;
            // rows fastest: a warp updates 32 consecutive rows of one column
            const unsigned unr = (unsigned)nr, total = unr * (unsigned)nc;   // rows <= 2^21 (launch_panel), nc < 512
            for (unsigned e = t; e < total; e += LU_THREADS) {
                const unsigned c = e / unr;
                const i64 i = ib + (e - c * unr);
                const int jj = j + 1 + (int)c;
                T* p = A + i + (i64)jj * lda;
                const T v = *p - colj[i] * u[jj];
                *p = v;
                if (PIVOT && c == 0) {
                    const double a = abs1(v);
                    if (better(a, i, bv, bi)) { bv = a; bi = i; }
                }
            }
        }
        if (PIVOT && j + 1 < n) {
            block_best(bv, bi, red);
            if (t == 0) { cand[blockIdx.x].val = bv; cand[blockIdx.x].idx = bi; }
        }
    }
}

// ---- second generation: IMPLICIT interchanges ------------------------------------------------------------------
// Rows do not move while the panel is factored: the pivot row of column j is read where it lies and marked done, the
// others are eliminated in place.  Nothing is overwritten that another CTA still reads, so ONE grid barrier per column
// suffices (the first generation needed a second one between reading and exchanging the two rows).  Every CTA keeps the
// same small bookkeeping (which physical row sits at positions 0 .. n-1, where each of the first n physical rows has
// been displaced to), which gives (i) LAPACK's tie-break -- first occurrence in the CURRENT order, not the physical
// one -- and (ii) ipiv[j] directly.  At the end the <= 2 n rows whose position differs from their physical index are
// staged and written to their places.
struct Cand2 {
    double val;
    long long key;   // current position of the row (tie-break)
    long long row;   // physical row
};
__device__ inline bool better2(double v, long long k, double bv, long long bk) { return v > bv || (v == bv && k < bk); }

template <class T, bool PIVOT>
__global__ void __launch_bounds__(LU_THREADS) lu_panel_ip_kernel(i64 M, int n, T* A, i64 lda, i64* ipiv, int* info, i64 col0,
                                                                 unsigned* bar, Cand2* cand, i64 rows, unsigned char* done,
                                                                 T* stage) {
    __shared__ T u[LU_MAX_N];
    __shared__ double rv[LU_THREADS / 32];
    __shared__ long long rk[LU_THREADS / 32], rr[LU_THREADS / 32];
    __shared__ long long rowAt[LU_MAX_N];   // physical row at position j < n
    __shared__ long long posTop[LU_MAX_N];  // current position of physical row r < n
    __shared__ long long sPiv;
    __shared__ int sSingular;
    const unsigned nblk = gridDim.x;
    unsigned epoch = 0;
    const i64 r0 = (i64)blockIdx.x * rows, r1 = (r0 + rows < M) ? r0 + rows : M;
    const int t = threadIdx.x;
    const long long BIG = 0x7fffffffffffffffLL;
    for (int c = t; c < n; c += LU_THREADS) { rowAt[c] = c; posTop[c] = c; }
    __syncthreads();

    auto block_reduce = [&](double& v, long long& k, long long& r) {
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, v, o);
            const long long ok = __shfl_down_sync(0xffffffffu, k, o);
            const long long orr = __shfl_down_sync(0xffffffffu, r, o);
            if (better2(ov, ok, v, k)) { v = ov; k = ok; r = orr; }
        }
        const int w = t >> 5, l = t & 31;
        if (l == 0) { rv[w] = v; rk[w] = k; rr[w] = r; }
        __syncthreads();
        if (w == 0) {
            v = l < (LU_THREADS >> 5) ? rv[l] : -1.0;
            k = l < (LU_THREADS >> 5) ? rk[l] : BIG;
            r = l < (LU_THREADS >> 5) ? rr[l] : 0;
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, v, o);
                const long long ok = __shfl_down_sync(0xffffffffu, k, o);
                const long long orr = __shfl_down_sync(0xffffffffu, r, o);
                if (better2(ov, ok, v, k)) { v = ov; k = ok; r = orr; }
            }
        }
        __syncthreads();
    };

    if (PIVOT) {   // candidates of column 0: every row, positions = physical indices
        double bv = -1.0;
        long long bk = BIG, br = 0;
        for (i64 i = r0 + t; i < r1; i += LU_THREADS) {
            const double v = abs1(A[i]);
            if (better2(v, i, bv, bk)) { bv = v; bk = i; br = i; }
        }
        block_reduce(bv, bk, br);
        if (t == 0) { cand[blockIdx.x].val = bv; cand[blockIdx.x].key = bk; cand[blockIdx.x].row = br; }
    }
    for (int j = 0; j < n; ++j) {
        grid_barrier(bar, nblk, epoch);   // every row carries the updates of columns < j; candidates of column j are out
        if (t < 32) {
            double bv = -1.0;
            long long bk = BIG, br = 0;
            if (PIVOT) {
                for (unsigned b = t; b < nblk; b += 32) {
                    const double v = __ldcg(&cand[b].val);
                    const long long k = __ldcg(&cand[b].key), r = __ldcg(&cand[b].row);
                    if (better2(v, k, bv, bk)) { bv = v; bk = k; br = r; }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    const double ov = __shfl_down_sync(0xffffffffu, bv, o);
                    const long long ok = __shfl_down_sync(0xffffffffu, bk, o);
                    const long long orr = __shfl_down_sync(0xffffffffu, br, o);
                    if (better2(ov, ok, bv, bk)) { bv = ov; bk = ok; br = orr; }
                }
            }
            if (t == 0) {
                long long p, q;
                int singular;
                if (PIVOT) {
                    singular = !(bv > 0.0);
                    p = singular ? rowAt[j] : br;     // nothing to choose from: LAPACK keeps the row at position j
                    q = singular ? (long long)j : bk;
                    const long long r = rowAt[j];     // always one of the first n physical rows
                    if (q != j) {
                        if (q < n) rowAt[q] = r;
                        posTop[r] = q;
                    }
                    rowAt[j] = p;
                    if (p < n) posTop[p] = j;
                } else {
                    p = j; q = j;
                    singular = scalar_traits<T>::is_zero(ldcg(&A[j + (i64)j * lda]));
                }
                sPiv = p;
                sSingular = singular;
                if (p >= r0 && p < r1) done[p] = 1;
                if (blockIdx.x == 0) {
                    if (ipiv) ipiv[j] = q;
                    if (singular) atomicCAS(info, 0, (int)(col0 + j + 1));
                }
            }
        }
        __syncthreads();
        const i64 piv = sPiv;
        const bool singular = sSingular != 0;
        for (int c = j + t; c < n; c += LU_THREADS) u[c] = ldcg(&A[piv + (i64)c * lda]);
        __syncthreads();
        const T inv = singular ? scalar_traits<T>::zero() : recip(u[j]);
        T* colj = A + (i64)j * lda;
        for (i64 i = r0 + t; i < r1; i += LU_THREADS)
            if (!done[i]) colj[i] = colj[i] * inv;
        __syncthreads();
        double bv = -1.0;
        long long bk = BIG, br = 0;
        // a warp takes 32 consecutive rows of one column at a time; its lanes keep their row's multiplier in a register
        // while the warp walks the columns
        const int nc = n - j - 1;
        const unsigned nchunk = (unsigned)((r1 - r0 + 31) >> 5);
        const unsigned lane = t & 31, warp = t >> 5, nwarp = LU_THREADS >> 5;
        // warps are spread over the chunks first, then over the columns of a chunk
        const unsigned perChunk = nwarp > nchunk ? nwarp / nchunk : 1;
        for (unsigned ch = warp / perChunk; ch < nchunk; ch += (nwarp + perChunk - 1) / perChunk) {
            const i64 i = r0 + (i64)ch * 32 + lane;
            const bool live = i < r1 && !done[i];
            const T l = live ? colj[i] : scalar_traits<T>::zero();
            if (!live) continue;
            T* row = A + i;
            int c = (int)(warp % perChunk);
            const int pc = (int)perChunk;
            if (PIVOT && c == 0 && nc > 0) {   // column j + 1 also yields the next pivot candidate of this row
                T* p = row + (i64)(j + 1) * lda;
                const T v = *p - l * u[j + 1];
                *p = v;
                const double a = abs1(v);
                const long long key = (i < n) ? posTop[i] : i;
                if (better2(a, key, bv, bk)) { bv = a; bk = key; br = i; }
                c += pc;
            }
            // four columns in flight per lane: the loads are independent, the compiler cannot hoist them past the
            // stores by itself
            for (; c + 3 * pc < nc; c += 4 * pc) {
                T* p0 = row + (i64)(j + 1 + c) * lda;
                T* p1 = p0 + (i64)pc * lda;
                T* p2 = p1 + (i64)pc * lda;
                T* p3 = p2 + (i64)pc * lda;
                const T a0 = *p0, a1 = *p1, a2 = *p2, a3 = *p3;
                *p0 = a0 - l * u[j + 1 + c];
                *p1 = a1 - l * u[j + 1 + c + pc];
                *p2 = a2 - l * u[j + 1 + c + 2 * pc];
                *p3 = a3 - l * u[j + 1 + c + 3 * pc];
            }
            for (; c < nc; c += pc) {
                T* p = row + (i64)(j + 1 + c) * lda;
                *p = *p - l * u[j + 1 + c];
            }
        }
        if (PIVOT && j + 1 < n) {
            block_reduce(bv, bk, br);
            if (t == 0) { cand[blockIdx.x].val = bv; cand[blockIdx.x].key = bk; cand[blockIdx.x].row = br; }
        }
    }
    if (!PIVOT) return;
    // the rows whose position is not their physical index: pivot rows go to 0 .. n-1, displaced top rows to where
    // their pivots came from.  Stage every source, meet, then write every destination.
    grid_barrier(bar, nblk, epoch);
    for (int slot = blockIdx.x; slot < 2 * n; slot += nblk) {
        i64 src;
        if (slot < n) { src = rowAt[slot]; if (src == slot) continue; }
        else { src = slot - n; if (posTop[src] < n) continue; }
        for (int c = t; c < n; c += LU_THREADS) stage[(i64)slot * n + c] = ldcg(&A[src + (i64)c * lda]);
    }
    grid_barrier(bar, nblk, epoch);
    for (int slot = blockIdx.x; slot < 2 * n; slot += nblk) {
        i64 dst;
        if (slot < n) { if (rowAt[slot] == slot) continue; dst = slot; }
        else { dst = posTop[slot - n]; if (dst < n) continue; }
        for (int c = t; c < n; c += LU_THREADS) A[dst + (i64)c * lda] = stage[(i64)slot * n + c];
    }
}

template <class T, bool PIVOT>
void launch_panel(i64 M, int n, T* A, i64 lda, i64* ipiv, int* info, i64 col0, cudaStream_t s) {
    static int maxGrid = 0;
    if (!maxGrid) {
        int per = 0, coop = 0, dev = 0;
        ELB_CUDA(cudaGetDevice(&dev));
        ELB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        if (!coop) throw std::runtime_error("getrf_panel: the device does not support cooperative launches");
        ELB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, lu_panel_kernel<T, PIVOT>, LU_THREADS, 0));
        if (per < 1) throw std::runtime_error("getrf_panel: kernel does not fit an SM");
        maxGrid = sm_count();   // one CTA per SM: the slabs are bandwidth-, not occupancy-bound
    }
    // as many CTAs as SMs once the panel has 64 rows per CTA: the slab of a CTA (rows x n) then stays in its L1, and the
    // per-column work of a thread is a handful of elements (measured: 256 rows per CTA left 3/4 of the SMs idle at
    // m = 8192 and made every column a 17 us walk through L2)
    static const int minRows = [] { const char* e = std::getenv("ELB200_LU_PANEL_ROWS"); return e ? std::atoi(e) : 64; }();
    i64 grid = ceil_div(M, minRows > 0 ? minRows : 64);
    if (grid > maxGrid) grid = maxGrid;
    if (grid < 1) grid = 1;
    i64 rows = ceil_div(ceil_div(M, grid), 32) * 32;
    if (rows > (i64(1) << 21)) throw std::logic_error("getrf_panel: panel taller than 2^21 rows per SM");
    grid = ceil_div(M, rows);
    static const int gen = [] { const char* e = std::getenv("ELB200_LU_PANEL_GEN"); return e ? std::atoi(e) : 2; }();
    if (gen == 1) {   // first generation: physical interchanges, two barriers per column (kept for A/B runs)
        unsigned* bar = (unsigned*)scratch_alloc(256 + sizeof(Cand) * (size_t)grid, s);
        Cand* cand = (Cand*)((char*)bar + 256);
        ELB_CUDA(cudaMemsetAsync(bar, 0, 256, s));
        void* args[] = {&M, &n, &A, &lda, &ipiv, &info, &col0, &bar, &cand, &rows};
        ELB_CUDA(cudaLaunchCooperativeKernel((const void*)lu_panel_kernel<T, PIVOT>, dim3((unsigned)grid), dim3(LU_THREADS), args, 0, s));
        ++g_kernel_launches;
        scratch_free(bar, s);
        return;
    }
    const size_t candBytes = (sizeof(Cand2) * (size_t)grid + 255) / 256 * 256;
    const size_t doneBytes = ((size_t)M + 255) / 256 * 256;
    const size_t stageBytes = sizeof(T) * 2 * (size_t)n * (size_t)n;
    char* ws = (char*)scratch_alloc(256 + candBytes + doneBytes + stageBytes, s);
    unsigned* bar = (unsigned*)ws;
    Cand2* cand = (Cand2*)(ws + 256);
    unsigned char* done = (unsigned char*)(ws + 256 + candBytes);
    T* stage = (T*)(ws + 256 + candBytes + doneBytes);
    ELB_CUDA(cudaMemsetAsync(ws, 0, 256 + candBytes + doneBytes, s));
    void* args[] = {&M, &n, &A, &lda, &ipiv, &info, &col0, &bar, &cand, &rows, &done, &stage};
    ELB_CUDA(cudaLaunchCooperativeKernel((const void*)lu_panel_ip_kernel<T, PIVOT>, dim3((unsigned)grid), dim3(LU_THREADS), args, 0, s));
    ++g_kernel_launches;
    scratch_free(ws, s);
}

// ---- interchange of the rows a panel's swap sequence touches -------------------------------------------------
// Slots: t < nb is row k + t; slot nb + j is row k + ipiv[j] when that row lies below the block and no earlier
// slot names it, else -1.  srcSlot[t] = the slot whose ORIGINAL row ends up in slot t's row after the swaps
// (j <-> ipiv[j], j = 0 .. nb-1, in order).
__global__ void __launch_bounds__(512) swap_plan_kernel(int nb, const i64* __restrict__ ipiv, i64 k, i64* slotRow, int* srcSlot) {
    __shared__ int partner[LU_MAX_N];   // slot that swap j exchanges with slot j
    __shared__ int lab[2 * LU_MAX_N];
    const int t = threadIdx.x;
    for (int j = t; j < nb; j += blockDim.x) {
        const i64 p = ipiv[j];
        int slot;
        if (p < nb) {
            slot = (int)p;
        } else {
            int first = j;
            for (int q = 0; q < j; ++q)
                if (ipiv[q] == p) { first = q; break; }
            slot = nb + first;
        }
        partner[j] = slot;
        slotRow[j] = k + j;
        slotRow[nb + j] = (p >= nb && slot == nb + j) ? k + p : -1;
    }
    for (int q = t; q < 2 * nb; q += blockDim.x) lab[q] = q;
    __syncthreads();
    if (t == 0)
        for (int j = 0; j < nb; ++j) {
            const int b = partner[j], x = lab[j];
            lab[j] = lab[b];
            lab[b] = x;
        }
    __syncthreads();
    for (int q = t; q < 2 * nb; q += blockDim.x) srcSlot[q] = lab[q];
}

// buf[slot + S * col] := A(localRow(slotRow[slot]), col) for the slots whose row this process owns
template <class T>
__global__ void __launch_bounds__(256) pack_rows_kernel(int S, const i64* __restrict__ slotRow, const int* __restrict__ srcSlot,
                                                        const T* __restrict__ A, i64 lda, i64 ncols, int align, int stride,
                                                        int rank, int shift, T* buf) {
    const int slot = threadIdx.x + blockIdx.x * blockDim.x;
    if (slot >= S) return;
    const i64 row = slotRow[slot];
    if (row < 0 || (int)((row + align) % stride) != rank) return;
    // a row that stays where it is and feeds nobody else need not travel -- but it may feed another slot, so only
    // skip when it is its own source
    if (srcSlot[slot] == slot) {
        bool feeds = false;
        for (int q = 0; q < S; ++q) feeds |= (q != slot && srcSlot[q] == slot);
        if (!feeds) return;
    }
    const i64 iLoc = (row - shift) / stride;
    for (i64 c = blockIdx.y; c < ncols; c += gridDim.y) buf[slot + (i64)S * c] = A[iLoc + c * lda];
}
// A(localRow(slotRow[slot]), col) := all[owner(source row)][srcSlot[slot] + S * col]
template <class T>
__global__ void __launch_bounds__(256) unpack_rows_kernel(int S, const i64* __restrict__ slotRow, const int* __restrict__ srcSlot,
                                                          T* A, i64 lda, i64 ncols, int align, int stride, int rank, int shift,
                                                          const T* __restrict__ all, i64 perRank) {
    const int slot = threadIdx.x + blockIdx.x * blockDim.x;
    if (slot >= S) return;
    const i64 row = slotRow[slot];
    if (row < 0 || (int)((row + align) % stride) != rank) return;
    const int src = srcSlot[slot];
    if (src == slot) return;
    const i64 srow = slotRow[src];
    const int owner = (int)((srow + align) % stride);
    const T* from = all + (i64)owner * perRank + src;
    const i64 iLoc = (row - shift) / stride;
    for (i64 c = blockIdx.y; c < ncols; c += gridDim.y) A[iLoc + c * lda] = from[(i64)S * c];
}

// dst(iLoc, c) := src(perm[shift + iLoc * stride], c): rows of a full-height source gathered into a distributed
// destination (rowwise = true), or the same along columns
template <class T>
__global__ void __launch_bounds__(256) permute_kernel(int rowwise, i64 mloc, i64 nloc, const i64* __restrict__ perm, int shift,
                                                      int stride, const T* __restrict__ src, i64 lds, T* dst, i64 ldd) {
    const i64 i = (i64)blockIdx.x * 256 + threadIdx.x;
    if (i >= mloc) return;
    if (rowwise) {
        const i64 from = perm[shift + i * stride];
        for (i64 c = blockIdx.y; c < nloc; c += gridDim.y) dst[i + c * ldd] = src[from + c * lds];
    } else {
        for (i64 c = blockIdx.y; c < nloc; c += gridDim.y) dst[i + c * ldd] = src[i + perm[shift + c * stride] * lds];
    }
}

__global__ void append_swaps_kernel(i64* origins, i64* dests, i64 at, const i64* __restrict__ ipiv, int count, i64 offset) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < count) { origins[at + j] = offset + j; dests[at + j] = offset + ipiv[j]; }
}

template <class T>
void getrf_panel_t(i64 m, i64 n, T* A, i64 lda, i64* ipiv, int pivot, int* info, i64 col0, cudaStream_t s) {
    if (m < 0 || n < 0 || lda < (m > 1 ? m : 1)) throw std::logic_error("getrf_panel: invalid argument");
    if (m < n) throw std::logic_error("getrf_panel: the panel must have at least as many rows as columns");
    if (n > LU_MAX_N) throw std::logic_error("getrf_panel: panel wider than 512 columns");
    if (pivot && !ipiv) throw std::logic_error("getrf_panel: ipiv is required with pivoting");
    if (n == 0) return;
    if (pivot) launch_panel<T, true>(m, (int)n, A, lda, ipiv, info, col0, s);
    else launch_panel<T, false>(m, (int)n, A, lda, ipiv, info, col0, s);
}

}  // namespace

template <class T>
void getrf_panel_device(i64 m, i64 n, T* A, i64 lda, i64* ipiv, bool pivot, int* info, i64 col0, cudaStream_t s) {
    getrf_panel_t<T>(m, n, A, lda, ipiv, pivot ? 1 : 0, info, col0, s);
}
void swap_plan_device(int nb, const i64* ipiv, i64 k, i64* slotRow, int* srcSlot, cudaStream_t s) {
    if (nb <= 0) return;
    if (nb > LU_MAX_N) throw std::logic_error("swap_plan: more than 512 swaps in one panel");
    swap_plan_kernel<<<1, 512, 0, s>>>(nb, ipiv, k, slotRow, srcSlot);
    ELB_LAUNCH_CHECK();
}
void append_swaps_device(i64* origins, i64* dests, i64 at, const i64* ipiv, int count, i64 offset, cudaStream_t s) {
    if (count <= 0) return;
    append_swaps_kernel<<<(unsigned)ceil_div(count, 256), 256, 0, s>>>(origins, dests, at, ipiv, count, offset);
    ELB_LAUNCH_CHECK();
}
template <class T>
void pack_rows_device(int S, const i64* slotRow, const int* srcSlot, const T* A, i64 lda, i64 ncols, int align, int stride,
                      int rank, int shift, T* buf, cudaStream_t s) {
    if (S <= 0 || ncols <= 0) return;
    dim3 grid((unsigned)ceil_div(S, 256), (unsigned)(ncols < 4096 ? ncols : 4096));
    pack_rows_kernel<T><<<grid, 256, 0, s>>>(S, slotRow, srcSlot, A, lda, ncols, align, stride, rank, shift, buf);
    ELB_LAUNCH_CHECK();
}
template <class T>
void unpack_rows_device(int S, const i64* slotRow, const int* srcSlot, T* A, i64 lda, i64 ncols, int align, int stride,
                        int rank, int shift, const T* all, i64 perRank, cudaStream_t s) {
    if (S <= 0 || ncols <= 0) return;
    dim3 grid((unsigned)ceil_div(S, 256), (unsigned)(ncols < 4096 ? ncols : 4096));
    unpack_rows_kernel<T><<<grid, 256, 0, s>>>(S, slotRow, srcSlot, A, lda, ncols, align, stride, rank, shift, all, perRank);
    ELB_LAUNCH_CHECK();
}
template <class T>
void permute_device(bool rowwise, i64 mloc, i64 nloc, const i64* perm, int shift, int stride, const T* src, i64 lds, T* dst,
                    i64 ldd, cudaStream_t s) {
    if (mloc <= 0 || nloc <= 0) return;
    dim3 grid((unsigned)ceil_div(mloc, 256), (unsigned)(nloc < 2048 ? nloc : 2048));
    permute_kernel<T><<<grid, 256, 0, s>>>(rowwise ? 1 : 0, mloc, nloc, perm, shift, stride, src, lds, dst, ldd);
    ELB_LAUNCH_CHECK();
}

#define ELB_LU_INST(T)                                                                                                  \
    template void getrf_panel_device<T>(i64, i64, T*, i64, i64*, bool, int*, i64, cudaStream_t);                        \
    template void pack_rows_device<T>(int, const i64*, const int*, const T*, i64, i64, int, int, int, int, T*,          \
                                      cudaStream_t);                                                                    \
    template void unpack_rows_device<T>(int, const i64*, const int*, T*, i64, i64, int, int, int, int, const T*, i64,   \
                                        cudaStream_t);                                                                  \
    template void permute_device<T>(bool, i64, i64, const i64*, int, int, const T*, i64, T*, i64, cudaStream_t);
ELB_LU_INST(float)
ELB_LU_INST(double)
ELB_LU_INST(c32_t)
ELB_LU_INST(c64_t)
template void permute_device<i64>(bool, i64, i64, const i64*, int, int, const i64*, i64, i64*, i64, cudaStream_t);

}  // namespace elb200

extern "C" {
using namespace elb200;
#define ELB_GETRF(P, T, CT)                                                                                            \
    int elb200_##P##getrf_panel(int64_t m, int64_t n, CT* A, int64_t lda, int64_t* ipiv, int pivot, int* info,          \
                                elb200_stream_t s) {                                                                    \
        return guarded([&] { getrf_panel_t<T>(m, n, (T*)A, lda, (i64*)ipiv, pivot, info, 0, (cudaStream_t)s); });       \
    }
ELB_GETRF(s, float, float)
ELB_GETRF(d, double, double)
ELB_GETRF(c, c32_t, elb200_c32)
ELB_GETRF(z, c64_t, elb200_c64)
}
