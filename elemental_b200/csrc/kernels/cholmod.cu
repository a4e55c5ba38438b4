// Rank-w update / downdate of a Cholesky factor, one column panel at a time.
//   cholmod_panel_device <- cholesky::mod::LowerUpdate / LowerDowndate
//                           (src/lapack_like/factor/Cholesky/LowerMod.hpp:19-58, 124-167) with the reflectors of
//                           src/lapack_like/reflect/Householder/Row.hpp:20-95 and Hyperbolic/Row.hpp:15-55
// Column j of the factor and the rows j.. of V are transformed by one (hyperbolic) Householder reflector built from
// (L(j,j), V(j,:)); nothing else of L is read, so the columns can be processed in panels with V carrying the state.
//
// B200-first: the reference runs m dependent level-2 steps, each with a broadcast and an all-reduce across the
// process row.  Here the panel of L and all of V are REPLICATED (V is n x w with small w; the panel is gathered like
// the LU panel) and ONE cooperative kernel sweeps the nb columns: row slabs per CTA, the reflector of column j is
// computed by the CTA that owns row j and published through global memory (double-buffered by the parity of j),
// one grid barrier per column, no host round trip.
#include "coop.cuh"
#include "device_api.hpp"
#include "elb200_blas.h"

namespace elb200 {
namespace {

constexpr int CM_THREADS = 256;

template <class T> struct re_of { typedef T type; };
template <class R> struct re_of<cplx<R>> { typedef R type; };

template <class T> __device__ inline T cm_scale(T x, typename re_of<T>::type s) { return x * s; }
template <class T> __device__ inline T cm_div(T a, T b) { return a / b; }
template <class R> __device__ inline cplx<R> cm_div_c(cplx<R> a, cplx<R> b) {
    if (fabs((double)b.re) >= fabs((double)b.im)) {
        const R r = b.im / b.re, d = b.re + b.im * r;
        return mk((a.re + a.im * r) / d, (a.im - a.re * r) / d);
    }
    const R r = b.re / b.im, d = b.re * r + b.im;
    return mk((a.re * r + a.im) / d, (a.im * r - a.re) / d);
}
template <> __device__ inline c32_t cm_div<c32_t>(c32_t a, c32_t b) { return cm_div_c(a, b); }
template <> __device__ inline c64_t cm_div<c64_t>(c64_t a, c64_t b) { return cm_div_c(a, b); }
template <class T> __device__ inline typename re_of<T>::type cm_imag(T) { return 0; }
template <class R> __device__ inline R cm_imag(cplx<R> x) { return x.im; }

// L: M x nb panel whose row 0 is the diagonal row of its first column; V: M x w, same rows.  CTA b owns rows
// [b rows, (b + 1) rows).  pub: 2 (w + 1) scalars.  info: 1 when a downdate would leave the matrix indefinite.
template <class T, bool DOWN>
__global__ void __launch_bounds__(CM_THREADS) cholmod_kernel(i64 M, int nb, T* L, i64 ldl, T* V, i64 ldv, int w, int* info,
                                                             unsigned* bar, T* pub, i64 rows) {
    typedef typename re_of<T>::type R;
    typedef scalar_traits<T> st;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* u = reinterpret_cast<T*>(smem_raw);   // [w] reflector vector, then [1] coefficient
    __shared__ R red[CM_THREADS / 32];
    __shared__ T sScl;
    const unsigned nblk = gridDim.x;
    unsigned epoch = 0;
    const i64 r0 = (i64)blockIdx.x * rows, r1 = (r0 + rows < M) ? r0 + rows : M;
    const int t = threadIdx.x;
    for (int j = 0; j < nb; ++j) {
        __syncthreads();   // row j was updated by this CTA in the previous step
        T* mine = pub + (size_t)(j & 1) * (size_t)(w + 1);
        const bool owner = (j >= r0 && j < r1);
        if (owner) {
            R ss = 0;
            for (int c = t; c < w; c += CM_THREADS) ss += st::abs2(V[j + (i64)c * ldv]);
            for (int o = 16; o > 0; o >>= 1) ss += __shfl_down_sync(0xffffffffu, ss, o);
            if ((t & 31) == 0) red[t >> 5] = ss;
            __syncthreads();
            if (t == 0) {
                ss = 0;
                for (int q = 0; q < CM_THREADS / 32; ++q) ss += red[q];
                const T alpha = L[j + (i64)j * ldl];
                const R ar = st::real_part(alpha), ai = cm_imag(alpha);
                T scl, coef, diag;
                if (!DOWN) {
                    // Householder/Row.hpp: beta = -sign(Re alpha) ||(alpha, x)||, tau = (beta - conj(alpha)) / beta,
                    // x := conj(x / (alpha - beta)); LowerMod.hpp:44 then negates the new diagonal entry
                    if (ss == R(0) && ai == R(0)) {
                        scl = st::from_real(R(1)); coef = st::from_real(R(2)); diag = alpha;
                    } else {
                        const R nrm = sqrt((R)(ar * ar + ai * ai + ss));
                        const R beta = (ar <= R(0)) ? nrm : -nrm;
                        coef = cm_div(st::from_real(beta) - st::conj(alpha), st::from_real(beta));
                        scl = cm_div(st::from_real(R(1)), alpha - st::from_real(beta));
                        diag = st::from_real(-beta);
                    }
                } else {
                    // Hyperbolic/Row.hpp: delta = alpha^2 - ||x||^2, lambda = sign(alpha) sqrt(delta),
                    // x := conj(x / (alpha + lambda)), tau = (delta + alpha lambda) / (alpha + lambda)^2; 1 / tau is applied
                    R delta = ar * ar - ss;
                    if (delta < R(0)) { atomicCAS(info, 0, 1); delta = R(0); }
                    const R lam = (ar >= R(0)) ? sqrt(delta) : -sqrt(delta);
                    const R kappa = ar + lam;
                    if (kappa == R(0)) {
                        scl = st::zero(); coef = st::from_real(R(1));
                    } else {
                        scl = st::from_real(R(1) / kappa);
                        coef = st::from_real((kappa * kappa) / (delta + ar * lam));
                    }
                    diag = st::from_real(lam);
                }
                L[j + (i64)j * ldl] = diag;
                sScl = scl;
                u[w] = coef;
                mine[w] = coef;
            }
            __syncthreads();
            const T scl = sScl;
            for (int c = t; c < w; c += CM_THREADS) {
                const T x = st::conj(V[j + (i64)c * ldv] * scl);
                V[j + (i64)c * ldv] = x;   // the reference leaves the reflector vector in v1
                u[c] = x;
                mine[c] = x;
            }
        }
        grid_barrier(bar, nblk, epoch);
        if (!owner)
            for (int c = t; c <= w; c += CM_THREADS) u[c] = ldcg(&mine[c]);
        __syncthreads();
        const T coef = u[w];
        const i64 ib = (r0 > (i64)j + 1) ? r0 : (i64)j + 1;
        T* colj = L + (i64)j * ldl;
        for (i64 i = ib + t; i < r1; i += CM_THREADS) {
            // z = l_ij +- V(i,:) u^T;  l_ij := -l_ij + coef z;  V(i,:) := -V(i,:) + coef z conj(u)
            T acc = st::zero();
            for (int c = 0; c < w; ++c) acc += V[i + (i64)c * ldv] * u[c];
            const T l = colj[i];
            const T z = DOWN ? l - acc : l + acc;
            const T cz = coef * z;
            colj[i] = cz - l;
            for (int c = 0; c < w; ++c) {
                T* p = V + i + (i64)c * ldv;
                *p = cz * st::conj(u[c]) - *p;
            }
        }
    }
}

template <class T, bool DOWN>
void launch_cholmod(i64 M, int nb, T* L, i64 ldl, T* V, i64 ldv, int w, int* info, cudaStream_t s) {
    const size_t smem = sizeof(T) * (size_t)(w + 1);
    if (smem > 200 * 1024) throw std::logic_error("cholmod_panel: V is too wide for the kernel's shared reflector");
    static size_t optedIn = 0;
    if (smem > 48 * 1024 && smem > optedIn) {
        ELB_CUDA(cudaFuncSetAttribute(cholmod_kernel<T, DOWN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        optedIn = smem;
    }
    i64 grid = ceil_div(M, 256);
    const i64 maxGrid = sm_count();
    if (grid > maxGrid) grid = maxGrid;
    if (grid < 1) grid = 1;
    i64 rows = ceil_div(ceil_div(M, grid), 32) * 32;
    grid = ceil_div(M, rows);
    unsigned* bar = (unsigned*)scratch_alloc(256 + sizeof(T) * 2 * (size_t)(w + 1), s);
    T* pub = (T*)((char*)bar + 256);
    ELB_CUDA(cudaMemsetAsync(bar, 0, 256, s));
    void* args[] = {&M, &nb, &L, &ldl, &V, &ldv, &w, &info, &bar, &pub, &rows};
    ELB_CUDA(cudaLaunchCooperativeKernel((const void*)cholmod_kernel<T, DOWN>, dim3((unsigned)grid), dim3(CM_THREADS), args,
                                         smem, s));
    ++g_kernel_launches;
    scratch_free(bar, s);
}

}  // namespace

template <class T>
void cholmod_panel_device(bool downdate, i64 M, i64 nb, T* L, i64 ldl, T* V, i64 ldv, i64 w, int* info, cudaStream_t s) {
    if (M < 0 || nb < 0 || w < 0 || ldl < (M > 1 ? M : 1) || ldv < (M > 1 ? M : 1))
        throw std::logic_error("cholmod_panel: invalid argument");
    if (nb > M) throw std::logic_error("cholmod_panel: the panel must have at least as many rows as columns");
    if (M == 0 || nb == 0) return;
    if (downdate) launch_cholmod<T, true>(M, (int)nb, L, ldl, V, ldv, (int)w, info, s);
    else launch_cholmod<T, false>(M, (int)nb, L, ldl, V, ldv, (int)w, info, s);
}
template void cholmod_panel_device<float>(bool, i64, i64, float*, i64, float*, i64, i64, int*, cudaStream_t);
template void cholmod_panel_device<double>(bool, i64, i64, double*, i64, double*, i64, i64, int*, cudaStream_t);
template void cholmod_panel_device<c32_t>(bool, i64, i64, c32_t*, i64, c32_t*, i64, i64, int*, cudaStream_t);
template void cholmod_panel_device<c64_t>(bool, i64, i64, c64_t*, i64, c64_t*, i64, i64, int*, cudaStream_t);

}  // namespace elb200
