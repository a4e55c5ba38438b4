/*
 * elb200_plan.h -- the HOST-side planning logic of the redistribution engine, exposed so
 * that it can be exercised without a GPU (tests/test_redist_plan_gloo.py executes these
 * plans over torch.distributed/gloo on CPU tensors and compares with the oracle's
 * definition of every distribution).  The CUDA path (redist.cpp) consumes the very same
 * plans: pack kernel -> grouped ncclSend/ncclRecv (or ncclReduceScatter) -> unpack kernel.
 *
 * Replaces the index arithmetic spread over the reference's copy::* primitives
 * (include/El/blas_like/level1/Copy/*.hpp) and axpy_contract::* (level1/AxpyContract.hpp).
 */
#ifndef ELB200_PLAN_H
#define ELB200_PLAN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* distribution codes: 0 MC, 2 MR, 3 VC, 4 VR, 5 STAR (the reference's El::Dist values) */
typedef struct {
    int colDist, rowDist, colAlign, rowAlign;
} elb200_layout;

/* message between me and the process at grid position (peerRow, peerCol); element (t,u),
 * t < nrows, u < ncols, lives at src_local[s_off + t*s_rs + u*s_cs] on the sender and at
 * dst_local[d_off + t*d_rs + u*d_cs] on the receiver (offsets in elements).
 * kind: 0 = I send it, 1 = I receive it, 2 = local (I am both) */
typedef struct {
    int kind, peerRow, peerCol;
    int64_t nrows, ncols;
    int64_t s_off, s_rs, s_cs;
    int64_t d_off, d_rs, d_cs;
} elb200_plan_msg;

/* Plan of B = op(A) for the process at (myRow,myCol) of an r x c grid.  height/width are
 * B's global dimensions, ldA/ldB the local leading dimensions.  out must hold 2*r*c
 * entries; *nout receives the number written.  Returns 0 on success. */
int elb200_redist_plan(int r, int c, int myRow, int myCol, int64_t height, int64_t width,
                       elb200_layout A, int64_t ldA, elb200_layout B, int64_t ldB, int transpose,
                       elb200_plan_msg* out, int* nout);

/* Plan of the sum-scatter B += sum_replicas A for a partially replicated A ([MC,*], [*,MR],
 * [MR,*], [*,MC] or [*,*]); bLayout is B's layout as seen in A's index space.
 *   *commKind : 0 = row communicator (MR), 1 = column communicator (MC), 2 = all (VC order),
 *               3 = nothing is replicated (plain redistribution)
 *   *T        : layout of the summed, non-replicated intermediate
 *   *chunk    : elements per member in the reduce-scatter send buffer
 *   packs     : one entry per communicator member q (kind = q): my A_local lattice that
 *               goes to sendbuf + q*chunk (+ d_off, strides d_rs/d_cs); nrows == 0 if empty
 * packs must hold max(r*c,1) entries. */
int elb200_contract_plan(int r, int c, int myRow, int myCol, int64_t height, int64_t width,
                         elb200_layout A, int64_t ldA, elb200_layout bLayout, int* commKind,
                         elb200_layout* T, int64_t* chunk, elb200_plan_msg* packs, int* npacks);

/* index helpers (include/El/core/indexing/impl.hpp) */
int64_t elb200_shift(int64_t rank, int64_t align, int64_t stride);
int64_t elb200_length(int64_t n, int64_t shift, int64_t stride);
int elb200_dist_stride(int dist, int r, int c);
int elb200_dist_rank(int dist, int r, int c, int row, int col);
int elb200_gemm_default_algorithm(int64_t m, int64_t n, int64_t k);

/* Permutation bookkeeping of DistPermutation (src/lapack_like/perm/Permutation.cpp:333-347 MakeArbitrary, :300-322
 * Image / Preimage, Parity): compose `nswaps` swaps (origins[j] <-> dests[j], applied in order) of `size` indices into
 * the explicit vectors: row i of P A is row preimages[i] of A, images[preimages[i]] = i.  Returns 0, or 1 when an
 * index is out of range.  elb200_perm_parity: 1 when the permutation is odd, 0 when even. */
int elb200_perm_compose(int64_t size, int64_t nswaps, const int64_t* origins, const int64_t* dests, int64_t* preimages,
                        int64_t* images);
int elb200_perm_parity(int64_t size, const int64_t* preimages);

#ifdef __cplusplus
}
#endif
#endif
