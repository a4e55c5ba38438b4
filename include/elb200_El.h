/*
 * elb200_El.h -- the reference's C API for the Gemm / Cholesky / HPDSolve path, served by
 * the B200-native layer.  Names, argument order and ElError convention follow the
 * reference's own C headers so that a binding written against them (its Python ctypes
 * package, python/) keeps working:
 *   include/El/core/Grid.h:30-141, include/El/core/DistMatrix.h:58-,
 *   include/El/blas_like/level1.h (ElCopyDist, ElTransposeDist ...),
 *   include/El/blas_like/level3.h:29-92,575-704 (ElGemmDist, ElGemmXDist, ElHerkDist,
 *   ElTrsmDist, ElTrrkDist), include/El/lapack_like/factor.h:29-32 (ElCholeskyDist),
 *   include/El/lapack_like/solve.h:151-173 (ElHPDSolveDist),
 *   include/El/core/environment.h (ElBlocksize, ElSetBlocksize, Push/Pop).
 * Differences, all forced by the device: DistMatrix buffers are DEVICE pointers, a Grid is
 * created from a broadcast ncclUniqueId instead of an MPI_Comm (ElGridCreateNccl), and
 * complex scalars are passed as {re,im} structs.  Suffixes: _s float, _d double,
 * _c complex<float>, _z complex<double>.
 */
#ifndef ELB200_EL_H
#define ELB200_EL_H

#include <stdbool.h>
#include <stdint.h>
#include "elb200_blas.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef int ElInt;
typedef enum { EL_TRSM_DEFAULT = 0, EL_TRSM_LARGE = 1, EL_TRSM_MEDIUM = 2, EL_TRSM_SMALL = 3 } ElTrsmAlgorithm; /* enum TrsmAlgorithm, include/El/blas_like/level3.hpp:445-452 */
typedef enum {
    EL_SUCCESS, EL_ALLOC_ERROR, EL_OUT_OF_BOUNDS_ERROR, EL_ARG_ERROR, EL_LOGIC_ERROR, EL_RUNTIME_ERROR,
    EL_NON_HPD_ERROR = 100, EL_SINGULAR_ERROR = 101, /* extensions: reference maps these to EL_RUNTIME_ERROR */
    EL_ERROR = -1
} ElError;
typedef enum { EL_MC, EL_MD, EL_MR, EL_VC, EL_VR, EL_STAR, EL_CIRC } ElDist;
typedef enum { EL_NORMAL, EL_TRANSPOSE, EL_ADJOINT } ElOrientation;
typedef enum { EL_LOWER, EL_UPPER } ElUpperOrLower;
typedef enum { EL_LEFT, EL_RIGHT } ElLeftOrRight;
typedef enum { EL_NON_UNIT, EL_UNIT } ElUnitOrNonUnit;
typedef enum { EL_ROW_MAJOR, EL_COLUMN_MAJOR } ElGridOrderType;
typedef enum { EL_GEMM_DEFAULT, EL_GEMM_SUMMA_A, EL_GEMM_SUMMA_B, EL_GEMM_SUMMA_C, EL_GEMM_SUMMA_DOT,
               EL_GEMM_CANNON } ElGemmAlgorithm;

typedef float ElScalar_s;
typedef double ElScalar_d;
typedef elb200_c32 ElScalar_c;
typedef elb200_c64 ElScalar_z;

typedef struct ElGrid_sDummy* ElGrid;
typedef const struct ElGrid_sDummy* ElConstGrid;

const char* ElErrorString(ElError error);
const char* ElLastErrorMessage(void);

/* environment */
ElError ElInitialize(int* argc, char*** argv);
ElError ElFinalize(void);
ElError ElBlocksize(ElInt* blocksize);
ElError ElSetBlocksize(ElInt blocksize);
ElError ElPushBlocksizeStack(ElInt blocksize);
ElError ElPopBlocksizeStack(void);
/* edge of the C blocks of SUMMA_Dot (reference: blockSizeDot = 2000 hard-coded, Gemm/NN.hpp:233);
 * 0 = sized for HBM (default).  Results do not depend on it. */
/* DistPermutation (include/El/core/Permutation.h:49-160): a swap sequence; row i of P A is row Preimage(i) of A.
 * The swap list is device-resident (a factorisation appends its pivots without a host round trip).            */
typedef struct ElDistPermutationDummy* ElDistPermutation;
typedef const struct ElDistPermutationDummy* ElConstDistPermutation;
ElError ElDistPermutationCreate(ElDistPermutation* P, ElConstGrid g);
ElError ElDistPermutationDestroy(ElConstDistPermutation P);
ElError ElDistPermutationEmpty(ElDistPermutation P);
ElError ElDistPermutationMakeIdentity(ElDistPermutation P, ElInt size);
ElError ElDistPermutationReserveSwaps(ElDistPermutation P, ElInt maxSwaps);
ElError ElDistPermutationSwap(ElDistPermutation P, ElInt origin, ElInt dest);
ElError ElDistPermutationSwapSequence(ElDistPermutation P, ElConstDistPermutation PAppend, ElInt offset);
ElError ElDistPermutationHeight(ElConstDistPermutation P, ElInt* height);
ElError ElDistPermutationWidth(ElConstDistPermutation P, ElInt* width);
ElError ElDistPermutationParity(ElConstDistPermutation P, bool* parity);
ElError ElDistPermutationIsSwapSequence(ElConstDistPermutation P, bool* isSwap);
ElError ElDistPermutationIsImplicitSwapSequence(ElConstDistPermutation P, bool* isImplicit);
ElError ElDistPermutationImage(ElConstDistPermutation P, ElInt origin, ElInt* dest);
ElError ElDistPermutationPreimage(ElConstDistPermutation P, ElInt dest, ElInt* origin);
/* all Height() preimages at once (host array; no counterpart in the reference, which exposes them one at a time) */
ElError ElDistPermutationPreimages(ElConstDistPermutation P, ElInt* preimages);

ElError ElSetGemmDotBlocksize(ElInt blocksize);
ElError ElSetStream(elb200_stream_t stream);   /* stream all work is enqueued on */
ElError ElSynchronize(void);

/* Grid */
ElError ElNcclUniqueId(void* out128);          /* rank 0 calls this and broadcasts the 128 bytes */
ElError ElGridCreateNccl(const void* uniqueId128, int rank, int size, int height, ElGridOrderType order,
                         ElGrid* grid);
ElError ElGridCreateTrivial(ElGrid* grid);     /* single process, 1x1 */
ElError ElGridDestroy(ElConstGrid grid);
ElError ElGridHeight(ElConstGrid grid, int* height);
ElError ElGridWidth(ElConstGrid grid, int* width);
ElError ElGridSize(ElConstGrid grid, int* size);
ElError ElGridRank(ElConstGrid grid, int* rank);
ElError ElGridRow(ElConstGrid grid, int* row);
ElError ElGridCol(ElConstGrid grid, int* col);
ElError ElGridVCRank(ElConstGrid grid, int* rank);
ElError ElGridVRRank(ElConstGrid grid, int* rank);

/* redistribution-engine statistics since the last reset (copies, messages, bytesSent,
 * packLaunches, zeroCopySends, reduceScatters, allGathers, p2pPushes) */
ElError ElRedistStats(uint64_t out[8], bool reset);

#define ELB200_DECLARE_TYPE(SUF, SCALAR, REAL)                                                              \
    typedef struct ElDistMatrix_##SUF##Dummy* ElDistMatrix_##SUF;                                           \
    typedef const struct ElDistMatrix_##SUF##Dummy* ElConstDistMatrix_##SUF;                                \
    ElError ElDistMatrixCreateSpecific_##SUF(ElDist U, ElDist V, ElConstGrid g, ElDistMatrix_##SUF* A);     \
    ElError ElDistMatrixDestroy_##SUF(ElConstDistMatrix_##SUF A);                                           \
    ElError ElDistMatrixEmpty_##SUF(ElDistMatrix_##SUF A);                                                  \
    ElError ElDistMatrixResize_##SUF(ElDistMatrix_##SUF A, ElInt height, ElInt width);                      \
    ElError ElDistMatrixAlign_##SUF(ElDistMatrix_##SUF A, int colAlign, int rowAlign, bool constrain);      \
    ElError ElDistMatrixAlignWith_##SUF(ElDistMatrix_##SUF A, ElConstDistMatrix_##SUF B);                   \
    ElError ElDistMatrixAttach_##SUF(ElDistMatrix_##SUF A, ElInt height, ElInt width, ElConstGrid g,        \
                                     int colAlign, int rowAlign, SCALAR* deviceBuffer, ElInt ldim, int root); \
    ElError ElDistMatrixLockedAttach_##SUF(ElDistMatrix_##SUF A, ElInt height, ElInt width, ElConstGrid g,  \
                                           int colAlign, int rowAlign, const SCALAR* deviceBuffer,          \
                                           ElInt ldim, int root);                                           \
    ElError ElDistMatrixView_##SUF(ElDistMatrix_##SUF A, ElDistMatrix_##SUF parent, ElInt i, ElInt j,       \
                                   ElInt height, ElInt width);                                              \
    ElError ElDistMatrixHeight_##SUF(ElConstDistMatrix_##SUF A, ElInt* v);                                  \
    ElError ElDistMatrixWidth_##SUF(ElConstDistMatrix_##SUF A, ElInt* v);                                   \
    ElError ElDistMatrixLocalHeight_##SUF(ElConstDistMatrix_##SUF A, ElInt* v);                             \
    ElError ElDistMatrixLocalWidth_##SUF(ElConstDistMatrix_##SUF A, ElInt* v);                              \
    ElError ElDistMatrixLDim_##SUF(ElConstDistMatrix_##SUF A, ElInt* v);                                    \
    ElError ElDistMatrixColAlign_##SUF(ElConstDistMatrix_##SUF A, int* v);                                  \
    ElError ElDistMatrixRowAlign_##SUF(ElConstDistMatrix_##SUF A, int* v);                                  \
    ElError ElDistMatrixColShift_##SUF(ElConstDistMatrix_##SUF A, int* v);                                  \
    ElError ElDistMatrixRowShift_##SUF(ElConstDistMatrix_##SUF A, int* v);                                  \
    ElError ElDistMatrixColStride_##SUF(ElConstDistMatrix_##SUF A, int* v);                                 \
    ElError ElDistMatrixRowStride_##SUF(ElConstDistMatrix_##SUF A, int* v);                                 \
    ElError ElDistMatrixBuffer_##SUF(ElDistMatrix_##SUF A, SCALAR** deviceBuffer);                          \
    ElError ElDistMatrixLockedBuffer_##SUF(ElConstDistMatrix_##SUF A, const SCALAR** deviceBuffer);         \
    /* whole local matrix <-> host column-major buffer (synchronous) */                                     \
    ElError ElDistMatrixLocalToHost_##SUF(ElConstDistMatrix_##SUF A, SCALAR* host, ElInt hostLDim);         \
    ElError ElDistMatrixLocalFromHost_##SUF(ElDistMatrix_##SUF A, const SCALAR* host, ElInt hostLDim);      \
    ElError ElDistMatrixHashFill_##SUF(ElDistMatrix_##SUF A, int kind, uint64_t seed, double diag);         \
    /* level 1 */                                                                                           \
    ElError ElCopyDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B);                              \
    ElError ElTransposeDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B);                         \
    ElError ElAdjointDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B);                           \
    ElError ElAxpyDist_##SUF(SCALAR alpha, ElConstDistMatrix_##SUF X, ElDistMatrix_##SUF Y);                \
    ElError ElAxpyContractDist_##SUF(SCALAR alpha, ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B);        \
    ElError ElContractDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B);                          \
    ElError ElScaleDist_##SUF(SCALAR alpha, ElDistMatrix_##SUF A);                                          \
    ElError ElZeroDist_##SUF(ElDistMatrix_##SUF A);                                                         \
    ElError ElScaleTrapezoidDist_##SUF(SCALAR alpha, ElUpperOrLower uplo, ElDistMatrix_##SUF A, ElInt offset); \
    /* BINARY (3) / BINARY_FLAT (4) matrix files of the reference (include/El/io.h: ElReadDist, ElWriteDist;   */ \
    /* src/io/Read/{Binary,BinaryFlat}.hpp, Write/{Binary,BinaryFlat}.hpp) into / from device-resident matrices */ \
    ElError ElReadBinaryFlatDist_##SUF(ElDistMatrix_##SUF A, ElInt height, ElInt width, const char* filename); \
    ElError ElReadBinaryDist_##SUF(ElDistMatrix_##SUF A, const char* filename);                             \
    ElError ElWriteDist_##SUF(ElConstDistMatrix_##SUF A, const char* basename, int format);                 \
    /* ElAxpyTrapezoidDist (include/El/blas_like/level1.h): Y_trap += alpha X_trap */                       \
    ElError ElAxpyTrapezoidDist_##SUF(ElUpperOrLower uplo, SCALAR alpha, ElConstDistMatrix_##SUF X,         \
                                      ElDistMatrix_##SUF Y, ElInt offset);                                  \
    ElError ElMakeTrapezoidalDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A, ElInt offset);           \
    ElError ElFrobeniusNormDist_##SUF(ElConstDistMatrix_##SUF A, REAL* norm);                               \
    ElError ElMaxNormDist_##SUF(ElConstDistMatrix_##SUF A, REAL* norm);                                     \
    /* level 3 */                                                                                           \
    ElError ElGemmDist_##SUF(ElOrientation orientationOfA, ElOrientation orientationOfB, SCALAR alpha,      \
                             ElConstDistMatrix_##SUF A, ElConstDistMatrix_##SUF B, SCALAR beta,             \
                             ElDistMatrix_##SUF C);                                                         \
    ElError ElGemmXDist_##SUF(ElOrientation orientationOfA, ElOrientation orientationOfB, SCALAR alpha,     \
                              ElConstDistMatrix_##SUF A, ElConstDistMatrix_##SUF B, SCALAR beta,            \
                              ElDistMatrix_##SUF C, ElGemmAlgorithm alg);                                   \
    /* El::Gemm on HOST-resident [MC,MR] local matrices (the buffers a reference DistMatrix owns): m, n, k */ \
    /* are the global sizes of op(A) op(B), A / B / C the column-major local matrices of this process      */ \
    /* (alignments 0) with leading dimensions lda / ldb / ldc.  Streamed through HBM with the host copies  */ \
    /* overlapped (csrc/host/stream_gemm.cpp); C is complete in host memory on return.                     */ \
    ElError ElGemmDistHost_##SUF(ElOrientation orientationOfA, ElOrientation orientationOfB, SCALAR alpha,  \
                                 ElConstGrid grid, ElInt m, ElInt n, ElInt k, const SCALAR* A, ElInt lda,   \
                                 const SCALAR* B, ElInt ldb, SCALAR beta, SCALAR* C, ElInt ldc,             \
                                 ElGemmAlgorithm alg);                                                      \
    ElError ElSyrkDist_##SUF(ElUpperOrLower uplo, ElOrientation orientation, SCALAR alpha,                  \
                             ElConstDistMatrix_##SUF A, SCALAR beta, ElDistMatrix_##SUF C);                 \
    ElError ElHerkDist_##SUF(ElUpperOrLower uplo, ElOrientation orientation, REAL alpha,                    \
                             ElConstDistMatrix_##SUF A, REAL beta, ElDistMatrix_##SUF C);                   \
    ElError ElTrrkDist_##SUF(ElUpperOrLower uplo, ElOrientation orientationOfA, ElOrientation orientationOfB, \
                             SCALAR alpha, ElConstDistMatrix_##SUF A, ElConstDistMatrix_##SUF B, SCALAR beta, \
                             ElDistMatrix_##SUF C);                                                         \
    ElError ElTrsmDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, ElOrientation orientation,            \
                             ElUnitOrNonUnit diag, SCALAR alpha, ElConstDistMatrix_##SUF A,                 \
                             ElDistMatrix_##SUF B);                                                         \
    /* El::Trsm with its two C++-only arguments (include/El/blas_like/level3.hpp:455-470): checkIfSingular throws  */ \
    /* EL_SINGULAR_ERROR on a zero diagonal entry (Trsm.cpp:54-60); alg picks Large / Medium / Small (LEFT only)   */ \
    ElError ElTrsmXDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, ElOrientation orientation,           \
                              ElUnitOrNonUnit diag, SCALAR alpha, ElConstDistMatrix_##SUF A,                \
                              ElDistMatrix_##SUF B, bool checkIfSingular, ElTrsmAlgorithm alg);             \
    /* El::Trsv (src/blas_like/level2/Trsv.cpp:47-68): x is n x 1 or 1 x n */                              \
    ElError ElTrsvDist_##SUF(ElUpperOrLower uplo, ElOrientation orientation, ElUnitOrNonUnit diag,          \
                             ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF x);                              \
    /* siblings over the same leaves (include/El/blas_like/level3.h:371-374 Symm, :477-480 Syr2k,           \
       :559-562 Trmm) */                                                                                    \
    ElError ElSymmDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, SCALAR alpha, ElConstDistMatrix_##SUF A, \
                             ElConstDistMatrix_##SUF B, SCALAR beta, ElDistMatrix_##SUF C);                 \
    ElError ElSyr2kDist_##SUF(ElUpperOrLower uplo, ElOrientation orientation, SCALAR alpha,                 \
                              ElConstDistMatrix_##SUF A, ElConstDistMatrix_##SUF B, SCALAR beta,            \
                              ElDistMatrix_##SUF C);                                                        \
    /* ElTwoSidedTrsmDist / ElTwoSidedTrmmDist (include/El/blas_like/level3.h) */                          \
    ElError ElTwoSidedTrsmDist_##SUF(ElUpperOrLower uplo, ElUnitOrNonUnit diag, ElDistMatrix_##SUF A,       \
                                     ElConstDistMatrix_##SUF B);                                            \
    ElError ElTwoSidedTrmmDist_##SUF(ElUpperOrLower uplo, ElUnitOrNonUnit diag, ElDistMatrix_##SUF A,       \
                                     ElConstDistMatrix_##SUF B);                                            \
    /* ElTrr2kDist (include/El/blas_like/level3.h): E_tri := alpha op(A) op(B) + beta op(C) op(D) + gamma E_tri */ \
    ElError ElTrr2kDist_##SUF(ElUpperOrLower uplo, ElOrientation orientA, ElOrientation orientB,            \
                              ElOrientation orientC, ElOrientation orientD, SCALAR alpha,                   \
                              ElConstDistMatrix_##SUF A, ElConstDistMatrix_##SUF B, SCALAR beta,            \
                              ElConstDistMatrix_##SUF C, ElConstDistMatrix_##SUF D, SCALAR gamma,           \
                              ElDistMatrix_##SUF E);                                                        \
    ElError ElTrmmDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, ElOrientation orientation,            \
                             ElUnitOrNonUnit diag, SCALAR alpha, ElConstDistMatrix_##SUF A,                 \
                             ElDistMatrix_##SUF B);                                                         \
    /* factor / solve */                                                                                    \
    ElError ElCholeskyDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A);                                \
    /* ElReverseCholeskyDist (include/El/lapack_like/factor.h): A = L^H L / U U^H; the left-looking variant 2 */ \
    /* (cholesky::{Lower,Upper}Variant2Blocked) has no C name in the reference: exposed for the tests          */ \
    ElError ElReverseCholeskyDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A);                         \
    ElError ElCholeskyVariant2Dist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A);                        \
    ElError ElCholeskySolveAfterDist_##SUF(ElUpperOrLower uplo, ElOrientation orientation,                  \
                                           ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B);                \
    ElError ElHPDSolveDist_##SUF(ElUpperOrLower uplo, ElOrientation orientation, ElConstDistMatrix_##SUF A, \
                                 ElDistMatrix_##SUF B);                                                     \
    /* LU (include/El/lapack_like/factor.h:465-520): without pivoting, with partial pivoting, the solves after */ \
    ElError ElLUDist_##SUF(ElDistMatrix_##SUF A);                                                           \
    ElError ElLUPartialPivDist_##SUF(ElDistMatrix_##SUF A, ElDistPermutation P);                            \
    ElError ElSolveAfterLUDist_##SUF(ElOrientation orientation, ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B); \
    ElError ElSolveAfterLUPartialPivDist_##SUF(ElOrientation orientation, ElConstDistMatrix_##SUF A,        \
                                               ElConstDistPermutation P, ElDistMatrix_##SUF B);             \
    /* ElCholeskyPivDist / ElSolveAfterCholeskyPivDist (include/El/lapack_like/factor.h:80-124): P A P^T = L L^H */ \
    ElError ElCholeskyPivDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A, ElDistPermutation P);        \
    ElError ElSolveAfterCholeskyPivDist_##SUF(ElUpperOrLower uplo, ElOrientation orientation,               \
                                              ElConstDistMatrix_##SUF A, ElConstDistPermutation P, ElDistMatrix_##SUF B); \
    /* ElCholeskyModDist (include/El/lapack_like/factor.h:127-145): T T^H + alpha V V^H = That That^H */    \
    ElError ElCholeskyModDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF T, REAL alpha, ElDistMatrix_##SUF V); \
    /* ElLinearSolveDist (include/El/lapack_like/solve.h:26-33) */                                          \
    ElError ElLinearSolveDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B);                       \
    /* ElDistPermutationPermuteRows / Cols and inverses (include/El/core/Permutation.h:118-160) */          \
    ElError ElDistPermutationPermuteRowsDist_##SUF(ElConstDistPermutation P, ElDistMatrix_##SUF A, ElInt offset); \
    ElError ElDistPermutationInversePermuteRowsDist_##SUF(ElConstDistPermutation P, ElDistMatrix_##SUF A, ElInt offset); \
    ElError ElDistPermutationPermuteColsDist_##SUF(ElConstDistPermutation P, ElDistMatrix_##SUF A, ElInt offset); \
    ElError ElDistPermutationInversePermuteColsDist_##SUF(ElConstDistPermutation P, ElDistMatrix_##SUF A, ElInt offset);

ELB200_DECLARE_TYPE(s, float, float)
ELB200_DECLARE_TYPE(d, double, double)
ELB200_DECLARE_TYPE(c, elb200_c32, float)
ELB200_DECLARE_TYPE(z, elb200_c64, double)
/* Hermitian forms exist for the complex types only (include/El/blas_like/level3.h:109-112 Hemm, :163-166 Her2k) */
#define ELB200_DECLARE_HERMITIAN(SUF, SCALAR, REAL)                                                         \
    ElError ElHemmDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, SCALAR alpha, ElConstDistMatrix_##SUF A, \
                             ElConstDistMatrix_##SUF B, SCALAR beta, ElDistMatrix_##SUF C);                 \
    ElError ElHer2kDist_##SUF(ElUpperOrLower uplo, ElOrientation orientation, SCALAR alpha,                 \
                              ElConstDistMatrix_##SUF A, ElConstDistMatrix_##SUF B, REAL beta,              \
                              ElDistMatrix_##SUF C);
ELB200_DECLARE_HERMITIAN(c, elb200_c32, float)
ELB200_DECLARE_HERMITIAN(z, elb200_c64, double)

#ifdef __cplusplus
}
#endif
#endif
