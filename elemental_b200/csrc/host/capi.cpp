// C API (include/elb200_El.h): the reference's ElXxxDist_{s,d,c,z} entry points over the
// B200-native host layer.  Exceptions become ElError codes the way the reference's
// EL_TRY / EL_CATCH do (include/El/core/CReflect.hpp:12-26).
#include <cstring>

#include "dev.hpp"
#include "elb200/factor.hpp"
#include "elb200/io.hpp"
#include "elb200/lu.hpp"
#include "elb200_El.h"

using namespace El;

namespace {
thread_local std::string g_msg;

template <class F>
ElError Try(F&& f) {
    try {
        f();
        return EL_SUCCESS;
    } catch (const NonHPDMatrixException& e) { g_msg = e.what(); return EL_NON_HPD_ERROR; }
    catch (const SingularMatrixException& e) { g_msg = e.what(); return EL_SINGULAR_ERROR; }
    catch (const std::bad_alloc& e) { g_msg = e.what(); return EL_ALLOC_ERROR; }
    catch (const std::out_of_range& e) { g_msg = e.what(); return EL_OUT_OF_BOUNDS_ERROR; }
    catch (const std::invalid_argument& e) { g_msg = e.what(); return EL_ARG_ERROR; }
    catch (const std::logic_error& e) { g_msg = e.what(); return EL_LOGIC_ERROR; }
    catch (const std::runtime_error& e) { g_msg = e.what(); return EL_RUNTIME_ERROR; }
    catch (const std::exception& e) { g_msg = e.what(); return EL_ERROR; }
    catch (...) { g_msg = "unknown exception"; return EL_ERROR; }
}

inline const Grid* G(ElConstGrid g) { return reinterpret_cast<const Grid*>(g); }
inline Orientation O(ElOrientation o) { return static_cast<Orientation>(o); }
inline UpperOrLower UL(ElUpperOrLower u) { return static_cast<UpperOrLower>(u); }
inline Dist D(ElDist d) { return static_cast<Dist>(d); }

template <typename T, typename S> inline T Sc(S x);
template <> inline float Sc<float, float>(float x) { return x; }
template <> inline double Sc<double, double>(double x) { return x; }
template <> inline Complex<float> Sc<Complex<float>, elb200_c32>(elb200_c32 x) { return Complex<float>(x.re, x.im); }
template <> inline Complex<double> Sc<Complex<double>, elb200_c64>(elb200_c64 x) { return Complex<double>(x.re, x.im); }
}  // namespace

extern "C" {

const char* ElErrorString(ElError e) {
    switch (e) {
        case EL_SUCCESS: return "EL_SUCCESS";
        case EL_ALLOC_ERROR: return "EL_ALLOC_ERROR";
        case EL_OUT_OF_BOUNDS_ERROR: return "EL_OUT_OF_BOUNDS_ERROR";
        case EL_ARG_ERROR: return "EL_ARG_ERROR";
        case EL_LOGIC_ERROR: return "EL_LOGIC_ERROR";
        case EL_RUNTIME_ERROR: return "EL_RUNTIME_ERROR";
        case EL_NON_HPD_ERROR: return "EL_NON_HPD_ERROR";
        case EL_SINGULAR_ERROR: return "EL_SINGULAR_ERROR";
        default: return "EL_ERROR";
    }
}
const char* ElLastErrorMessage(void) { return g_msg.c_str(); }

ElError ElInitialize(int*, char***) {
    return Try([] { dev::c_check(elb200_device_check(), "elb200_device_check"); });
}
ElError ElFinalize(void) { return EL_SUCCESS; }
ElError ElBlocksize(ElInt* b) { return Try([&] { *b = Blocksize(); }); }
ElError ElSetBlocksize(ElInt b) { return Try([&] { SetBlocksize(b); }); }
ElError ElPushBlocksizeStack(ElInt b) { return Try([&] { PushBlocksizeStack(b); }); }
ElError ElPopBlocksizeStack(void) { return Try([] { PopBlocksizeStack(); }); }
ElError ElSetStream(elb200_stream_t s) { return Try([&] { SetCurrentStream((Stream)s); }); }
ElError ElSynchronize(void) { return Try([] { SynchronizeStream(); }); }

ElError ElNcclUniqueId(void* out128) {
    return Try([&] {
        ncclUniqueId id;
        ELB_NCCL(ncclGetUniqueId(&id));
        static_assert(sizeof(id) == 128, "ncclUniqueId is expected to be 128 bytes");
        std::memcpy(out128, &id, sizeof(id));
    });
}
ElError ElGridCreateNccl(const void* uid, int rank, int size, int height, ElGridOrderType order, ElGrid* grid) {
    return Try([&] {
        *grid = reinterpret_cast<ElGrid>(new Grid(uid, rank, size, height, order == EL_ROW_MAJOR ? ROW_MAJOR : COLUMN_MAJOR));
    });
}
ElError ElGridCreateTrivial(ElGrid* grid) { return Try([&] { *grid = reinterpret_cast<ElGrid>(new Grid()); }); }
ElError ElGridDestroy(ElConstGrid g) { return Try([&] { delete G(g); }); }
ElError ElGridHeight(ElConstGrid g, int* v) { return Try([&] { *v = G(g)->Height(); }); }
ElError ElGridWidth(ElConstGrid g, int* v) { return Try([&] { *v = G(g)->Width(); }); }
ElError ElGridSize(ElConstGrid g, int* v) { return Try([&] { *v = G(g)->Size(); }); }
ElError ElGridRank(ElConstGrid g, int* v) { return Try([&] { *v = G(g)->Rank(); }); }
ElError ElGridRow(ElConstGrid g, int* v) { return Try([&] { *v = G(g)->Row(); }); }
ElError ElGridCol(ElConstGrid g, int* v) { return Try([&] { *v = G(g)->Col(); }); }
ElError ElGridVCRank(ElConstGrid g, int* v) { return Try([&] { *v = G(g)->VCRank(); }); }
ElError ElGridVRRank(ElConstGrid g, int* v) { return Try([&] { *v = G(g)->VRRank(); }); }

static inline DistPermutation* PM(ElDistPermutation P) { return reinterpret_cast<DistPermutation*>(P); }
static inline const DistPermutation* CPM(ElConstDistPermutation P) { return reinterpret_cast<const DistPermutation*>(P); }
ElError ElDistPermutationCreate(ElDistPermutation* P, ElConstGrid g) {
    return Try([&] { *P = reinterpret_cast<ElDistPermutation>(new DistPermutation(*G(g))); });
}
ElError ElDistPermutationDestroy(ElConstDistPermutation P) { return Try([&] { delete CPM(P); }); }
ElError ElDistPermutationEmpty(ElDistPermutation P) { return Try([&] { PM(P)->Empty(); }); }
ElError ElDistPermutationMakeIdentity(ElDistPermutation P, ElInt size) { return Try([&] { PM(P)->MakeIdentity(size); }); }
ElError ElDistPermutationReserveSwaps(ElDistPermutation P, ElInt maxSwaps) { return Try([&] { PM(P)->ReserveSwaps(maxSwaps); }); }
ElError ElDistPermutationSwap(ElDistPermutation P, ElInt origin, ElInt dest) { return Try([&] { PM(P)->Swap(origin, dest); }); }
ElError ElDistPermutationSwapSequence(ElDistPermutation P, ElConstDistPermutation PAppend, ElInt offset) {
    return Try([&] { PM(P)->SwapSequence(*CPM(PAppend), offset); });
}
ElError ElDistPermutationHeight(ElConstDistPermutation P, ElInt* h) { return Try([&] { *h = CPM(P)->Height(); }); }
ElError ElDistPermutationWidth(ElConstDistPermutation P, ElInt* w) { return Try([&] { *w = CPM(P)->Width(); }); }
ElError ElDistPermutationParity(ElConstDistPermutation P, bool* parity) { return Try([&] { *parity = CPM(P)->Parity(); }); }
ElError ElDistPermutationIsSwapSequence(ElConstDistPermutation P, bool* b) { return Try([&] { *b = CPM(P)->IsSwapSequence(); }); }
ElError ElDistPermutationIsImplicitSwapSequence(ElConstDistPermutation P, bool* b) {
    return Try([&] { *b = CPM(P)->IsImplicitSwapSequence(); });
}
ElError ElDistPermutationImage(ElConstDistPermutation P, ElInt origin, ElInt* dest) { return Try([&] { *dest = CPM(P)->Image(origin); }); }
ElError ElDistPermutationPreimage(ElConstDistPermutation P, ElInt dest, ElInt* origin) {
    return Try([&] { *origin = CPM(P)->Preimage(dest); });
}
ElError ElDistPermutationPreimages(ElConstDistPermutation P, ElInt* out) {
    return Try([&] {
        const std::vector<Int> v = CPM(P)->Preimages();
        std::copy(v.begin(), v.end(), out);
    });
}

ElError ElSetGemmDotBlocksize(ElInt b) { return Try([&] { SetGemmDotBlocksize(b); }); }
ElError ElRedistStats(uint64_t out[8], bool reset) {
    return Try([&] {
        RedistStats& s = GetRedistStats();
        out[0] = s.copies; out[1] = s.messages; out[2] = s.bytesSent; out[3] = s.packLaunches;
        out[4] = s.zeroCopySends; out[5] = s.reduceScatters; out[6] = s.allGathers; out[7] = s.p2pPushes;
        if (reset) s = RedistStats();
    });
}

#define ELB200_DEFINE_TYPE(SUF, SCALAR, REAL, T)                                                                   \
    static inline AbstractDistMatrix<T>* M_##SUF(ElDistMatrix_##SUF A) { return reinterpret_cast<AbstractDistMatrix<T>*>(A); } \
    static inline const AbstractDistMatrix<T>* CM_##SUF(ElConstDistMatrix_##SUF A) { return reinterpret_cast<const AbstractDistMatrix<T>*>(A); } \
    ElError ElDistMatrixCreateSpecific_##SUF(ElDist U, ElDist V, ElConstGrid g, ElDistMatrix_##SUF* A) {           \
        return Try([&] { *A = reinterpret_cast<ElDistMatrix_##SUF>(new AbstractDistMatrix<T>(*G(g), D(U), D(V))); }); \
    }                                                                                                              \
    ElError ElDistMatrixDestroy_##SUF(ElConstDistMatrix_##SUF A) { return Try([&] { delete CM_##SUF(A); }); }      \
    ElError ElDistMatrixEmpty_##SUF(ElDistMatrix_##SUF A) { return Try([&] { M_##SUF(A)->Empty(); }); }            \
    ElError ElDistMatrixResize_##SUF(ElDistMatrix_##SUF A, ElInt h, ElInt w) { return Try([&] { M_##SUF(A)->Resize(h, w); }); } \
    ElError ElDistMatrixAlign_##SUF(ElDistMatrix_##SUF A, int ca, int ra, bool constrain) {                        \
        return Try([&] { M_##SUF(A)->Align(ca, ra, constrain); });                                                 \
    }                                                                                                              \
    ElError ElDistMatrixAlignWith_##SUF(ElDistMatrix_##SUF A, ElConstDistMatrix_##SUF B) {                         \
        return Try([&] { M_##SUF(A)->AlignWith(*CM_##SUF(B)); });                                                  \
    }                                                                                                              \
    ElError ElDistMatrixAttach_##SUF(ElDistMatrix_##SUF A, ElInt h, ElInt w, ElConstGrid g, int ca, int ra,        \
                                     SCALAR* buf, ElInt ld, int) {                                                 \
        return Try([&] { M_##SUF(A)->Attach(h, w, *G(g), ca, ra, reinterpret_cast<T*>(buf), ld); });               \
    }                                                                                                              \
    ElError ElDistMatrixLockedAttach_##SUF(ElDistMatrix_##SUF A, ElInt h, ElInt w, ElConstGrid g, int ca, int ra,  \
                                           const SCALAR* buf, ElInt ld, int) {                                     \
        return Try([&] { M_##SUF(A)->LockedAttach(h, w, *G(g), ca, ra, reinterpret_cast<const T*>(buf), ld); });   \
    }                                                                                                              \
    ElError ElDistMatrixView_##SUF(ElDistMatrix_##SUF A, ElDistMatrix_##SUF P, ElInt i, ElInt j, ElInt h, ElInt w) { \
        return Try([&] { M_##SUF(A)->ViewOf(*M_##SUF(P), i, j, h, w); });                                          \
    }                                                                                                              \
    ElError ElDistMatrixHeight_##SUF(ElConstDistMatrix_##SUF A, ElInt* v) { return Try([&] { *v = CM_##SUF(A)->Height(); }); } \
    ElError ElDistMatrixWidth_##SUF(ElConstDistMatrix_##SUF A, ElInt* v) { return Try([&] { *v = CM_##SUF(A)->Width(); }); } \
    ElError ElDistMatrixLocalHeight_##SUF(ElConstDistMatrix_##SUF A, ElInt* v) { return Try([&] { *v = CM_##SUF(A)->LocalHeight(); }); } \
    ElError ElDistMatrixLocalWidth_##SUF(ElConstDistMatrix_##SUF A, ElInt* v) { return Try([&] { *v = CM_##SUF(A)->LocalWidth(); }); } \
    ElError ElDistMatrixLDim_##SUF(ElConstDistMatrix_##SUF A, ElInt* v) { return Try([&] { *v = CM_##SUF(A)->LDim(); }); } \
    ElError ElDistMatrixColAlign_##SUF(ElConstDistMatrix_##SUF A, int* v) { return Try([&] { *v = CM_##SUF(A)->ColAlign(); }); } \
    ElError ElDistMatrixRowAlign_##SUF(ElConstDistMatrix_##SUF A, int* v) { return Try([&] { *v = CM_##SUF(A)->RowAlign(); }); } \
    ElError ElDistMatrixColShift_##SUF(ElConstDistMatrix_##SUF A, int* v) { return Try([&] { *v = CM_##SUF(A)->ColShift(); }); } \
    ElError ElDistMatrixRowShift_##SUF(ElConstDistMatrix_##SUF A, int* v) { return Try([&] { *v = CM_##SUF(A)->RowShift(); }); } \
    ElError ElDistMatrixColStride_##SUF(ElConstDistMatrix_##SUF A, int* v) { return Try([&] { *v = CM_##SUF(A)->ColStride(); }); } \
    ElError ElDistMatrixRowStride_##SUF(ElConstDistMatrix_##SUF A, int* v) { return Try([&] { *v = CM_##SUF(A)->RowStride(); }); } \
    ElError ElDistMatrixBuffer_##SUF(ElDistMatrix_##SUF A, SCALAR** b) {                                           \
        return Try([&] { *b = reinterpret_cast<SCALAR*>(M_##SUF(A)->Buffer()); });                                 \
    }                                                                                                              \
    ElError ElDistMatrixLockedBuffer_##SUF(ElConstDistMatrix_##SUF A, const SCALAR** b) {                          \
        return Try([&] { *b = reinterpret_cast<const SCALAR*>(CM_##SUF(A)->LockedBuffer()); });                    \
    }                                                                                                              \
    ElError ElDistMatrixLocalToHost_##SUF(ElConstDistMatrix_##SUF A, SCALAR* host, ElInt ld) {                     \
        return Try([&] { CM_##SUF(A)->LockedMatrix().ToHost(reinterpret_cast<T*>(host), ld); });                   \
    }                                                                                                              \
    ElError ElDistMatrixLocalFromHost_##SUF(ElDistMatrix_##SUF A, const SCALAR* host, ElInt ld) {                  \
        return Try([&] { M_##SUF(A)->Matrix().FromHost(reinterpret_cast<const T*>(host), ld); });                  \
    }                                                                                                              \
    ElError ElDistMatrixHashFill_##SUF(ElDistMatrix_##SUF A, int kind, uint64_t seed, double diag) {               \
        return Try([&] { HashFill(*M_##SUF(A), kind, seed, diag); });                                              \
    }                                                                                                              \
    ElError ElCopyDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) { return Try([&] { Copy(*CM_##SUF(A), *M_##SUF(B)); }); } \
    ElError ElTransposeDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) { return Try([&] { Transpose(*CM_##SUF(A), *M_##SUF(B), false); }); } \
    ElError ElAdjointDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) { return Try([&] { Adjoint(*CM_##SUF(A), *M_##SUF(B)); }); } \
    ElError ElAxpyDist_##SUF(SCALAR alpha, ElConstDistMatrix_##SUF X, ElDistMatrix_##SUF Y) {                      \
        return Try([&] { Axpy(Sc<T, SCALAR>(alpha), *CM_##SUF(X), *M_##SUF(Y)); });                                \
    }                                                                                                              \
    ElError ElAxpyContractDist_##SUF(SCALAR alpha, ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) {              \
        return Try([&] { AxpyContract(Sc<T, SCALAR>(alpha), *CM_##SUF(A), *M_##SUF(B)); });                        \
    }                                                                                                              \
    ElError ElContractDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) { return Try([&] { Contract(*CM_##SUF(A), *M_##SUF(B)); }); } \
    ElError ElScaleDist_##SUF(SCALAR alpha, ElDistMatrix_##SUF A) { return Try([&] { Scale(Sc<T, SCALAR>(alpha), *M_##SUF(A)); }); } \
    ElError ElZeroDist_##SUF(ElDistMatrix_##SUF A) { return Try([&] { Zero(*M_##SUF(A)); }); }                     \
    ElError ElScaleTrapezoidDist_##SUF(SCALAR alpha, ElUpperOrLower uplo, ElDistMatrix_##SUF A, ElInt offset) {    \
        return Try([&] { ScaleTrapezoid(Sc<T, SCALAR>(alpha), UL(uplo), *M_##SUF(A), offset); });                  \
    }                                                                                                              \
    ElError ElReadBinaryFlatDist_##SUF(ElDistMatrix_##SUF A, ElInt height, ElInt width, const char* filename) {    \
        return Try([&] { read::BinaryFlat(*M_##SUF(A), height, width, filename); });                               \
    }                                                                                                              \
    ElError ElReadBinaryDist_##SUF(ElDistMatrix_##SUF A, const char* filename) {                                   \
        return Try([&] { read::Binary(*M_##SUF(A), filename); });                                                  \
    }                                                                                                              \
    ElError ElWriteDist_##SUF(ElConstDistMatrix_##SUF A, const char* basename, int format) {                       \
        return Try([&] {                                                                                           \
            if (format == BINARY) write::Binary(*CM_##SUF(A), basename);                                           \
            else if (format == BINARY_FLAT) write::BinaryFlat(*CM_##SUF(A), basename);                             \
            else LogicError("Only the BINARY and BINARY_FLAT file formats are on this path");                      \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElAxpyTrapezoidDist_##SUF(ElUpperOrLower uplo, SCALAR alpha, ElConstDistMatrix_##SUF X,                \
                                      ElDistMatrix_##SUF Y, ElInt offset) {                                        \
        return Try([&] { AxpyTrapezoid(UL(uplo), Sc<T, SCALAR>(alpha), *CM_##SUF(X), *M_##SUF(Y), offset); });     \
    }                                                                                                              \
    ElError ElMakeTrapezoidalDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A, ElInt offset) {                 \
        return Try([&] { MakeTrapezoidal(UL(uplo), *M_##SUF(A), offset); });                                       \
    }                                                                                                              \
    ElError ElFrobeniusNormDist_##SUF(ElConstDistMatrix_##SUF A, REAL* norm) { return Try([&] { *norm = FrobeniusNorm(*CM_##SUF(A)); }); } \
    ElError ElMaxNormDist_##SUF(ElConstDistMatrix_##SUF A, REAL* norm) { return Try([&] { *norm = MaxNorm(*CM_##SUF(A)); }); } \
    ElError ElGemmDist_##SUF(ElOrientation oA, ElOrientation oB, SCALAR alpha, ElConstDistMatrix_##SUF A,          \
                             ElConstDistMatrix_##SUF B, SCALAR beta, ElDistMatrix_##SUF C) {                       \
        return Try([&] { Gemm(O(oA), O(oB), Sc<T, SCALAR>(alpha), *CM_##SUF(A), *CM_##SUF(B), Sc<T, SCALAR>(beta), *M_##SUF(C)); }); \
    }                                                                                                              \
    ElError ElGemmXDist_##SUF(ElOrientation oA, ElOrientation oB, SCALAR alpha, ElConstDistMatrix_##SUF A,         \
                              ElConstDistMatrix_##SUF B, SCALAR beta, ElDistMatrix_##SUF C, ElGemmAlgorithm alg) { \
        return Try([&] {                                                                                           \
            Gemm(O(oA), O(oB), Sc<T, SCALAR>(alpha), *CM_##SUF(A), *CM_##SUF(B), Sc<T, SCALAR>(beta), *M_##SUF(C), \
                 static_cast<GemmAlgorithm>(alg));                                                                 \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElGemmDistHost_##SUF(ElOrientation oA, ElOrientation oB, SCALAR alpha, ElConstGrid grid, ElInt m, ElInt n, \
                                 ElInt k, const SCALAR* A, ElInt lda, const SCALAR* B, ElInt ldb, SCALAR beta,     \
                                 SCALAR* C, ElInt ldc, ElGemmAlgorithm alg) {                                      \
        return Try([&] {                                                                                           \
            HostLocalMatrix<T> hA{O(oA) == NORMAL ? m : k, O(oA) == NORMAL ? k : m,                                \
                                  const_cast<T*>(reinterpret_cast<const T*>(A)), lda};                             \
            HostLocalMatrix<T> hB{O(oB) == NORMAL ? k : n, O(oB) == NORMAL ? n : k,                                \
                                  const_cast<T*>(reinterpret_cast<const T*>(B)), ldb};                             \
            HostLocalMatrix<T> hC{m, n, reinterpret_cast<T*>(C), ldc};                                             \
            GemmHost(O(oA), O(oB), Sc<T, SCALAR>(alpha), *G(grid), hA, hB, Sc<T, SCALAR>(beta), hC,               \
                     static_cast<GemmAlgorithm>(alg));                                                             \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElSyrkDist_##SUF(ElUpperOrLower uplo, ElOrientation o, SCALAR alpha, ElConstDistMatrix_##SUF A,        \
                             SCALAR beta, ElDistMatrix_##SUF C) {                                                  \
        return Try([&] { Syrk(UL(uplo), O(o), Sc<T, SCALAR>(alpha), *CM_##SUF(A), Sc<T, SCALAR>(beta), *M_##SUF(C), false); }); \
    }                                                                                                              \
    ElError ElHerkDist_##SUF(ElUpperOrLower uplo, ElOrientation o, REAL alpha, ElConstDistMatrix_##SUF A,          \
                             REAL beta, ElDistMatrix_##SUF C) {                                                    \
        return Try([&] { Herk(UL(uplo), O(o), alpha, *CM_##SUF(A), beta, *M_##SUF(C)); });                         \
    }                                                                                                              \
    ElError ElTrrkDist_##SUF(ElUpperOrLower uplo, ElOrientation oA, ElOrientation oB, SCALAR alpha,                \
                             ElConstDistMatrix_##SUF A, ElConstDistMatrix_##SUF B, SCALAR beta,                    \
                             ElDistMatrix_##SUF C) {                                                               \
        return Try([&] {                                                                                           \
            Trrk(UL(uplo), O(oA), O(oB), Sc<T, SCALAR>(alpha), *CM_##SUF(A), *CM_##SUF(B), Sc<T, SCALAR>(beta), *M_##SUF(C)); \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElTrsmDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, ElOrientation o, ElUnitOrNonUnit diag,       \
                             SCALAR alpha, ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) {                      \
        return Try([&] {                                                                                           \
            Trsm(static_cast<LeftOrRight>(side), UL(uplo), O(o), static_cast<UnitOrNonUnit>(diag),                 \
                 Sc<T, SCALAR>(alpha), *CM_##SUF(A), *M_##SUF(B));                                                 \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElTrsmXDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, ElOrientation o, ElUnitOrNonUnit diag,      \
                              SCALAR alpha, ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B, bool checkIfSingular, \
                              ElTrsmAlgorithm alg) {                                                               \
        return Try([&] {                                                                                           \
            Trsm(static_cast<LeftOrRight>(side), UL(uplo), O(o), static_cast<UnitOrNonUnit>(diag),                 \
                 Sc<T, SCALAR>(alpha), *CM_##SUF(A), *M_##SUF(B), checkIfSingular, static_cast<TrsmAlgorithm>(alg)); \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElTrsvDist_##SUF(ElUpperOrLower uplo, ElOrientation o, ElUnitOrNonUnit diag, ElConstDistMatrix_##SUF A, \
                             ElDistMatrix_##SUF x) {                                                               \
        return Try([&] { Trsv(UL(uplo), O(o), static_cast<UnitOrNonUnit>(diag), *CM_##SUF(A), *M_##SUF(x)); });    \
    }                                                                                                              \
    ElError ElSymmDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, SCALAR alpha, ElConstDistMatrix_##SUF A,     \
                             ElConstDistMatrix_##SUF B, SCALAR beta, ElDistMatrix_##SUF C) {                       \
        return Try([&] {                                                                                           \
            Symm(static_cast<LeftOrRight>(side), UL(uplo), Sc<T, SCALAR>(alpha), *CM_##SUF(A), *CM_##SUF(B),       \
                 Sc<T, SCALAR>(beta), *M_##SUF(C), false);                                                         \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElSyr2kDist_##SUF(ElUpperOrLower uplo, ElOrientation o, SCALAR alpha, ElConstDistMatrix_##SUF A,       \
                              ElConstDistMatrix_##SUF B, SCALAR beta, ElDistMatrix_##SUF C) {                      \
        return Try([&] {                                                                                           \
            Syr2k(UL(uplo), O(o), Sc<T, SCALAR>(alpha), *CM_##SUF(A), *CM_##SUF(B), Sc<T, SCALAR>(beta),           \
                  *M_##SUF(C), false);                                                                             \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElTwoSidedTrsmDist_##SUF(ElUpperOrLower uplo, ElUnitOrNonUnit diag, ElDistMatrix_##SUF A,              \
                                     ElConstDistMatrix_##SUF B) {                                                  \
        return Try([&] { TwoSidedTrsm(UL(uplo), static_cast<UnitOrNonUnit>(diag), *M_##SUF(A), *CM_##SUF(B)); });  \
    }                                                                                                              \
    ElError ElTwoSidedTrmmDist_##SUF(ElUpperOrLower uplo, ElUnitOrNonUnit diag, ElDistMatrix_##SUF A,              \
                                     ElConstDistMatrix_##SUF B) {                                                  \
        return Try([&] { TwoSidedTrmm(UL(uplo), static_cast<UnitOrNonUnit>(diag), *M_##SUF(A), *CM_##SUF(B)); });  \
    }                                                                                                              \
    ElError ElTrr2kDist_##SUF(ElUpperOrLower uplo, ElOrientation oA, ElOrientation oB, ElOrientation oC,           \
                              ElOrientation oD, SCALAR alpha, ElConstDistMatrix_##SUF A, ElConstDistMatrix_##SUF B, \
                              SCALAR beta, ElConstDistMatrix_##SUF C, ElConstDistMatrix_##SUF D, SCALAR gamma,     \
                              ElDistMatrix_##SUF E) {                                                              \
        return Try([&] {                                                                                           \
            Trr2k(UL(uplo), O(oA), O(oB), O(oC), O(oD), Sc<T, SCALAR>(alpha), *CM_##SUF(A), *CM_##SUF(B),          \
                  Sc<T, SCALAR>(beta), *CM_##SUF(C), *CM_##SUF(D), Sc<T, SCALAR>(gamma), *M_##SUF(E));             \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElTrmmDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, ElOrientation o, ElUnitOrNonUnit diag,       \
                             SCALAR alpha, ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) {                      \
        return Try([&] {                                                                                           \
            Trmm(static_cast<LeftOrRight>(side), UL(uplo), O(o), static_cast<UnitOrNonUnit>(diag),                 \
                 Sc<T, SCALAR>(alpha), *CM_##SUF(A), *M_##SUF(B));                                                 \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElCholeskyDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A) { return Try([&] { Cholesky(UL(uplo), *M_##SUF(A)); }); } \
    ElError ElReverseCholeskyDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A) {                               \
        return Try([&] { ReverseCholesky(UL(uplo), *M_##SUF(A)); });                                               \
    }                                                                                                              \
    ElError ElCholeskyVariant2Dist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A) {                              \
        return Try([&] {                                                                                           \
            if (UL(uplo) == LOWER) cholesky::LowerVariant2Blocked(*M_##SUF(A));                                    \
            else cholesky::UpperVariant2Blocked(*M_##SUF(A));                                                      \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElCholeskySolveAfterDist_##SUF(ElUpperOrLower uplo, ElOrientation o, ElConstDistMatrix_##SUF A,        \
                                           ElDistMatrix_##SUF B) {                                                 \
        return Try([&] { cholesky::SolveAfter(UL(uplo), O(o), *CM_##SUF(A), *M_##SUF(B)); });                      \
    }                                                                                                              \
    ElError ElHPDSolveDist_##SUF(ElUpperOrLower uplo, ElOrientation o, ElConstDistMatrix_##SUF A,                  \
                                 ElDistMatrix_##SUF B) {                                                           \
        return Try([&] { HPDSolve(UL(uplo), O(o), *CM_##SUF(A), *M_##SUF(B)); });                                  \
    }                                                                                                              \
    ElError ElLUDist_##SUF(ElDistMatrix_##SUF A) { return Try([&] { LU(*M_##SUF(A)); }); }                         \
    ElError ElLUPartialPivDist_##SUF(ElDistMatrix_##SUF A, ElDistPermutation P) {                                  \
        return Try([&] { LU(*M_##SUF(A), *PM(P)); });                                                              \
    }                                                                                                              \
    ElError ElSolveAfterLUDist_##SUF(ElOrientation o, ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) {           \
        return Try([&] { lu::SolveAfter(O(o), *CM_##SUF(A), *M_##SUF(B)); });                                      \
    }                                                                                                              \
    ElError ElSolveAfterLUPartialPivDist_##SUF(ElOrientation o, ElConstDistMatrix_##SUF A, ElConstDistPermutation P, \
                                               ElDistMatrix_##SUF B) {                                             \
        return Try([&] { lu::SolveAfter(O(o), *CM_##SUF(A), *CPM(P), *M_##SUF(B)); });                             \
    }                                                                                                              \
    ElError ElCholeskyPivDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF A, ElDistPermutation P) {              \
        return Try([&] { Cholesky(UL(uplo), *M_##SUF(A), *PM(P)); });                                              \
    }                                                                                                              \
    ElError ElSolveAfterCholeskyPivDist_##SUF(ElUpperOrLower uplo, ElOrientation o, ElConstDistMatrix_##SUF A,     \
                                              ElConstDistPermutation P, ElDistMatrix_##SUF B) {                    \
        return Try([&] { cholesky::SolveAfter(UL(uplo), O(o), *CM_##SUF(A), *CPM(P), *M_##SUF(B)); });             \
    }                                                                                                              \
    ElError ElCholeskyModDist_##SUF(ElUpperOrLower uplo, ElDistMatrix_##SUF T_, REAL alpha, ElDistMatrix_##SUF V) { \
        return Try([&] { CholeskyMod(UL(uplo), *M_##SUF(T_), alpha, *M_##SUF(V)); });                              \
    }                                                                                                              \
    ElError ElLinearSolveDist_##SUF(ElConstDistMatrix_##SUF A, ElDistMatrix_##SUF B) {                             \
        return Try([&] { LinearSolve(*CM_##SUF(A), *M_##SUF(B)); });                                               \
    }                                                                                                              \
    ElError ElDistPermutationPermuteRowsDist_##SUF(ElConstDistPermutation P, ElDistMatrix_##SUF A, ElInt off) {    \
        return Try([&] { CPM(P)->PermuteRows(*M_##SUF(A), off); });                                                \
    }                                                                                                              \
    ElError ElDistPermutationInversePermuteRowsDist_##SUF(ElConstDistPermutation P, ElDistMatrix_##SUF A, ElInt off) { \
        return Try([&] { CPM(P)->InversePermuteRows(*M_##SUF(A), off); });                                         \
    }                                                                                                              \
    ElError ElDistPermutationPermuteColsDist_##SUF(ElConstDistPermutation P, ElDistMatrix_##SUF A, ElInt off) {    \
        return Try([&] { CPM(P)->PermuteCols(*M_##SUF(A), off); });                                                \
    }                                                                                                              \
    ElError ElDistPermutationInversePermuteColsDist_##SUF(ElConstDistPermutation P, ElDistMatrix_##SUF A, ElInt off) { \
        return Try([&] { CPM(P)->InversePermuteCols(*M_##SUF(A), off); });                                         \
    }

ELB200_DEFINE_TYPE(s, float, float, float)
ELB200_DEFINE_TYPE(d, double, double, double)
ELB200_DEFINE_TYPE(c, elb200_c32, float, Complex<float>)
ELB200_DEFINE_TYPE(z, elb200_c64, double, Complex<double>)

#define ELB200_DEFINE_HERMITIAN(SUF, SCALAR, REAL, T)                                                              \
    ElError ElHemmDist_##SUF(ElLeftOrRight side, ElUpperOrLower uplo, SCALAR alpha, ElConstDistMatrix_##SUF A,     \
                             ElConstDistMatrix_##SUF B, SCALAR beta, ElDistMatrix_##SUF C) {                       \
        return Try([&] {                                                                                           \
            Hemm(static_cast<LeftOrRight>(side), UL(uplo), Sc<T, SCALAR>(alpha), *CM_##SUF(A), *CM_##SUF(B),       \
                 Sc<T, SCALAR>(beta), *M_##SUF(C));                                                                \
        });                                                                                                        \
    }                                                                                                              \
    ElError ElHer2kDist_##SUF(ElUpperOrLower uplo, ElOrientation o, SCALAR alpha, ElConstDistMatrix_##SUF A,       \
                              ElConstDistMatrix_##SUF B, REAL beta, ElDistMatrix_##SUF C) {                        \
        return Try([&] { Her2k(UL(uplo), O(o), Sc<T, SCALAR>(alpha), *CM_##SUF(A), *CM_##SUF(B), beta, *M_##SUF(C)); }); \
    }
ELB200_DEFINE_HERMITIAN(c, elb200_c32, float, Complex<float>)
ELB200_DEFINE_HERMITIAN(z, elb200_c64, double, Complex<double>)

}  // extern "C"
