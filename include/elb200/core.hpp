// elb200 host layer, part 1: the object model of the reference's hot path --
// Grid, Matrix, DistMatrix, views, alignments, Blocksize() -- with DEVICE-resident
// local matrices and NCCL communicators.  Same names, argument meaning and error
// behaviour as the reference so that code written against
//   include/El/core/Grid.hpp:15-160, include/El/core/Matrix/decl.hpp,
//   include/El/core/DistMatrix/{Abstract,Element}.hpp, include/El/core/View/impl.hpp,
//   include/El/core/environment/decl.hpp:87-94 (blocksize API)
// compiles against this header for the Gemm / Cholesky / HPDSolve path.
//
// Design differences (B200-first, not a port):
//   * one runtime-typed AbstractDistMatrix<T> carries (colDist,rowDist) as values;
//     DistMatrix<T,U,V> is a thin typed shell, so every redistribution `B = A` goes
//     through ONE generic element-cyclic engine (redist.cpp) instead of ~40 pairwise
//     pack/collective/unpack primitives;
//   * local buffers live in HBM (stream-ordered cudaMallocAsync pool); element access
//     from the host is explicit (Get/Set do a synchronous copy) -- there is no host
//     mirror and no CPU fallback;
//   * Grid owns NCCL communicators (world=VC order, row=MR, column=MC) created from a
//     ncclUniqueId that the launcher (torch.distributed / MPI / file) broadcasts.
#pragma once
#include <complex>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

struct ncclComm;
struct CUstream_st;

namespace El {

typedef int Int;
template <typename R> using Complex = std::complex<R>;
template <typename T> struct BaseHelper { typedef T type; };
template <typename R> struct BaseHelper<Complex<R>> { typedef R type; };
template <typename T> using Base = typename BaseHelper<T>::type;
template <typename T> struct IsComplex { static const bool value = false; };
template <typename R> struct IsComplex<Complex<R>> { static const bool value = true; };

// ---- enums: same order/values as include/El/core/types.hpp ----
enum Dist { MC, MD, MR, VC, VR, STAR, CIRC };
enum Orientation { NORMAL, TRANSPOSE, ADJOINT };
enum UpperOrLower { LOWER, UPPER };
enum LeftOrRight { LEFT, RIGHT };
enum UnitOrNonUnit { NON_UNIT, UNIT };
enum GridOrder { ROW_MAJOR, COLUMN_MAJOR };
enum GemmAlgorithm { GEMM_DEFAULT, GEMM_SUMMA_A, GEMM_SUMMA_B, GEMM_SUMMA_C, GEMM_SUMMA_DOT, GEMM_CANNON };
enum TrsmAlgorithm { TRSM_DEFAULT, TRSM_LARGE, TRSM_MEDIUM, TRSM_SMALL };

inline char OrientationToChar(Orientation o) { return o == NORMAL ? 'N' : (o == TRANSPOSE ? 'T' : 'C'); }
inline char UpperOrLowerToChar(UpperOrLower u) { return u == LOWER ? 'L' : 'U'; }
inline char LeftOrRightToChar(LeftOrRight s) { return s == LEFT ? 'L' : 'R'; }
inline char UnitOrNonUnitToChar(UnitOrNonUnit d) { return d == UNIT ? 'U' : 'N'; }

// ---- exceptions (include/El/core/environment/decl.hpp:200-240) ----
struct SingularMatrixException : std::runtime_error {
    explicit SingularMatrixException(const char* m = "Matrix was singular") : std::runtime_error(m) {}
};
struct NonHPDMatrixException : std::runtime_error {
    explicit NonHPDMatrixException(const char* m = "Matrix was not HPD") : std::runtime_error(m) {}
};
[[noreturn]] void LogicError(const std::string& msg);    // throws std::logic_error
[[noreturn]] void RuntimeError(const std::string& msg);  // throws std::runtime_error

// ---- blocksize stack (src/blas_like/blocksizes.cpp:37-73; default 128, environment.cpp:181-183) ----
Int Blocksize();
void SetBlocksize(Int blocksize);
void PushBlocksizeStack(Int blocksize);
void PopBlocksizeStack();
// Edge of the C blocks SUMMA_Dot forms (the reference hard-codes blockSizeDot = 2000, Gemm/NN.hpp:233, a
// CPU-cache size).  0 (default) = sized for HBM: the largest multiple of 128 whose replicated block stays
// within 1 GiB, so an 8192 x 8192 float C is ONE product of 4096 tiles (27.7 waves of 148 SMs) instead of
// 25 products of 1.7 waves each.  The result does not depend on it (every C entry is a full-k product).
Int GemmDotBlocksize(size_t scalarBytes);
void SetGemmDotBlocksize(Int bs);
template <typename T> Int LocalTrrkBlocksize();            // kept for API parity (default 64);
template <typename T> void SetLocalTrrkBlocksize(Int bs);  // the masked-GEMM leaf does not need it

// ---- the layer's CUDA stream (all kernels and NCCL calls are enqueued on it) ----
typedef CUstream_st* Stream;
Stream CurrentStream();
void SetCurrentStream(Stream s);
void SynchronizeStream();
// Overlapped panel loops (SUMMA-C panel prefetch, Cholesky look-ahead) on/off; default on,
// ELB200_OVERLAP=0 in the environment turns them off.  Results are identical either way.
void SetOverlap(bool on);

// ---- index math (include/El/core/indexing/impl.hpp) ----
inline Int Mod(Int a, Int b) { Int r = a % b; return r < 0 ? r + b : r; }
inline Int Shift_(Int rank, Int align, Int stride) { return Mod(rank - align, stride); }
inline Int Length_(Int n, Int shift, Int stride) { return n > shift ? (n - shift - 1) / stride + 1 : 0; }
inline Int Length(Int n, Int rank, Int align, Int stride) { return Length_(n, Shift_(rank, align, stride), stride); }
inline Int MaxLength(Int n, Int stride) { return n > 0 ? (n - 1) / stride + 1 : 0; }

struct Range { Int beg, end; Range(Int b = 0, Int e = 0) : beg(b), end(e) {} };
typedef Range IR;
static const Int END = -1;
struct AllRange {};
static const AllRange ALL = AllRange();

// ---------------------------------------------------------------------------
// Grid (src/core/Grid.cpp:112-247): r x c process grid; ranks in the column
// communicator MC, the row communicator MR, and the vector communicators VC
// (column-major) / VR (row-major).
// ---------------------------------------------------------------------------
struct Comm {
    ncclComm* nccl = nullptr;  // null when size == 1
    int rank = 0;
    int size = 1;
    std::vector<int> toWorld;  // comm rank -> rank in the grid's world NCCL communicator
};

class Grid {
public:
    // single-process 1x1 grid (no NCCL)
    Grid();
    // one process per GPU: `uniqueId` is the 128-byte ncclUniqueId broadcast by the launcher
    Grid(const void* ncclUniqueId128, int worldRank, int worldSize, int height,
         GridOrder order = COLUMN_MAJOR);
    // shape-only grid for planning / index math (no communicators; any collective on it throws)
    struct PlanningOnly {};
    Grid(int height, int width, int mcRank, int mrRank, PlanningOnly);
    ~Grid();
    Grid(const Grid&) = delete;
    Grid& operator=(const Grid&) = delete;

    int Height() const { return height_; }
    int Width() const { return width_; }
    int Size() const { return height_ * width_; }
    int Rank() const { return vcRank_; }
    int Row() const { return mcRank_; }
    int Col() const { return mrRank_; }
    int MCRank() const { return mcRank_; }
    int MRRank() const { return mrRank_; }
    int VCRank() const { return vcRank_; }
    int VRRank() const { return vrRank_; }
    int MCSize() const { return height_; }
    int MRSize() const { return width_; }
    int VCSize() const { return Size(); }
    int VRSize() const { return Size(); }
    GridOrder Order() const { return order_; }
    const Comm& MCComm() const { return mc_; }
    const Comm& MRComm() const { return mr_; }
    const Comm& VCComm() const { return vc_; }
    const Comm& VRComm() const { return vr_; }
    ncclComm* WorldNccl() const { return world_; }
    int WorldRank() const { return worldRank_; }
    // world (NCCL) rank of the process at grid position (mcRank, mrRank)
    int WorldRankOf(int mcRank, int mrRank) const;
    int VCToVR(int vc) const { return (vc / height_) + width_ * (vc % height_); }
    int VRToVC(int vr) const { return (vr / width_) + height_ * (vr % width_); }

    // NVLink peer-memory exchange windows (ELB200_P2P=1): every rank's window is mapped by all its peers
    // (CUDA IPC); redistributions push their pieces into the destination's window instead of going
    // through ncclSend/ncclRecv.  Layout of a window: 4 KB of flags, then per channel (stream) two halves
    // of Size() regions of regionBytes each: half (epoch & 1), region = world rank of the source.
    struct P2PState {
        bool on = false;
        size_t regionBytes = 0;
        char* local = nullptr;
        std::vector<char*> peer;   // by world rank; peer[WorldRank()] == local
        unsigned epoch[2] = {0, 0};
        int* error = nullptr;      // pinned host flag raised by a timed-out wait
        unsigned* Ready(int owner, int ch, int src) const { return (unsigned*)(peer[owner] + ch * 256) + src; }
        unsigned* Ack(int owner, int ch, int src) const { return (unsigned*)(peer[owner] + 1024 + ch * 256) + src; }
        char* Region(int owner, int ch, unsigned ep, int src) const {
            return peer[owner] + 4096 + ((size_t)(ch * 2 + (ep & 1u)) * peer.size() + (size_t)src) * regionBytes;
        }
    };
    P2PState& P2P() const { return p2p_; }

    static const Grid& Default();  // lazily-built trivial 1x1 grid
    static int DefaultHeight(int gridSize);  // src/core/Grid.cpp:66-72

private:
    int height_ = 1, width_ = 1;
    int mcRank_ = 0, mrRank_ = 0, vcRank_ = 0, vrRank_ = 0;
    int worldRank_ = 0;
    GridOrder order_ = COLUMN_MAJOR;
    ncclComm* world_ = nullptr;
    Comm mc_, mr_, vc_, vr_;
    mutable P2PState p2p_;
    void SetupP2P(int worldSize);
};

// stride / rank of a distribution on a grid
int DistStride(Dist d, const Grid& g);
int DistRank(Dist d, const Grid& g);                       // of this process
int DistRankOf(Dist d, const Grid& g, int mcRank, int mrRank);
Dist PartialDist(Dist d);       // Partial(VC)=MC, Partial(VR)=MR, else identity
Dist PartialUnionDist(Dist d);  // VC->MR, VR->MC, else STAR
const char* DistName(Dist d);

// ---------------------------------------------------------------------------
// Matrix<T>: local column-major matrix in device memory (Matrix/decl.hpp:175-181)
// ---------------------------------------------------------------------------
template <typename T>
class Matrix {
public:
    Matrix();
    Matrix(Int height, Int width);
    Matrix(Int height, Int width, Int ldim);
    Matrix(const Matrix<T>& A);  // deep copy (device to device)
    Matrix(Matrix<T>&& A) noexcept;
    ~Matrix();
    Matrix<T>& operator=(const Matrix<T>& A);  // deep copy
    Matrix<T>& operator=(Matrix<T>&& A) noexcept;

    void Empty(bool freeMemory = true);
    void Resize(Int height, Int width);
    void Resize(Int height, Int width, Int ldim);
    void Attach(Int height, Int width, T* deviceBuffer, Int ldim);
    void LockedAttach(Int height, Int width, const T* deviceBuffer, Int ldim);

    Int Height() const { return height_; }
    Int Width() const { return width_; }
    Int LDim() const { return ldim_; }
    bool Viewing() const { return !owner_; }
    bool Locked() const { return locked_; }
    T* Buffer();
    T* Buffer(Int i, Int j);
    const T* LockedBuffer() const { return data_; }
    const T* LockedBuffer(Int i, Int j) const { return data_ + size_t(i) + size_t(j) * size_t(ldim_); }

    Matrix<T> operator()(Range I, Range J);
    const Matrix<T> operator()(Range I, Range J) const;

    // host <-> device element access (synchronous; tests and small fix-ups only)
    T Get(Int i, Int j) const;
    void Set(Int i, Int j, T value);
    // whole-matrix transfers to/from a host column-major buffer (synchronous)
    void ToHost(T* host, Int hostLDim) const;
    void FromHost(const T* host, Int hostLDim);

private:
    Int height_ = 0, width_ = 0, ldim_ = 1;
    T* data_ = nullptr;
    size_t capacity_ = 0;  // elements, when owner
    bool owner_ = true, locked_ = false;
    void Release();
};

// ---------------------------------------------------------------------------
// AbstractDistMatrix<T>: element-cyclic distributed matrix with runtime (U,V)
// ---------------------------------------------------------------------------
template <typename T>
class AbstractDistMatrix {
public:
    AbstractDistMatrix(const El::Grid& g, Dist colDist, Dist rowDist);
    AbstractDistMatrix(const AbstractDistMatrix<T>& A);  // deep copy, same distribution
    AbstractDistMatrix(AbstractDistMatrix<T>&& A) noexcept;
    virtual ~AbstractDistMatrix();
    AbstractDistMatrix<T>& operator=(const AbstractDistMatrix<T>& A);  // redistributes: *this = A
    AbstractDistMatrix<T>& operator=(AbstractDistMatrix<T>&& A) noexcept;

    // --- size and distribution queries (DistMatrix/Abstract.hpp:88-129) ---
    const El::Grid& Grid() const { return *grid_; }
    Dist ColDist() const { return colDist_; }
    Dist RowDist() const { return rowDist_; }
    Int Height() const { return height_; }
    Int Width() const { return width_; }
    Int LocalHeight() const { return matrix_.Height(); }
    Int LocalWidth() const { return matrix_.Width(); }
    Int LDim() const { return matrix_.LDim(); }
    int ColAlign() const { return colAlign_; }
    int RowAlign() const { return rowAlign_; }
    int ColShift() const { return colShift_; }
    int RowShift() const { return rowShift_; }
    int ColStride() const { return DistStride(colDist_, *grid_); }
    int RowStride() const { return DistStride(rowDist_, *grid_); }
    int ColRank() const { return DistRank(colDist_, *grid_); }
    int RowRank() const { return DistRank(rowDist_, *grid_); }
    bool ColConstrained() const { return colConstrained_; }
    bool RowConstrained() const { return rowConstrained_; }
    bool Viewing() const { return viewing_; }
    bool Locked() const { return locked_; }
    bool Participating() const { return true; }
    El::Matrix<T>& Matrix() { return matrix_; }
    const El::Matrix<T>& LockedMatrix() const { return matrix_; }
    T* Buffer() { return matrix_.Buffer(); }
    const T* LockedBuffer() const { return matrix_.LockedBuffer(); }

    // --- global <-> local index math (src/core/DistMatrix/Element.cpp:544-600) ---
    int RowOwner(Int i) const { return Mod(i + colAlign_, ColStride()); }
    int ColOwner(Int j) const { return Mod(j + rowAlign_, RowStride()); }
    Int LocalRowOffset(Int i) const { return Length_(i, colShift_, ColStride()); }
    Int LocalColOffset(Int j) const { return Length_(j, rowShift_, RowStride()); }
    Int GlobalRow(Int iLoc) const { return colShift_ + iLoc * ColStride(); }
    Int GlobalCol(Int jLoc) const { return rowShift_ + jLoc * RowStride(); }
    bool IsLocalRow(Int i) const { return RowOwner(i) == ColRank(); }
    bool IsLocalCol(Int j) const { return ColOwner(j) == RowRank(); }
    bool IsLocal(Int i, Int j) const { return IsLocalRow(i) && IsLocalCol(j); }

    // --- alignment and sizing (DistMatrix/Element.hpp:51-88) ---
    void Empty(bool freeMemory = true);
    void Resize(Int height, Int width);
    void Resize(Int height, Int width, Int ldim);
    void Align(int colAlign, int rowAlign, bool constrain = true);
    void AlignCols(int colAlign, bool constrain = true);
    void AlignRows(int rowAlign, bool constrain = true);
    void FreeAlignments();
    void AlignWith(const AbstractDistMatrix<T>& A, bool constrain = true, bool allowMismatch = false);
    void AlignColsWith(const AbstractDistMatrix<T>& A, bool constrain = true, bool allowMismatch = false);
    void AlignRowsWith(const AbstractDistMatrix<T>& A, bool constrain = true, bool allowMismatch = false);
    void AlignAndResize(int colAlign, int rowAlign, Int height, Int width, bool force = false,
                        bool constrain = false);
    // attach to an existing DEVICE buffer holding the local matrix
    void Attach(Int height, Int width, const El::Grid& g, int colAlign, int rowAlign, T* buffer, Int ldim);
    void LockedAttach(Int height, Int width, const El::Grid& g, int colAlign, int rowAlign,
                      const T* buffer, Int ldim);

    // --- views (include/El/core/View/impl.hpp:394-428) ---
    // *this becomes a view of A(i:i+height, j:j+width)
    void ViewOf(AbstractDistMatrix<T>& A, Int i, Int j, Int height, Int width);
    void LockedViewOf(const AbstractDistMatrix<T>& A, Int i, Int j, Int height, Int width);

    // --- local element access (synchronous host<->device; tests only) ---
    T GetLocal(Int iLoc, Int jLoc) const { return matrix_.Get(iLoc, jLoc); }
    void SetLocal(Int iLoc, Int jLoc, T v) { matrix_.Set(iLoc, jLoc, v); }

protected:
    const El::Grid* grid_;
    Dist colDist_, rowDist_;
    Int height_ = 0, width_ = 0;
    int colAlign_ = 0, rowAlign_ = 0, colShift_ = 0, rowShift_ = 0;
    bool colConstrained_ = false, rowConstrained_ = false;
    bool viewing_ = false, locked_ = false;
    El::Matrix<T> matrix_;
    void SetShifts();
    template <typename S> friend class AbstractDistMatrix;
};
template <typename T> using ElementalMatrix = AbstractDistMatrix<T>;

// Typed shell: DistMatrix<double,MC,MR> etc. (include/El/core/DistMatrix/Element/*.hpp)
template <typename T, Dist U = MC, Dist V = MR>
class DistMatrix : public AbstractDistMatrix<T> {
public:
    typedef AbstractDistMatrix<T> base;
    explicit DistMatrix(const El::Grid& g = El::Grid::Default()) : base(g, U, V) {}
    DistMatrix(Int height, Int width, const El::Grid& g = El::Grid::Default()) : base(g, U, V) {
        this->Resize(height, width);
    }
    DistMatrix(const DistMatrix<T, U, V>& A) : base(A) {}
    DistMatrix(DistMatrix<T, U, V>&& A) noexcept : base(std::move(A)) {}
    // redistributing constructor / assignment from any distribution
    DistMatrix(const AbstractDistMatrix<T>& A) : base(A.Grid(), U, V) { base::operator=(A); }
    DistMatrix<T, U, V>& operator=(const AbstractDistMatrix<T>& A) { base::operator=(A); return *this; }
    DistMatrix<T, U, V>& operator=(const DistMatrix<T, U, V>& A) { base::operator=(A); return *this; }
    DistMatrix<T, U, V>& operator=(DistMatrix<T, U, V>&& A) noexcept { base::operator=(std::move(A)); return *this; }

    DistMatrix<T, U, V> operator()(Range I, Range J) {
        DistMatrix<T, U, V> V_(this->Grid());
        V_.ViewOf(*this, I.beg, J.beg, ClampEnd(I.end, this->Height()) - I.beg, ClampEnd(J.end, this->Width()) - J.beg);
        return V_;
    }
    DistMatrix<T, U, V> operator()(AllRange, Range J) { return (*this)(Range(0, this->Height()), J); }
    DistMatrix<T, U, V> operator()(Range I, AllRange) { return (*this)(I, Range(0, this->Width())); }
    const DistMatrix<T, U, V> operator()(Range I, Range J) const {
        DistMatrix<T, U, V> V_(this->Grid());
        V_.LockedViewOf(*this, I.beg, J.beg, ClampEnd(I.end, this->Height()) - I.beg, ClampEnd(J.end, this->Width()) - J.beg);
        return V_;
    }
    const DistMatrix<T, U, V> operator()(AllRange, Range J) const { return (*this)(Range(0, this->Height()), J); }
    const DistMatrix<T, U, V> operator()(Range I, AllRange) const { return (*this)(I, Range(0, this->Width())); }

private:
    static Int ClampEnd(Int e, Int n) { return (e == END || e > n) ? n : e; }
};

// ---------------------------------------------------------------------------
// Redistribution and level-1 helpers on the hot path (level1/Copy.hpp, Transpose.hpp,
// Contract.hpp, AxpyContract.hpp, TransposeAxpyContract.hpp, ScaleTrapezoid.hpp, ...)
// ---------------------------------------------------------------------------
template <typename T> void Copy(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B);
template <typename T> void Transpose(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B, bool conjugate = false);
template <typename T> void Adjoint(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B);
// B += alpha * (sum over the ranks holding partial sums of A), A partially replicated
template <typename T> void AxpyContract(T alpha, const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B);
template <typename T> void Contract(const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B);
template <typename T> void TransposeAxpyContract(T alpha, const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B, bool conjugate = false);
// B += alpha A (any pair of distributions)
template <typename T> void Axpy(T alpha, const AbstractDistMatrix<T>& A, AbstractDistMatrix<T>& B);
template <typename T> void Scale(T alpha, AbstractDistMatrix<T>& A);
template <typename T> void Scale(T alpha, Matrix<T>& A);
template <typename T> void Zero(AbstractDistMatrix<T>& A);
template <typename T> void Zero(Matrix<T>& A);
template <typename T> void Zeros(AbstractDistMatrix<T>& A, Int m, Int n);
template <typename T> void Conjugate(AbstractDistMatrix<T>& A);
template <typename T> void ScaleTrapezoid(T alpha, UpperOrLower uplo, AbstractDistMatrix<T>& A, Int offset = 0);
template <typename T> void MakeTrapezoidal(UpperOrLower uplo, AbstractDistMatrix<T>& A, Int offset = 0);
// Y_trap += alpha X_trap (include/El/blas_like/level1/AxpyTrapezoid.hpp:14-49,73-160)
template <typename T> void AxpyTrapezoid(UpperOrLower uplo, T alpha, const AbstractDistMatrix<T>& X, AbstractDistMatrix<T>& Y, Int offset = 0);
template <typename T> void LocalAxpyTrapezoid(UpperOrLower uplo, T alpha, const AbstractDistMatrix<T>& X, AbstractDistMatrix<T>& Y, Int offset = 0);
template <typename T> void Copy(const Matrix<T>& A, Matrix<T>& B);
// grid-independent counter-hash fill on global indices (kind 0 general, 1 Hermitian + diag)
template <typename T> void HashFill(AbstractDistMatrix<T>& A, int kind, uint64_t seed, double diag = 0.0);
template <typename T> Base<T> FrobeniusNorm(const AbstractDistMatrix<T>& A);
template <typename T> Base<T> MaxNorm(const AbstractDistMatrix<T>& A);

// Collective statistics of the redistribution engine (for tests and the bench breakdown)
struct RedistStats {
    uint64_t copies = 0, messages = 0, bytesSent = 0, packLaunches = 0, zeroCopySends = 0,
             reduceScatters = 0, allGathers = 0, p2pPushes = 0;  // p2pPushes: messages written into a peer's window
};
RedistStats& GetRedistStats();

}  // namespace El
