// Object model of the host layer: runtime state, Grid (NCCL communicators),
// device-resident Matrix and the runtime-typed element-cyclic DistMatrix.
// See include/elb200/core.hpp for the reference interfaces each piece mirrors.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "dev.hpp"

namespace El {

void LogicError(const std::string& msg) { throw std::logic_error(msg); }
void RuntimeError(const std::string& msg) { throw std::runtime_error(msg); }

// ---- blocksize stack (src/blas_like/blocksizes.cpp:16,37-73; default pushed at init) ----
namespace {
std::vector<Int>& BlocksizeStack() {
    static std::vector<Int> stack(1, 128);
    return stack;
}
Int g_localTrrkFloat = 64, g_localTrrkDouble = 64, g_localTrrkCFloat = 64, g_localTrrkCDouble = 64;
}  // namespace
Int Blocksize() { return BlocksizeStack().back(); }
namespace { Int g_dotBlocksize = 0; }
void SetGemmDotBlocksize(Int bs) {
    if (bs < 0) LogicError("Dot blocksize must be non-negative");
    g_dotBlocksize = bs;
}
Int GemmDotBlocksize(size_t scalarBytes) {
    if (g_dotBlocksize > 0) return g_dotBlocksize;
    const double edge = std::sqrt(double(size_t(1) << 30) / double(scalarBytes));
    return std::max<Int>(128, Int(edge) / 128 * 128);
}
void SetBlocksize(Int b) {
    if (b <= 0) LogicError("Blocksize must be positive");
    BlocksizeStack().back() = b;
}
void PushBlocksizeStack(Int b) {
    if (b <= 0) LogicError("Blocksize must be positive");
    BlocksizeStack().push_back(b);
}
void PopBlocksizeStack() {
    if (BlocksizeStack().size() <= 1) LogicError("Blocksize stack would become empty");
    BlocksizeStack().pop_back();
}
template <> Int LocalTrrkBlocksize<float>() { return g_localTrrkFloat; }
template <> Int LocalTrrkBlocksize<double>() { return g_localTrrkDouble; }
template <> Int LocalTrrkBlocksize<Complex<float>>() { return g_localTrrkCFloat; }
template <> Int LocalTrrkBlocksize<Complex<double>>() { return g_localTrrkCDouble; }
template <> void SetLocalTrrkBlocksize<float>(Int b) { g_localTrrkFloat = b; }
template <> void SetLocalTrrkBlocksize<double>(Int b) { g_localTrrkDouble = b; }
template <> void SetLocalTrrkBlocksize<Complex<float>>(Int b) { g_localTrrkCFloat = b; }
template <> void SetLocalTrrkBlocksize<Complex<double>>(Int b) { g_localTrrkCDouble = b; }

namespace dev {
int PanelSms(int dflt) {
    static int v = -2;
    if (v == -2) {
        const char* e = std::getenv("ELB200_PANEL_SMS");
        v = e ? std::atoi(e) : -1;
    }
    return v >= 0 ? v : dflt;
}
bool PhaseTimer::Enabled() {
    static int v = -1;
    if (v < 0) { const char* e = std::getenv("ELB200_TRACE"); v = (e && std::atoi(e) != 0) ? 1 : 0; }
    return v != 0;
}
PhaseTimer::PhaseTimer() {
    if (!Enabled()) return;
    ELB_CUDA(cudaEventCreate(&a));
    ELB_CUDA(cudaEventCreate(&b));
}
PhaseTimer::~PhaseTimer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
}
void PhaseTimer::Begin(cudaStream_t s) {
    if (!a) return;
    ELB_CUDA(cudaStreamSynchronize(s));
    ELB_CUDA(cudaEventRecord(a, s));
}
void PhaseTimer::End(cudaStream_t s, const char* name) {
    if (!a) return;
    ELB_CUDA(cudaEventRecord(b, s));
    ELB_CUDA(cudaEventSynchronize(b));
    float ms = 0.f;
    ELB_CUDA(cudaEventElapsedTime(&ms, a, b));
    for (auto& kv : acc)
        if (kv.first == name) { kv.second += ms; return; }
    acc.emplace_back(name, (double)ms);
}
void PhaseTimer::Report(const char* title) {
    if (!a) return;
    double tot = 0;
    for (auto& kv : acc) tot += kv.second;
    std::fprintf(stderr, "[elb200 trace] %s: total %.2f ms\n", title, tot);
    for (auto& kv : acc) std::fprintf(stderr, "[elb200 trace]   %-28s %10.2f ms  %5.1f%%\n", kv.first.c_str(), kv.second, 100.0 * kv.second / tot);
}
static int g_overlap = -1;
bool OverlapEnabled() {
    if (g_overlap < 0) {
        const char* e = std::getenv("ELB200_OVERLAP");
        g_overlap = (e && std::atoi(e) == 0) ? 0 : 1;
    }
    return g_overlap != 0;
}
}  // namespace dev
void SetOverlap(bool on) { dev::g_overlap = on ? 1 : 0; }

Stream CurrentStream() { return (Stream)elb200::current_stream(); }
void SetCurrentStream(Stream s) { elb200::set_current_stream((cudaStream_t)s); }
void SynchronizeStream() { ELB_CUDA(cudaStreamSynchronize(dev::stream())); }

// ---------------------------------------------------------------------------
// Grid
// ---------------------------------------------------------------------------
Grid::Grid() {
    mc_.toWorld = mr_.toWorld = vc_.toWorld = vr_.toWorld = std::vector<int>(1, 0);
}

Grid::Grid(int height, int width, int mcRank, int mrRank, PlanningOnly) {
    if (height <= 0 || width <= 0 || mcRank < 0 || mcRank >= height || mrRank < 0 || mrRank >= width)
        LogicError("Invalid planning grid");
    height_ = height; width_ = width; mcRank_ = mcRank; mrRank_ = mrRank;
    vcRank_ = mcRank_ + height_ * mrRank_;
    vrRank_ = mrRank_ + width_ * mcRank_;
    worldRank_ = vcRank_;
    mc_.rank = mcRank_; mc_.size = height_;
    mr_.rank = mrRank_; mr_.size = width_;
    vc_.rank = vcRank_; vc_.size = height_ * width_;
    vr_.rank = vrRank_; vr_.size = height_ * width_;
}

int Grid::WorldRankOf(int mcRank, int mrRank) const {
    return order_ == COLUMN_MAJOR ? mcRank + height_ * mrRank : mrRank + width_ * mcRank;
}

int Grid::DefaultHeight(int gridSize) {
    // largest divisor of gridSize not exceeding sqrt(gridSize) (Grid.cpp:66-72)
    int h = 1;
    for (int d = 1; d * d <= gridSize; ++d)
        if (gridSize % d == 0) h = d;
    return h;
}

Grid::Grid(const void* uid, int worldRank, int worldSize, int height, GridOrder order) {
    if (worldSize <= 0 || worldRank < 0 || worldRank >= worldSize) LogicError("Invalid rank/size for Grid");
    if (height <= 0) height = DefaultHeight(worldSize);
    if (worldSize % height != 0) LogicError("Grid height does not evenly divide the number of processes");
    height_ = height;
    width_ = worldSize / height;
    order_ = order;
    worldRank_ = worldRank;
    if (order == COLUMN_MAJOR) { mcRank_ = worldRank % height_; mrRank_ = worldRank / height_; }
    else { mcRank_ = worldRank / width_; mrRank_ = worldRank % width_; }
    vcRank_ = mcRank_ + height_ * mrRank_;
    vrRank_ = mrRank_ + width_ * mcRank_;

    mc_.rank = mcRank_; mc_.size = height_;
    for (int i = 0; i < height_; ++i) mc_.toWorld.push_back(WorldRankOf(i, mrRank_));
    mr_.rank = mrRank_; mr_.size = width_;
    for (int j = 0; j < width_; ++j) mr_.toWorld.push_back(WorldRankOf(mcRank_, j));
    vc_.rank = vcRank_; vc_.size = worldSize;
    for (int v = 0; v < worldSize; ++v) vc_.toWorld.push_back(WorldRankOf(v % height_, v / height_));
    vr_.rank = vrRank_; vr_.size = worldSize;
    for (int v = 0; v < worldSize; ++v) vr_.toWorld.push_back(WorldRankOf(v / width_, v % width_));

    if (worldSize > 1) {
        if (!uid) LogicError("A multi-process Grid needs the broadcast ncclUniqueId");
        ncclUniqueId id;
        std::memcpy(&id, uid, sizeof(id));
        // The panel gathers run beside a persistent GEMM that leaves only a few SMs free
        // (dev::PanelSms): keep NCCL's kernels within that budget.  ELB200_NCCL_MAX_CTAS=0 keeps
        // NCCL's own choice.
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        const char* e = std::getenv("ELB200_NCCL_MAX_CTAS");
        const int maxCtas = e ? std::atoi(e) : 8;
        if (maxCtas > 0) { cfg.minCTAs = 1; cfg.maxCTAs = maxCtas; }
        ELB_NCCL(ncclCommInitRankConfig(&world_, worldSize, id, worldRank, &cfg));
        // column communicator: same grid column, ordered by grid row (Grid.cpp:151-156)
        ELB_NCCL(ncclCommSplit(world_, mrRank_, mcRank_, &mc_.nccl, nullptr));
        // row communicator: same grid row, ordered by grid column
        ELB_NCCL(ncclCommSplit(world_, mcRank_, mrRank_, &mr_.nccl, nullptr));
        // VC / VR orderings of all processes (Grid.cpp:165-166)
        if (order == COLUMN_MAJOR) {
            vc_.nccl = world_;
            ELB_NCCL(ncclCommSplit(world_, 0, vrRank_, &vr_.nccl, nullptr));
        } else {
            vr_.nccl = world_;
            ELB_NCCL(ncclCommSplit(world_, 0, vcRank_, &vc_.nccl, nullptr));
        }
        if (height_ == 1) { /* size-1 column comm is valid but never used for collectives */ }
        SetupP2P(worldSize);
    }
}

// Allocate this rank's exchange window, publish its IPC handle through one ncclAllGather and map the
// windows of all peers.  Collective: the path is enabled only if EVERY rank succeeded.
void Grid::SetupP2P(int worldSize) {
    // on by default since round 2 (parity suite green on 1x2, 2x1, 2x2 and 2x4); ELB200_P2P=0 keeps every wire
    // step on ncclSend / ncclRecv
    const char* e = std::getenv("ELB200_P2P");
    if ((e && std::atoi(e) == 0) || worldSize > elb200::P2P_MAX_PEERS) return;
    const char* r = std::getenv("ELB200_P2P_REGION_MB");
    const size_t regionBytes = (size_t)(r ? std::max(1, std::atoi(r)) : 64) << 20;
    const size_t bytes = 4096 + (size_t)4 * worldSize * regionBytes;
    cudaStream_t s = dev::stream();
    int ok = 1;
    char* local = nullptr;
    cudaIpcMemHandle_t mine;
    if (cudaMalloc((void**)&local, bytes) != cudaSuccess) { ok = 0; local = nullptr; cudaGetLastError(); }
    if (ok && cudaMemset(local, 0, 4096) != cudaSuccess) ok = 0;
    if (ok && cudaIpcGetMemHandle(&mine, local) != cudaSuccess) { ok = 0; cudaGetLastError(); }
    if (!ok) std::memset(&mine, 0, sizeof(mine));
    ELB_CUDA(cudaDeviceSynchronize());
    // handles (64 B each) and success flags travel through the world communicator
    const size_t hs = sizeof(cudaIpcMemHandle_t);
    char* dbuf = nullptr;
    ELB_CUDA(cudaMalloc((void**)&dbuf, hs * worldSize + sizeof(int)));
    ELB_CUDA(cudaMemcpy(dbuf + hs * worldRank_, &mine, hs, cudaMemcpyHostToDevice));
    ELB_NCCL(ncclAllGather(dbuf + hs * worldRank_, dbuf, hs, ncclInt8, world_, s));
    std::vector<cudaIpcMemHandle_t> handles(worldSize);
    ELB_CUDA(cudaStreamSynchronize(s));
    ELB_CUDA(cudaMemcpy(handles.data(), dbuf, hs * worldSize, cudaMemcpyDeviceToHost));
    std::vector<char*> peer(worldSize, nullptr);
    if (ok) {
        for (int w = 0; w < worldSize && ok; ++w) {
            if (w == worldRank_) { peer[w] = local; continue; }
            void* p = nullptr;
            if (cudaIpcOpenMemHandle(&p, handles[w], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            peer[w] = (char*)p;
        }
    }
    int* dflag = (int*)(dbuf + hs * worldSize);
    ELB_CUDA(cudaMemcpy(dflag, &ok, sizeof(int), cudaMemcpyHostToDevice));
    ELB_NCCL(ncclAllReduce(dflag, dflag, 1, ncclInt, ncclMin, world_, s));
    ELB_CUDA(cudaStreamSynchronize(s));
    int all = 0;
    ELB_CUDA(cudaMemcpy(&all, dflag, sizeof(int), cudaMemcpyDeviceToHost));
    cudaFree(dbuf);
    if (!all) {
        for (int w = 0; w < worldSize; ++w)
            if (w != worldRank_ && peer[w]) cudaIpcCloseMemHandle(peer[w]);
        if (local) cudaFree(local);
        if (worldRank_ == 0 && e) std::fprintf(stderr, "[elb200] ELB200_P2P=1 but the exchange windows could not be mapped; using NCCL\n");
        return;
    }
    p2p_.regionBytes = regionBytes;
    p2p_.local = local;
    p2p_.peer = peer;
    ELB_CUDA(cudaHostAlloc((void**)&p2p_.error, sizeof(int), cudaHostAllocMapped));
    *p2p_.error = 0;
    p2p_.on = true;
}

Grid::~Grid() {
    auto destroy = [&](ncclComm*& c) {
        if (c && c != world_) ncclCommDestroy(c);
        c = nullptr;
    };
    destroy(mc_.nccl);
    destroy(mr_.nccl);
    destroy(vc_.nccl);
    destroy(vr_.nccl);
    if (p2p_.on) {
        // nobody may still be writing flags into a window that is about to be unmapped
        cudaDeviceSynchronize();
        int* d = nullptr;
        if (world_ && cudaMalloc((void**)&d, sizeof(int)) == cudaSuccess) {
            cudaMemset(d, 0, sizeof(int));
            if (ncclAllReduce(d, d, 1, ncclInt, ncclSum, world_, nullptr) == ncclSuccess) cudaDeviceSynchronize();
            cudaFree(d);
        }
        for (size_t w = 0; w < p2p_.peer.size(); ++w)
            if ((int)w != worldRank_ && p2p_.peer[w]) cudaIpcCloseMemHandle(p2p_.peer[w]);
        if (p2p_.local) cudaFree(p2p_.local);
        if (p2p_.error) cudaFreeHost(p2p_.error);
        p2p_.on = false;
    }
    if (world_) ncclCommDestroy(world_);
    world_ = nullptr;
}

const Grid& Grid::Default() {
    static Grid g;
    return g;
}

int DistStride(Dist d, const Grid& g) {
    switch (d) {
        case MC: return g.Height();
        case MR: return g.Width();
        case VC: case VR: return g.Size();
        case STAR: return 1;
        default: LogicError("Distribution not supported on this path (MD/CIRC are out of scope)");
    }
}
int DistRankOf(Dist d, const Grid& g, int i, int j) {
    switch (d) {
        case MC: return i;
        case MR: return j;
        case VC: return i + g.Height() * j;
        case VR: return j + g.Width() * i;
        case STAR: return 0;
        default: LogicError("Distribution not supported on this path (MD/CIRC are out of scope)");
    }
}
int DistRank(Dist d, const Grid& g) { return DistRankOf(d, g, g.Row(), g.Col()); }
Dist PartialDist(Dist d) { return d == VC ? MC : (d == VR ? MR : d); }
Dist PartialUnionDist(Dist d) { return d == VC ? MR : (d == VR ? MC : STAR); }
const char* DistName(Dist d) {
    static const char* names[] = {"MC", "MD", "MR", "VC", "VR", "STAR", "CIRC"};
    return names[(int)d];
}

// ---------------------------------------------------------------------------
// Matrix<T>
// ---------------------------------------------------------------------------
namespace {
template <typename T>
Int PaddedLDim(Int height) {
    Int ld = height > 1 ? height : 1;
    const Int q = 16 / (Int)sizeof(T) > 0 ? 16 / (Int)sizeof(T) : 1;  // elements per 16 bytes
    return ((ld + q - 1) / q) * q;
}
}  // namespace

template <typename T> Matrix<T>::Matrix() {}
template <typename T> Matrix<T>::Matrix(Int h, Int w) { Resize(h, w); }
template <typename T> Matrix<T>::Matrix(Int h, Int w, Int ld) { Resize(h, w, ld); }
template <typename T> Matrix<T>::~Matrix() { Release(); }

template <typename T>
void Matrix<T>::Release() {
    if (owner_ && data_) {
        // destructors must not throw
        cudaFreeAsync(data_, dev::stream());
    }
    data_ = nullptr;
    capacity_ = 0;
}

template <typename T>
Matrix<T>::Matrix(const Matrix<T>& A) {
    Resize(A.height_, A.width_);
    Copy(A, *this);
}
template <typename T>
Matrix<T>::Matrix(Matrix<T>&& A) noexcept
    : height_(A.height_), width_(A.width_), ldim_(A.ldim_), data_(A.data_), capacity_(A.capacity_),
      owner_(A.owner_), locked_(A.locked_) {
    A.data_ = nullptr; A.capacity_ = 0; A.height_ = A.width_ = 0; A.ldim_ = 1; A.owner_ = true; A.locked_ = false;
}
template <typename T>
Matrix<T>& Matrix<T>::operator=(const Matrix<T>& A) {
    if (this != &A) { Resize(A.height_, A.width_); Copy(A, *this); }
    return *this;
}
template <typename T>
Matrix<T>& Matrix<T>::operator=(Matrix<T>&& A) noexcept {
    if (this != &A) {
        Release();
        height_ = A.height_; width_ = A.width_; ldim_ = A.ldim_; data_ = A.data_; capacity_ = A.capacity_;
        owner_ = A.owner_; locked_ = A.locked_;
        A.data_ = nullptr; A.capacity_ = 0; A.height_ = A.width_ = 0; A.ldim_ = 1; A.owner_ = true; A.locked_ = false;
    }
    return *this;
}

template <typename T>
void Matrix<T>::Empty(bool freeMemory) {
    if (!owner_) { data_ = nullptr; owner_ = true; locked_ = false; capacity_ = 0; }
    else if (freeMemory) Release();
    height_ = width_ = 0;
    ldim_ = 1;
}

template <typename T>
void Matrix<T>::Resize(Int h, Int w) {
    // Matrix/impl.hpp:698-717: the leading dimension changes only when the buffer has to grow ("simply shrink our
    // view if possible"), so the leading h x w block survives a shrink
    if (owner_ && data_ && h >= 0 && w >= 0 && h <= ldim_ && w <= width_) { height_ = h; width_ = w; return; }
    Resize(h, w, PaddedLDim<T>(h));
}

template <typename T>
void Matrix<T>::Resize(Int h, Int w, Int ld) {
    if (h < 0 || w < 0) LogicError("Height and width must be non-negative");
    if (ld < (h > 1 ? h : 1)) LogicError("Leading dimension must be no less than height");
    if (!owner_) {
        // a view may only shrink (Matrix/impl.hpp Resize on views)
        if (h > height_ || w > width_) LogicError("Cannot increase the size of a view");
        height_ = h; width_ = w;
        return;
    }
    const size_t need = size_t(ld) * size_t(w > 0 ? w : 0);
    if (need > capacity_) {
        Release();
        data_ = (T*)elb200::scratch_alloc(need * sizeof(T), dev::stream());
        capacity_ = need;
    }
    height_ = h; width_ = w; ldim_ = ld;
}

template <typename T>
void Matrix<T>::Attach(Int h, Int w, T* buf, Int ld) {
    Release();
    height_ = h; width_ = w; ldim_ = ld > 1 ? ld : 1; data_ = buf; owner_ = false; locked_ = false; capacity_ = 0;
}
template <typename T>
void Matrix<T>::LockedAttach(Int h, Int w, const T* buf, Int ld) {
    Attach(h, w, const_cast<T*>(buf), ld);
    locked_ = true;
}
template <typename T>
T* Matrix<T>::Buffer() {
    if (locked_) LogicError("Cannot return non-const buffer of a locked Matrix");
    return data_;
}
template <typename T>
T* Matrix<T>::Buffer(Int i, Int j) {
    if (locked_) LogicError("Cannot return non-const buffer of a locked Matrix");
    return data_ + size_t(i) + size_t(j) * size_t(ldim_);
}
template <typename T>
Matrix<T> Matrix<T>::operator()(Range I, Range J) {
    Matrix<T> V;
    const Int ie = (I.end == END || I.end > height_) ? height_ : I.end;
    const Int je = (J.end == END || J.end > width_) ? width_ : J.end;
    if (locked_) V.LockedAttach(ie - I.beg, je - J.beg, LockedBuffer(I.beg, J.beg), ldim_);
    else V.Attach(ie - I.beg, je - J.beg, Buffer(I.beg, J.beg), ldim_);
    return V;
}
template <typename T>
const Matrix<T> Matrix<T>::operator()(Range I, Range J) const {
    Matrix<T> V;
    const Int ie = (I.end == END || I.end > height_) ? height_ : I.end;
    const Int je = (J.end == END || J.end > width_) ? width_ : J.end;
    V.LockedAttach(ie - I.beg, je - J.beg, LockedBuffer(I.beg, J.beg), ldim_);
    return V;
}
template <typename T>
T Matrix<T>::Get(Int i, Int j) const {
    if (i < 0 || i >= height_ || j < 0 || j >= width_) LogicError("Matrix index out of bounds");
    T v;
    ELB_CUDA(cudaMemcpyAsync(&v, LockedBuffer(i, j), sizeof(T), cudaMemcpyDeviceToHost, dev::stream()));
    ELB_CUDA(cudaStreamSynchronize(dev::stream()));
    return v;
}
template <typename T>
void Matrix<T>::Set(Int i, Int j, T v) {
    if (i < 0 || i >= height_ || j < 0 || j >= width_) LogicError("Matrix index out of bounds");
    ELB_CUDA(cudaMemcpyAsync(Buffer(i, j), &v, sizeof(T), cudaMemcpyHostToDevice, dev::stream()));
    ELB_CUDA(cudaStreamSynchronize(dev::stream()));
}
template <typename T>
void Matrix<T>::ToHost(T* host, Int hld) const {
    if (height_ == 0 || width_ == 0) return;
    ELB_CUDA(cudaMemcpy2DAsync(host, size_t(hld) * sizeof(T), data_, size_t(ldim_) * sizeof(T),
                               size_t(height_) * sizeof(T), size_t(width_), cudaMemcpyDeviceToHost, dev::stream()));
    ELB_CUDA(cudaStreamSynchronize(dev::stream()));
}
template <typename T>
void Matrix<T>::FromHost(const T* host, Int hld) {
    if (height_ == 0 || width_ == 0) return;
    ELB_CUDA(cudaMemcpy2DAsync(Buffer(), size_t(ldim_) * sizeof(T), host, size_t(hld) * sizeof(T),
                               size_t(height_) * sizeof(T), size_t(width_), cudaMemcpyHostToDevice, dev::stream()));
    ELB_CUDA(cudaStreamSynchronize(dev::stream()));
}

template <typename T>
void Copy(const Matrix<T>& A, Matrix<T>& B) {
    B.Resize(A.Height(), A.Width());
    if (A.Height() == 0 || A.Width() == 0) return;
    ELB_CUDA(cudaMemcpy2DAsync(B.Buffer(), size_t(B.LDim()) * sizeof(T), A.LockedBuffer(),
                               size_t(A.LDim()) * sizeof(T), size_t(A.Height()) * sizeof(T), size_t(A.Width()),
                               cudaMemcpyDeviceToDevice, dev::stream()));
}

// ---------------------------------------------------------------------------
// AbstractDistMatrix<T>
// ---------------------------------------------------------------------------
template <typename T>
AbstractDistMatrix<T>::AbstractDistMatrix(const El::Grid& g, Dist U, Dist V) : grid_(&g), colDist_(U), rowDist_(V) {
    DistStride(U, g);  // validates that the distribution is one this path supports
    DistStride(V, g);
    SetShifts();
}
template <typename T> AbstractDistMatrix<T>::~AbstractDistMatrix() {}

template <typename T>
void AbstractDistMatrix<T>::SetShifts() {
    colShift_ = Shift_(ColRank(), colAlign_, ColStride());
    rowShift_ = Shift_(RowRank(), rowAlign_, RowStride());
}

template <typename T>
AbstractDistMatrix<T>::AbstractDistMatrix(const AbstractDistMatrix<T>& A)
    : grid_(A.grid_), colDist_(A.colDist_), rowDist_(A.rowDist_) {
    colAlign_ = A.colAlign_; rowAlign_ = A.rowAlign_;
    SetShifts();
    Resize(A.height_, A.width_);
    El::Copy(A.matrix_, matrix_);
}
template <typename T>
AbstractDistMatrix<T>::AbstractDistMatrix(AbstractDistMatrix<T>&& A) noexcept
    : grid_(A.grid_), colDist_(A.colDist_), rowDist_(A.rowDist_), height_(A.height_), width_(A.width_),
      colAlign_(A.colAlign_), rowAlign_(A.rowAlign_), colShift_(A.colShift_), rowShift_(A.rowShift_),
      colConstrained_(A.colConstrained_), rowConstrained_(A.rowConstrained_), viewing_(A.viewing_),
      locked_(A.locked_), matrix_(std::move(A.matrix_)) {
    A.height_ = A.width_ = 0; A.viewing_ = false; A.locked_ = false;
}
template <typename T>
AbstractDistMatrix<T>& AbstractDistMatrix<T>::operator=(const AbstractDistMatrix<T>& A) {
    if (this != &A) El::Copy(A, *this);
    return *this;
}
template <typename T>
AbstractDistMatrix<T>& AbstractDistMatrix<T>::operator=(AbstractDistMatrix<T>&& A) noexcept {
    if (this != &A) {
        if (colDist_ == A.colDist_ && rowDist_ == A.rowDist_ && !viewing_) {
            grid_ = A.grid_; height_ = A.height_; width_ = A.width_; colAlign_ = A.colAlign_; rowAlign_ = A.rowAlign_;
            colShift_ = A.colShift_; rowShift_ = A.rowShift_; colConstrained_ = A.colConstrained_;
            rowConstrained_ = A.rowConstrained_; viewing_ = A.viewing_; locked_ = A.locked_;
            matrix_ = std::move(A.matrix_);
            A.height_ = A.width_ = 0; A.viewing_ = false; A.locked_ = false;
        } else {
            El::Copy(static_cast<const AbstractDistMatrix<T>&>(A), *this);
        }
    }
    return *this;
}

template <typename T>
void AbstractDistMatrix<T>::Empty(bool freeMemory) {
    matrix_.Empty(freeMemory);
    height_ = width_ = 0;
    colAlign_ = rowAlign_ = 0;
    colConstrained_ = rowConstrained_ = false;
    viewing_ = locked_ = false;
    SetShifts();
}
template <typename T>
void AbstractDistMatrix<T>::Resize(Int h, Int w) {
    if (h < 0 || w < 0) LogicError("Height and width must be non-negative");
    if (viewing_ && (h > height_ || w > width_)) LogicError("Tried to increase the size of a view");
    height_ = h; width_ = w;
    matrix_.Resize(Length_(h, colShift_, ColStride()), Length_(w, rowShift_, RowStride()));
}
template <typename T>
void AbstractDistMatrix<T>::Resize(Int h, Int w, Int ld) {
    if (viewing_ && (h > height_ || w > width_)) LogicError("Tried to increase the size of a view");
    height_ = h; width_ = w;
    matrix_.Resize(Length_(h, colShift_, ColStride()), Length_(w, rowShift_, RowStride()), ld);
}
template <typename T>
void AbstractDistMatrix<T>::AlignCols(int a, bool constrain) {
    if (a < 0 || a >= ColStride()) LogicError("Invalid column alignment");
    if (colAlign_ != a) { matrix_.Empty(false); height_ = width_ = 0; viewing_ = locked_ = false; }
    if (constrain) colConstrained_ = true;
    colAlign_ = a;
    SetShifts();
}
template <typename T>
void AbstractDistMatrix<T>::AlignRows(int a, bool constrain) {
    if (a < 0 || a >= RowStride()) LogicError("Invalid row alignment");
    if (rowAlign_ != a) { matrix_.Empty(false); height_ = width_ = 0; viewing_ = locked_ = false; }
    if (constrain) rowConstrained_ = true;
    rowAlign_ = a;
    SetShifts();
}
template <typename T>
void AbstractDistMatrix<T>::Align(int ca, int ra, bool constrain) { AlignCols(ca, constrain); AlignRows(ra, constrain); }
template <typename T>
void AbstractDistMatrix<T>::FreeAlignments() { if (!viewing_) { colConstrained_ = rowConstrained_ = false; } }

// ElementalMatrix<T>::AlignColsWith / AlignRowsWith (src/core/DistMatrix/Element.cpp:204-258)
template <typename T>
void AbstractDistMatrix<T>::AlignColsWith(const AbstractDistMatrix<T>& A, bool constrain, bool allowMismatch) {
    grid_ = A.grid_;
    const Dist U = colDist_;
    if (A.colDist_ == U || A.colDist_ == PartialDist(U)) AlignCols(A.colAlign_ % ColStride(), constrain);
    else if (A.rowDist_ == U || A.rowDist_ == PartialDist(U)) AlignCols(A.rowAlign_ % ColStride(), constrain);
    else if (A.colDist_ == PartialUnionDist(U) && U != STAR) AlignCols(A.colAlign_ % ColStride(), constrain);
    else if (A.rowDist_ == PartialUnionDist(U) && U != STAR) AlignCols(A.rowAlign_ % ColStride(), constrain);
    else if (U != STAR && A.colDist_ != STAR && A.rowDist_ != STAR && !allowMismatch) LogicError("Nonsensical alignment");
}
template <typename T>
void AbstractDistMatrix<T>::AlignRowsWith(const AbstractDistMatrix<T>& A, bool constrain, bool allowMismatch) {
    grid_ = A.grid_;
    const Dist V = rowDist_;
    if (A.colDist_ == V || A.colDist_ == PartialDist(V)) AlignRows(A.colAlign_ % RowStride(), constrain);
    else if (A.rowDist_ == V || A.rowDist_ == PartialDist(V)) AlignRows(A.rowAlign_ % RowStride(), constrain);
    else if (A.colDist_ == PartialUnionDist(V) && V != STAR) AlignRows(A.colAlign_ % RowStride(), constrain);
    else if (A.rowDist_ == PartialUnionDist(V) && V != STAR) AlignRows(A.rowAlign_ % RowStride(), constrain);
    else if (V != STAR && A.colDist_ != STAR && A.rowDist_ != STAR && !allowMismatch) LogicError("Nonsensical alignment");
}
template <typename T>
void AbstractDistMatrix<T>::AlignWith(const AbstractDistMatrix<T>& A, bool constrain, bool allowMismatch) {
    AlignColsWith(A, constrain, allowMismatch);
    AlignRowsWith(A, constrain, allowMismatch);
}
template <typename T>
void AbstractDistMatrix<T>::AlignAndResize(int ca, int ra, Int h, Int w, bool force, bool constrain) {
    if (!viewing_) {
        if (force || !colConstrained_) { colAlign_ = Mod(ca, ColStride()); }
        if (force || !rowConstrained_) { rowAlign_ = Mod(ra, RowStride()); }
        SetShifts();
    }
    if (constrain) colConstrained_ = rowConstrained_ = true;
    if (force && (colAlign_ != Mod(ca, ColStride()) || rowAlign_ != Mod(ra, RowStride()))) LogicError("Could not set alignments");
    Resize(h, w);
}
template <typename T>
void AbstractDistMatrix<T>::Attach(Int h, Int w, const El::Grid& g, int ca, int ra, T* buf, Int ld) {
    matrix_.Empty();
    grid_ = &g; height_ = h; width_ = w; colAlign_ = ca; rowAlign_ = ra;
    colConstrained_ = rowConstrained_ = true; viewing_ = true; locked_ = false;
    SetShifts();
    matrix_.Attach(Length_(h, colShift_, ColStride()), Length_(w, rowShift_, RowStride()), buf, ld);
}
template <typename T>
void AbstractDistMatrix<T>::LockedAttach(Int h, Int w, const El::Grid& g, int ca, int ra, const T* buf, Int ld) {
    Attach(h, w, g, ca, ra, const_cast<T*>(buf), ld);
    matrix_.LockedAttach(matrix_.Height(), matrix_.Width(), buf, ld);
    locked_ = true;
}
template <typename T>
void AbstractDistMatrix<T>::LockedViewOf(const AbstractDistMatrix<T>& A, Int i, Int j, Int h, Int w) {
    if (A.colDist_ != colDist_ || A.rowDist_ != rowDist_) LogicError("A view must have the distribution of its parent");
    if (i < 0 || j < 0 || h < 0 || w < 0 || i + h > A.height_ || j + w > A.width_)
        LogicError("View is out of bounds of the parent matrix");
    matrix_.Empty();
    grid_ = A.grid_; height_ = h; width_ = w;
    colAlign_ = A.RowOwner(i); rowAlign_ = A.ColOwner(j);
    colConstrained_ = rowConstrained_ = true; viewing_ = true; locked_ = true;
    SetShifts();
    const Int iLoc = A.LocalRowOffset(i), jLoc = A.LocalColOffset(j);
    matrix_.LockedAttach(Length_(h, colShift_, ColStride()), Length_(w, rowShift_, RowStride()),
                         A.matrix_.LockedBuffer(iLoc, jLoc), A.LDim());
}
template <typename T>
void AbstractDistMatrix<T>::ViewOf(AbstractDistMatrix<T>& A, Int i, Int j, Int h, Int w) {
    if (A.locked_) LogicError("Cannot take a mutable view of a locked matrix");
    LockedViewOf(A, i, j, h, w);
    matrix_.Attach(matrix_.Height(), matrix_.Width(), const_cast<T*>(matrix_.LockedBuffer()), matrix_.LDim());
    locked_ = false;
}

#define ELB_INST(T)                         \
    template class Matrix<T>;               \
    template class AbstractDistMatrix<T>;   \
    template void Copy(const Matrix<T>&, Matrix<T>&);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
