// BINARY and BINARY_FLAT file formats of the reference, read into / written from DEVICE-resident matrices.
//   read::BinaryFlat   src/io/Read/BinaryFlat.hpp:16-34 (Matrix), :37-102 (AbstractDistMatrix)
//   read::Binary       src/io/Read/Binary.hpp:16-41, :44-...  (two El::Int header words: height, width)
//   write::Binary      src/io/Write/Binary.hpp:16-36;  write::BinaryFlat  src/io/Write/BinaryFlat.hpp:16-33
//   Write(AbstractDistMatrix) src/io/Write.cpp:46-63 gathers to [CIRC,CIRC] and lets the root write.
// Bytes on disk are identical to the reference's: column-major entries of T, no padding; BINARY prefixes two
// 32-bit Ints; extensions "bin" / "dat" (src/io/File.cpp:34-35).
//
// B200-first: the reference reads element by element with seekg for a 2-D distribution (BinaryFlat.hpp:82-99) and
// funnels every write through one process.  Here every process moves whole COLUMNS -- the file's contiguous unit --
// through a [*,VC] intermediate: process q reads (writes) the columns j = q mod p with pread (pwrite) into pinned
// staging buffers, double-buffered against the host<->device copies, and the redistribution engine turns [*,VC]
// into the caller's distribution over NVLink.  I/O is p-way parallel and every request is height * sizeof(T) bytes.
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>

#include "dev.hpp"
#include "elb200/io.hpp"

namespace El {

namespace {

struct Fd {
    int fd = -1;
    Fd(const std::string& name, int flags, int mode = 0644) : fd(::open(name.c_str(), flags, mode)) {
        if (fd < 0) RuntimeError("Could not open " + name);
    }
    ~Fd() { if (fd >= 0) ::close(fd); }
};
struct Pinned {
    char* p = nullptr;
    explicit Pinned(size_t bytes) { ELB_CUDA(cudaHostAlloc((void**)&p, bytes ? bytes : 1, cudaHostAllocDefault)); }
    ~Pinned() { if (p) cudaFreeHost(p); }
};

void PreadAll(int fd, char* dst, size_t bytes, off_t pos, const std::string& name) {
    while (bytes) {
        const ssize_t got = ::pread(fd, dst, bytes, pos);
        if (got <= 0) RuntimeError("Short read from " + name);
        dst += got; pos += got; bytes -= (size_t)got;
    }
}
void PwriteAll(int fd, const char* src, size_t bytes, off_t pos, const std::string& name) {
    while (bytes) {
        const ssize_t put = ::pwrite(fd, src, bytes, pos);
        if (put <= 0) RuntimeError("Short write to " + name);
        src += put; pos += put; bytes -= (size_t)put;
    }
}

// local columns of M (local height == height: every column is whole) <-> file columns GlobalCol(jLoc), through two
// pinned staging buffers so that the file system and the copy engine work at the same time
template <typename T>
void MoveColumns(bool reading, AbstractDistMatrix<T>& M, int fd, off_t base, Int height, const std::string& name) {
    const Int lw = M.LocalWidth();
    if (lw == 0 || height == 0) return;
    const size_t colBytes = size_t(height) * sizeof(T);
    const Int batch = (Int)std::max<size_t>(1, std::min<size_t>((size_t)lw, (size_t(32) << 20) / colBytes));
    Pinned stage[2] = {Pinned(colBytes * batch), Pinned(colBytes * batch)};
    dev::Event done[2];
    bool used[2] = {false, false};
    cudaStream_t s = dev::stream();
    for (Int j0 = 0, it = 0; j0 < lw; j0 += batch, ++it) {
        const int b = it & 1;
        const Int nb = std::min(batch, lw - j0);
        if (used[b]) ELB_CUDA(cudaEventSynchronize(done[b].e));   // the copy that last used this buffer
        if (reading) {
            for (Int j = 0; j < nb; ++j)
                PreadAll(fd, stage[b].p + colBytes * j, colBytes, base + off_t(M.GlobalCol(j0 + j)) * off_t(colBytes), name);
            ELB_CUDA(cudaMemcpy2DAsync(M.Buffer() + size_t(j0) * M.LDim(), sizeof(T) * size_t(M.LDim()), stage[b].p, colBytes,
                                       colBytes, size_t(nb), cudaMemcpyHostToDevice, s));
            done[b].Record(s);
            used[b] = true;
        } else {
            ELB_CUDA(cudaMemcpy2DAsync(stage[b].p, colBytes, M.LockedBuffer() + size_t(j0) * M.LDim(),
                                       sizeof(T) * size_t(M.LDim()), colBytes, size_t(nb), cudaMemcpyDeviceToHost, s));
            ELB_CUDA(cudaStreamSynchronize(s));
            for (Int j = 0; j < nb; ++j)
                PwriteAll(fd, stage[b].p + colBytes * j, colBytes, base + off_t(M.GlobalCol(j0 + j)) * off_t(colBytes), name);
        }
    }
    ELB_CUDA(cudaStreamSynchronize(s));
}

off_t FileSizeOf(int fd) {
    struct stat st;
    if (::fstat(fd, &st) != 0) RuntimeError("fstat failed");
    return st.st_size;
}

template <typename T>
void ReadColumns(AbstractDistMatrix<T>& A, Int height, Int width, const std::string& filename, off_t base) {
    Fd f(filename, O_RDONLY);
    const off_t expect = base + off_t(height) * off_t(width) * off_t(sizeof(T));
    const off_t have = FileSizeOf(f.fd);
    if (have != expect)
        RuntimeError("Expected file to be " + std::to_string((long long)expect) + " bytes but found " + std::to_string((long long)have));
    const Grid& g = A.Grid();
    A.Resize(height, width);
    if (A.ColDist() == STAR && (A.RowDist() == VC || A.RowDist() == VR || A.RowDist() == STAR || A.RowDist() == MR || A.RowDist() == MC)) {
        MoveColumns(true, A, f.fd, base, height, filename);   // whole columns already
        return;
    }
    AbstractDistMatrix<T> cols(g, STAR, VC);
    cols.Resize(height, width);
    MoveColumns(true, cols, f.fd, base, height, filename);
    Copy(static_cast<const AbstractDistMatrix<T>&>(cols), A);
}

template <typename T>
void WriteColumns(const AbstractDistMatrix<T>& A, const std::string& filename, bool header) {
    const Grid& g = A.Grid();
    const Int height = A.Height(), width = A.Width();
    const off_t base = header ? off_t(2 * sizeof(Int)) : 0;
    // every process holds whole columns of a [*,VC] copy; process 0 also writes the header and sizes the file
    AbstractDistMatrix<T> cols(g, STAR, VC);
    Copy(A, cols);
    ELB_CUDA(cudaStreamSynchronize(dev::stream()));
    if (g.VCRank() == 0) {
        Fd f(filename, O_WRONLY | O_CREAT | O_TRUNC);
        if (header) {
            Int hw[2] = {height, width};
            PwriteAll(f.fd, (const char*)hw, sizeof(hw), 0, filename);
        }
        if (::ftruncate(f.fd, base + off_t(height) * off_t(width) * off_t(sizeof(T))) != 0) RuntimeError("ftruncate failed");
    }
    // the file must exist with its final size before the other processes open it
    if (g.Size() > 1) {
        double* d = (double*)elb200::scratch_alloc(sizeof(double), dev::stream());
        ELB_CUDA(cudaMemsetAsync(d, 0, sizeof(double), dev::stream()));
        ELB_NCCL(ncclAllReduce(d, d, 1, ncclDouble, ncclSum, (ncclComm_t)g.VCComm().nccl, dev::stream()));
        ELB_CUDA(cudaStreamSynchronize(dev::stream()));
        elb200::scratch_free(d, dev::stream());
    }
    Fd f(filename, O_WRONLY);
    MoveColumns(false, cols, f.fd, base, height, filename);
}

}  // namespace

namespace read {
template <typename T>
void BinaryFlat(AbstractDistMatrix<T>& A, Int height, Int width, const std::string& filename) {
    ReadColumns(A, height, width, filename, 0);
}
template <typename T>
void Binary(AbstractDistMatrix<T>& A, const std::string& filename) {
    Int hw[2] = {0, 0};
    {
        Fd f(filename, O_RDONLY);
        if (FileSizeOf(f.fd) < (off_t)sizeof(hw)) RuntimeError("File too short for a BINARY header: " + filename);
        PreadAll(f.fd, (char*)hw, sizeof(hw), 0, filename);
    }
    if (hw[0] < 0 || hw[1] < 0) RuntimeError("Corrupt BINARY header in " + filename);
    ReadColumns(A, hw[0], hw[1], filename, off_t(sizeof(hw)));
}
}  // namespace read

namespace write {
template <typename T>
void Binary(const AbstractDistMatrix<T>& A, const std::string& basename) { WriteColumns(A, basename + ".bin", true); }
template <typename T>
void BinaryFlat(const AbstractDistMatrix<T>& A, const std::string& basename) { WriteColumns(A, basename + ".dat", false); }
}  // namespace write

#define ELB_INST(T)                                                                                   \
    template void read::BinaryFlat(AbstractDistMatrix<T>&, Int, Int, const std::string&);             \
    template void read::Binary(AbstractDistMatrix<T>&, const std::string&);                           \
    template void write::Binary(const AbstractDistMatrix<T>&, const std::string&);                    \
    template void write::BinaryFlat(const AbstractDistMatrix<T>&, const std::string&);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
