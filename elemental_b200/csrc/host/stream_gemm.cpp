// GemmHost: El::Gemm for operands whose [MC,MR] local matrices live in HOST memory -- what a caller of the
// reference has (its DistMatrix buffers are host allocations, include/El/core/Memory/impl.hpp) -- streamed through
// HBM instead of "copy three matrices in, multiply, copy one out".
//
// C is cut into bands of columns.  Band j needs: all of op(A) (resident after the first band), op(B)(:, J) and
// C(:, J).  Three streams:
//   copy-in  : C(:, J), op(B)(:, J) of band j + 1 while band j is multiplied; during band 0 also A, in chunks of
//              the summation index in the order the SUMMA panel loop consumes them;
//   compute  : band 0 = one Gemm per A chunk (started as soon as that chunk has landed), later bands = one Gemm;
//   copy-out : the finished band goes back to the host while the next one is multiplied (PCIe is full duplex).
// Every entry of C still receives its rank-Blocksize() updates in the same order (chunk edges are multiples of the
// blocksize, band edges never split a k-panel), so the result is bit-identical to El::Gemm on device-resident
// operands (tests/test_el_gpu.py::test_gemm_host_streamed_matches_device_gemm).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "dev.hpp"
#include "elb200/level3.hpp"

namespace El {

namespace {

Int Gcd(Int a, Int b) { return b == 0 ? a : Gcd(b, a % b); }
Int Lcm(Int a, Int b) { return a / Gcd(a, b) * b; }
Int RoundUp(Int x, Int q) { return (x + q - 1) / q * q; }

template <typename T>
AbstractDistMatrix<T> LockedView(const AbstractDistMatrix<T>& A, Int i, Int j, Int h, Int w) {
    AbstractDistMatrix<T> V(A.Grid(), A.ColDist(), A.RowDist());
    V.LockedViewOf(A, i, j, h, w);
    return V;
}

// rows [i0, i0 + h) x columns [j0, j0 + w) (GLOBAL, both offsets multiples of the grid strides) of a host-resident
// [MC,MR] matrix <-> the device matrix D (h x w, alignments 0), on stream s
template <typename T>
void CopyBlock(bool toDevice, const HostLocalMatrix<T>& H, Int i0, Int j0, AbstractDistMatrix<T>& D, cudaStream_t s) {
    const Grid& g = D.Grid();
    const Int li = i0 / g.Height(), lj = j0 / g.Width();
    const Int lh = D.LocalHeight(), lw = D.LocalWidth();
    if (lh == 0 || lw == 0) return;
    T* host = H.buffer + size_t(li) + size_t(lj) * size_t(H.ldim);
    if (toDevice)
        ELB_CUDA(cudaMemcpy2DAsync(D.Buffer(), sizeof(T) * size_t(D.LDim()), host, sizeof(T) * size_t(H.ldim),
                                   sizeof(T) * size_t(lh), size_t(lw), cudaMemcpyHostToDevice, s));
    else
        ELB_CUDA(cudaMemcpy2DAsync(host, sizeof(T) * size_t(H.ldim), D.LockedBuffer(), sizeof(T) * size_t(D.LDim()),
                                   sizeof(T) * size_t(lh), size_t(lw), cudaMemcpyDeviceToHost, s));
}

}  // namespace

template <typename T>
void GemmHost(Orientation oA, Orientation oB, T alpha, const Grid& g, const HostLocalMatrix<T>& A,
              const HostLocalMatrix<T>& B, T beta, const HostLocalMatrix<T>& C, GemmAlgorithm alg,
              GemmHostStats* stats) {
    const Int m = C.height, n = C.width;
    const Int am = (oA == NORMAL) ? A.height : A.width, ak = (oA == NORMAL) ? A.width : A.height;
    const Int bk = (oB == NORMAL) ? B.height : B.width, bn = (oB == NORMAL) ? B.width : B.height;
    if (am != m || bn != n || ak != bk) LogicError("Nonconformal matrices in GemmHost");
    const Int k = ak;
    if (alg == GEMM_DEFAULT) alg = GemmDefaultAlgorithm(m, n, k);
    const Int bsize = Blocksize();
    // offsets that keep every local offset integral on both grid dimensions (and 16-element aligned for TMA)
    const Int granule = 16 * Lcm(g.Height(), g.Width());
    // ELB200_GEMMHOST_BANDS / _CHUNKS (read at every call) override the number of column bands of C / chunks of the
    // summation index; results do not depend on them
    auto envInt = [](const char* name, Int dflt) { const char* e = std::getenv(name); const Int v = e ? std::atoi(e) : 0; return v > 0 ? v : dflt; };
    // Default number of bands: local bands of about 512 MB, between 2 and 8.  Every band repeats the whole panel loop
    // (the A panels are gathered again) and pays its own first copy-in / last copy-out, so small local matrices want
    // few bands; large ones want many, so that the copy of band j + 1 hides behind the product of band j.  Measured
    // (profiles/r02_gemmhost_bands_n8.txt, n = 32768 on 2x4, local C 1.07 GB): 8 bands 492 ms, 4 bands 434 ms,
    // 2 bands 413 ms; on one GPU (local C 8.6 GB) 8 bands hide all but 9 % of the copies.
    const double localCBytes = double(sizeof(T)) * double((m + g.Height() - 1) / g.Height()) * double((n + g.Width() - 1) / g.Width());
    const Int autoBands = (Int)std::min(8.0, std::max(2.0, std::ceil(localCBytes / double(size_t(512) << 20))));
    const Int wantBands = envInt("ELB200_GEMMHOST_BANDS", autoBands), wantChunks = envInt("ELB200_GEMMHOST_CHUNKS", 8);
    const Int wb = std::max<Int>(granule, RoundUp((n + wantBands - 1) / wantBands, granule));
    const Int kc = std::max<Int>(Lcm(granule, bsize), RoundUp((k + wantChunks - 1) / wantChunks, Lcm(granule, bsize)));
    const Int nbands = std::max<Int>(1, (n + wb - 1) / wb);
    const Int nchunks = std::max<Int>(1, (k + kc - 1) / kc);

    cudaStream_t compute = dev::stream(), copyIn = elb200::aux_stream(1), copyOut = elb200::aux_stream(2);
    AbstractDistMatrix<T> Ad(g, MC, MR);
    Ad.Resize(A.height, A.width);
    AbstractDistMatrix<T> Bd[2] = {AbstractDistMatrix<T>(g, MC, MR), AbstractDistMatrix<T>(g, MC, MR)};
    AbstractDistMatrix<T> Cd[2] = {AbstractDistMatrix<T>(g, MC, MR), AbstractDistMatrix<T>(g, MC, MR)};
    const Int wb0 = std::min(wb, n);
    for (int s = 0; s < 2; ++s) {
        if (s == 1 && nbands == 1) break;
        if (oB == NORMAL) Bd[s].Resize(k, wb0); else Bd[s].Resize(wb0, k);
        Cd[s].Resize(m, wb0);
    }
    dev::Event allocated, joinIn, joinOut;
    std::vector<dev::Event> aReady(nchunks), inReady(2), computed(2), slotFree(2);
    // the buffers were allocated in compute-stream order: the copy streams may touch them only after that point
    allocated.Record(compute);
    allocated.Wait(copyIn);
    allocated.Wait(copyOut);

    auto bandWidth = [&](Int j) { return std::min(wb, n - j * wb); };
    auto loadBand = [&](Int j) {
        const int s = int(j & 1);
        const Int w = bandWidth(j);
        if (j >= 2) slotFree[s].Wait(copyIn);   // band j - 2 has left this slot
        if (oB == NORMAL) { Bd[s].Resize(k, w); CopyBlock(true, B, 0, j * wb, Bd[s], copyIn); }
        else { Bd[s].Resize(w, k); CopyBlock(true, B, j * wb, 0, Bd[s], copyIn); }
        Cd[s].Resize(m, w);
        CopyBlock(true, C, 0, j * wb, Cd[s], copyIn);
        inReady[s].Record(copyIn);
    };
    // band 0 first, then A in the order the panel loop consumes it
    loadBand(0);
    for (Int i = 0; i < nchunks; ++i) {
        const Int k0 = i * kc, kw = std::min(kc, k - k0);
        AbstractDistMatrix<T> V(g, MC, MR);
        if (oA == NORMAL) { V.ViewOf(Ad, 0, k0, m, kw); CopyBlock(true, A, 0, k0, V, copyIn); }
        else { V.ViewOf(Ad, k0, 0, kw, m); CopyBlock(true, A, k0, 0, V, copyIn); }
        aReady[i].Record(copyIn);
    }
    for (Int j = 0; j < nbands; ++j) {
        const int s = int(j & 1);
        if (j + 1 < nbands) loadBand(j + 1);   // overlaps the product of band j
        inReady[s].Wait(compute);
        if (j == 0) {
            for (Int i = 0; i < nchunks; ++i) {
                const Int k0 = i * kc, kw = std::min(kc, k - k0);
                aReady[i].Wait(compute);
                auto Av = (oA == NORMAL) ? LockedView(static_cast<const AbstractDistMatrix<T>&>(Ad), 0, k0, m, kw)
                                         : LockedView(static_cast<const AbstractDistMatrix<T>&>(Ad), k0, 0, kw, m);
                auto Bv = (oB == NORMAL) ? LockedView(static_cast<const AbstractDistMatrix<T>&>(Bd[s]), k0, 0, kw, Bd[s].Width())
                                         : LockedView(static_cast<const AbstractDistMatrix<T>&>(Bd[s]), 0, k0, Bd[s].Height(), kw);
                Gemm(oA, oB, alpha, static_cast<const AbstractDistMatrix<T>&>(Av), static_cast<const AbstractDistMatrix<T>&>(Bv),
                     i == 0 ? beta : T(1), Cd[s], alg);
            }
        } else {
            Gemm(oA, oB, alpha, static_cast<const AbstractDistMatrix<T>&>(Ad), static_cast<const AbstractDistMatrix<T>&>(Bd[s]),
                 beta, Cd[s], alg);
        }
        computed[s].Record(compute);
        computed[s].Wait(copyOut);
        CopyBlock(false, C, 0, j * wb, Cd[s], copyOut);
        slotFree[s].Record(copyOut);
    }
    // the call returns with C in host memory and every buffer released in compute-stream order
    joinIn.Record(copyIn);
    joinOut.Record(copyOut);
    joinIn.Wait(compute);
    joinOut.Wait(compute);
    ELB_CUDA(cudaStreamSynchronize(compute));
    if (stats) { stats->bands = (int)nbands; stats->chunks = (int)nchunks; stats->bandWidth = wb; stats->chunkWidth = kc; }
}

#define ELB_INST(T)                                                                                             \
    template void GemmHost(Orientation, Orientation, T, const Grid&, const HostLocalMatrix<T>&,                 \
                           const HostLocalMatrix<T>&, T, const HostLocalMatrix<T>&, GemmAlgorithm, GemmHostStats*);
ELB_INST(float)
ELB_INST(double)
ELB_INST(Complex<float>)
ELB_INST(Complex<double>)

}  // namespace El
