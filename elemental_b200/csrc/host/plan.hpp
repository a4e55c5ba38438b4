// Planning half of the redistribution engine (no CUDA, no NCCL): which strided lattice of
// the local matrices travels between which pair of ranks.  Shared by redist.cpp (device
// execution) and the C-ABI in include/elb200_plan.h (CPU/gloo tests).
#pragma once
#include <vector>

#include "elb200/core.hpp"

namespace El {
namespace plan {

typedef long long i64;

struct Layout {
    Dist U, V;
    int colAlign, rowAlign;
};

// One (source rank -> destination rank) message, in DESTINATION orientation:
// element (t,u), t < nrows, u < ncols, is
//   A_local[s_off + t*s_rs + u*s_cs]   on the source rank and
//   B_local[d_off + t*d_rs + u*d_cs]   on the destination rank.
struct Msg {
    bool empty = true;
    i64 nrows = 0, ncols = 0;
    i64 s_off = 0, s_rs = 0, s_cs = 0;
    i64 d_off = 0, d_rs = 0, d_cs = 0;
    i64 count() const { return nrows * ncols; }
};

bool PinsRow(Dist d);  // the distribution fixes the grid-row coordinate of the owner
bool PinsCol(Dist d);

// chooseOwner: when A is replicated, only the replica sharing the destination's free
// grid coordinate(s) sends.  Sum-scatter passes false (every replica contributes).
Msg ComputeMsg(const Grid& g, Int height, Int width, const Layout& A, i64 ldA, int si, int sj, const Layout& B,
               i64 ldB, int di, int dj, bool transpose, bool chooseOwner);

// send[v] / recv[v]: message to / from the rank with VC rank v (= mcRank + r*mrRank);
// send[me] is the local part.
struct RedistPlan {
    std::vector<Msg> send, recv;
};
RedistPlan BuildRedistPlan(const Grid& g, Int height, Int width, const Layout& A, i64 ldA, const Layout& B, i64 ldB,
                           bool transpose);

enum CommKind { OVER_MR = 0, OVER_MC = 1, OVER_VC = 2, NOT_REPLICATED = 3 };
struct ContractPlan {
    CommKind kind = NOT_REPLICATED;
    Layout T{MC, MR, 0, 0};
    i64 chunk = 0;            // elements per communicator member in the send buffer
    i64 myRows = 0;           // local height of my piece of T
    bool needZero = false;    // some chunk is smaller than `chunk`: the buffer must be zeroed first
    std::vector<Msg> packs;   // per member q: my A_local lattice -> sendbuf + q*chunk
};
// bView: B's layout as seen in A's index space (swap the roles for a transposed output)
ContractPlan BuildContractPlan(const Grid& g, Int height, Int width, const Layout& A, i64 ldA, const Layout& bView);

}  // namespace plan
}  // namespace El
