// comm.h -- the one collective the solver needs: a tiny packed all-gather over NCCL/NVLink.
//
// Every scalar sync point of the iteration ships ONE packet of kPacket doubles per rank
// (partial sums + the shard's boundary x / g / d values for the one-element halo).  Every
// rank then adds the packets in rank order, so the totals are bit-identical on all ranks
// and independent of NCCL's algorithm choice.  The reference has no multi-GPU path; this
// is new work.
#pragma once
#include <cuda_runtime.h>

#include "../../include/lbfgsb200.h"

struct lbfgsb200_comm {
    void *nccl; // ncclComm_t
    int rank;
    int nranks;
};

namespace lb {
// all-gather `count` doubles per rank: recv[r*count + i] = send_r[i].  Stream-ordered; safe to
// capture in a CUDA graph.  Returns 0 or LBFGSB200_ERR_NCCL.
int comm_allgather(lbfgsb200_comm *c, const double *send, double *recv, int count,
                   cudaStream_t stream);
} // namespace lb
