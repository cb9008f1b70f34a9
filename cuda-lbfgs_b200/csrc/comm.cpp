// comm.cpp -- NCCL plumbing behind the C ABI (lbfgsb200_comm_*).
//
// One process per GPU.  The launcher (torchrun + torch.distributed in the Python harness, or
// MPI / any out-of-band channel in a C++ host) broadcasts the 128-byte unique id from rank 0;
// each rank then creates its communicator on its current CUDA device.
#include "comm.h"

#include <nccl.h>
#include <stdio.h>
#include <string.h>

namespace lb {
void set_error(const char *fmt, ...);

int comm_allgather(lbfgsb200_comm *c, const double *send, double *recv, int count,
                   cudaStream_t stream)
{
    ncclResult_t r =
        ncclAllGather(send, recv, (size_t)count, ncclDouble, (ncclComm_t)c->nccl, stream);
    if (r != ncclSuccess) {
        set_error("ncclAllGather: %s", ncclGetErrorString(r));
        return LBFGSB200_ERR_NCCL;
    }
    return 0;
}
} // namespace lb

static_assert(sizeof(ncclUniqueId) <= LBFGSB200_UNIQUE_ID_BYTES, "unique id does not fit");

extern "C" int lbfgsb200_comm_unique_id(char id[LBFGSB200_UNIQUE_ID_BYTES])
{
    ncclUniqueId u;
    ncclResult_t r = ncclGetUniqueId(&u);
    if (r != ncclSuccess) {
        lb::set_error("ncclGetUniqueId: %s", ncclGetErrorString(r));
        return LBFGSB200_ERR_NCCL;
    }
    memset(id, 0, LBFGSB200_UNIQUE_ID_BYTES);
    memcpy(id, &u, sizeof u);
    return 0;
}

extern "C" int lbfgsb200_comm_create(lbfgsb200_comm_t **out, const char id[LBFGSB200_UNIQUE_ID_BYTES],
                                     int rank, int nranks)
{
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) {
        lb::set_error("comm_create: bad arguments (rank %d of %d)", rank, nranks);
        return LBFGSB200_ERR_INVALID;
    }
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm;
    ncclResult_t r = ncclCommInitRank(&comm, nranks, u, rank);
    if (r != ncclSuccess) {
        lb::set_error("ncclCommInitRank: %s", ncclGetErrorString(r));
        return LBFGSB200_ERR_NCCL;
    }
    lbfgsb200_comm *c = new lbfgsb200_comm;
    c->nccl = comm;
    c->rank = rank;
    c->nranks = nranks;
    *out = c;
    return 0;
}

extern "C" void lbfgsb200_comm_destroy(lbfgsb200_comm_t *c)
{
    if (!c) return;
    ncclCommDestroy((ncclComm_t)c->nccl);
    delete c;
}
