// comm.cpp -- NCCL plumbing behind the C ABI (lbfgsb200_comm_*).
//
// One process per GPU.  The launcher (torchrun + torch.distributed in the Python harness, or
// MPI / any out-of-band channel in a C++ host) broadcasts the 128-byte unique id from rank 0;
// each rank then creates its communicator on its current CUDA device.
//
// NCCL is bound at run time (dlopen) instead of at link time: a single-GPU host never needs it,
// and inside a process that also carries PyTorch the library must use the SAME libnccl.so.2
// torch loaded (torch bundles a newer NCCL than the system one; two copies of one SONAME cannot
// coexist).  RTLD_NOLOAD picks up an already-loaded copy first.
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h> // types and enums only; no NCCL symbol is linked
#include <stdio.h>
#include <string.h>

namespace lb {
void set_error(const char *fmt, ...);

namespace {
struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl()
{
    if (g_nccl.handle) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        set_error("cannot load libnccl.so.2: %s", dlerror());
        return LBFGSB200_ERR_NCCL;
    }
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather ||
        !g_nccl.GetErrorString) {
        set_error("libnccl.so.2 lacks a required symbol");
        return LBFGSB200_ERR_NCCL;
    }
    g_nccl.handle = h;
    return 0;
}
} // namespace

int comm_allgather(lbfgsb200_comm *c, const double *send, double *recv, int count,
                   cudaStream_t stream)
{
    ncclResult_t r = g_nccl.AllGather(send, recv, (size_t)count, ncclDouble, (ncclComm_t)c->nccl, stream);
    if (r != ncclSuccess) {
        set_error("ncclAllGather: %s", g_nccl.GetErrorString(r));
        return LBFGSB200_ERR_NCCL;
    }
    return 0;
}
} // namespace lb

static_assert(sizeof(ncclUniqueId) <= LBFGSB200_UNIQUE_ID_BYTES, "unique id does not fit");

extern "C" int lbfgsb200_comm_unique_id(char id[LBFGSB200_UNIQUE_ID_BYTES])
{
    if (int rc = lb::load_nccl()) return rc;
    ncclUniqueId u;
    ncclResult_t r = lb::g_nccl.GetUniqueId(&u);
    if (r != ncclSuccess) {
        lb::set_error("ncclGetUniqueId: %s", lb::g_nccl.GetErrorString(r));
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
    if (int rc = lb::load_nccl()) return rc;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    ncclComm_t comm;
    ncclResult_t r = lb::g_nccl.CommInitRank(&comm, nranks, u, rank);
    if (r != ncclSuccess) {
        lb::set_error("ncclCommInitRank: %s", lb::g_nccl.GetErrorString(r));
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
    if (lb::g_nccl.handle) lb::g_nccl.CommDestroy((ncclComm_t)c->nccl);
    delete c;
}
