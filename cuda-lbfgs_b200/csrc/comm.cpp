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
#include <stdlib.h>
#include <string.h>

#include <vector>

namespace lb {
int setup_mailboxes(lbfgsb200_comm *c);
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

namespace lb {
// Allocate this rank's mailbox, exchange the CUDA IPC handles through NCCL, map every peer's mailbox.  Every
// fallible LOCAL step (allocations, mappings, the upload of the pointer table) happens before the last agreement
// round, and its outcome is part of what that round gathers: all ranks reach the same verdict, so none is left
// spinning on a mailbox while its peers sit in NCCL.  On a negative verdict everything is released again.
int setup_mailboxes(lbfgsb200_comm *c)
{
    const int P = c->nranks;
    const size_t mail_bytes = LBFGSB200_MAIL_DOUBLES * sizeof(double);
    int dev = 0;
    cudaGetDevice(&dev);
    c->device = dev;
    struct Blob {
        cudaIpcMemHandle_t handle;
        int device;
        int ok;
    };
    Blob mine;
    memset(&mine, 0, sizeof mine);
    mine.device = dev;
    mine.ok = 1;
    char *d_buf = nullptr;
    if (cudaMalloc(&c->mail, mail_bytes) != cudaSuccess) mine.ok = 0;
    if (mine.ok && cudaMemset(c->mail, 0, mail_bytes) != cudaSuccess) mine.ok = 0;
    if (mine.ok && cudaIpcGetMemHandle(&mine.handle, c->mail) != cudaSuccess) mine.ok = 0;
    if (cudaMalloc(&c->peers_dev, sizeof(double *) * (size_t)P) != cudaSuccess) mine.ok = 0;
    // the gather buffer itself: without it this rank cannot even take part in the agreement, so it is the one
    // failure that has to be reported as "no NCCL exchange possible either"
    if (cudaMalloc(&d_buf, sizeof(Blob) * (size_t)(P + 1)) != cudaSuccess) {
        cudaGetLastError();
        return -2;
    }
    cudaGetLastError();
    std::vector<Blob> all((size_t)P);
    auto gather = [&]() -> int {
        if (cudaMemcpy(d_buf, &mine, sizeof mine, cudaMemcpyHostToDevice) != cudaSuccess) return -1;
        if (g_nccl.AllGather(d_buf, d_buf + sizeof(Blob), sizeof(Blob), ncclChar, (ncclComm_t)c->nccl, 0) != ncclSuccess) return -1;
        if (cudaDeviceSynchronize() != cudaSuccess) return -1;
        if (cudaMemcpy(all.data(), d_buf + sizeof(Blob), sizeof(Blob) * (size_t)P, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        return 0;
    };
    int rc = gather(); // round 1: handles
    std::vector<double *> peers((size_t)P, nullptr);
    if (rc == 0) {
        for (int r = 0; r < P && mine.ok; ++r) {
            if (!all[r].ok) { mine.ok = 0; break; }
            if (r == c->rank) { peers[r] = c->mail; continue; }
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, dev, all[r].device) != cudaSuccess || !can) { mine.ok = 0; break; }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { mine.ok = 0; break; }
            c->opened[r] = p;
            peers[r] = (double *)p;
        }
        if (mine.ok && cudaMemcpy(c->peers_dev, peers.data(), sizeof(double *) * (size_t)P, cudaMemcpyHostToDevice) != cudaSuccess) mine.ok = 0;
        cudaGetLastError();
        rc = gather(); // round 2: every rank learns whether EVERY rank mapped everything
    }
    cudaFree(d_buf);
    int all_ok = (rc == 0);
    for (int r = 0; r < P && all_ok; ++r) all_ok &= all[r].ok;
    if (!all_ok) { // identical on every rank (a failed gather fails on all of them): fall back together
        for (int r = 0; r < LBFGSB200_MAIL_RANKS; ++r)
            if (c->opened[r]) { cudaIpcCloseMemHandle(c->opened[r]); c->opened[r] = nullptr; }
        if (c->peers_dev) { cudaFree(c->peers_dev); c->peers_dev = nullptr; }
        if (c->mail) { cudaFree(c->mail); c->mail = nullptr; }
        cudaGetLastError();
        return -1;
    }
    c->p2p = 1;
    return 0;
}
} // namespace lb

// In-process communicator: P devices driven by threads of ONE process (lbfgsb200_solve with num_gpus > 1).
// Mailboxes are plain device allocations made reachable with cudaDeviceEnablePeerAccess; no NCCL, no IPC.
extern "C" int lbfgsb200_comm_create_local(lbfgsb200_comm_t **out, const int *devices, int nranks)
{
    if (!out || !devices || nranks < 1 || nranks > LBFGSB200_MAIL_RANKS) {
        lb::set_error("comm_create_local: bad arguments (%d ranks, at most %d)", nranks, LBFGSB200_MAIL_RANKS);
        return LBFGSB200_ERR_INVALID;
    }
    int saved = 0;
    cudaGetDevice(&saved);
    for (int r = 0; r < nranks; ++r) out[r] = nullptr;
    const size_t mail_bytes = LBFGSB200_MAIL_DOUBLES * sizeof(double);
    int rc = 0;
    std::vector<double *> mails((size_t)nranks, nullptr);
    for (int r = 0; r < nranks && !rc; ++r) {
        lbfgsb200_comm *c = new lbfgsb200_comm;
        memset(c, 0, sizeof *c);
        c->rank = r;
        c->nranks = nranks;
        c->local = 1;
        c->device = devices[r];
        out[r] = c;
        if (cudaSetDevice(devices[r]) != cudaSuccess) { lb::set_error("comm_create_local: cannot select device %d", devices[r]); rc = LBFGSB200_ERR_CUDA; break; }
        for (int q = 0; q < nranks && !rc; ++q) {
            if (q == r || devices[q] == devices[r]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[r], devices[q]) != cudaSuccess || !can) {
                lb::set_error("comm_create_local: device %d cannot access device %d (no NVLink / peer access)", devices[r], devices[q]);
                rc = LBFGSB200_ERR_CUDA;
                break;
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(devices[q], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                lb::set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", devices[r], devices[q], cudaGetErrorString(e));
                rc = LBFGSB200_ERR_CUDA;
            }
            cudaGetLastError();
        }
        if (!rc && (cudaMalloc(&c->mail, mail_bytes) != cudaSuccess || cudaMemset(c->mail, 0, mail_bytes) != cudaSuccess ||
                    cudaMalloc(&c->peers_dev, sizeof(double *) * (size_t)nranks) != cudaSuccess)) {
            lb::set_error("comm_create_local: mailbox allocation failed on device %d", devices[r]);
            rc = LBFGSB200_ERR_NOMEM;
        }
        mails[r] = c->mail;
    }
    for (int r = 0; r < nranks && !rc; ++r) {
        cudaSetDevice(devices[r]);
        if (cudaMemcpy(out[r]->peers_dev, mails.data(), sizeof(double *) * (size_t)nranks, cudaMemcpyHostToDevice) != cudaSuccess) {
            lb::set_error("comm_create_local: upload of the mailbox table failed on device %d", devices[r]);
            rc = LBFGSB200_ERR_CUDA;
        }
        out[r]->p2p = nranks > 1 ? 1 : 0;
    }
    if (rc) {
        for (int r = 0; r < nranks; ++r)
            if (out[r]) { lbfgsb200_comm_destroy(out[r]); out[r] = nullptr; }
    }
    cudaSetDevice(saved);
    return rc;
}

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
    memset(c, 0, sizeof *c);
    c->nccl = comm;
    c->rank = rank;
    c->nranks = nranks;
    cudaGetDevice(&c->device);
    *out = c;
    const char *env = getenv("LBFGSB200_P2P");
    if (nranks > 1 && nranks <= LBFGSB200_MAIL_RANKS && !(env && atoi(env) == 0)) {
        if (lb::setup_mailboxes(c) != 0) { // not fatal, and the same verdict on every rank: keep the NCCL exchange
            c->p2p = 0;
        }
    }
    return 0;
}

extern "C" void lbfgsb200_comm_destroy(lbfgsb200_comm_t *c)
{
    if (!c) return;
    if (c->local) cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < LBFGSB200_MAIL_RANKS; ++r)
        if (c->opened[r]) cudaIpcCloseMemHandle(c->opened[r]);
    if (c->peers_dev) cudaFree(c->peers_dev);
    if (c->mail) cudaFree(c->mail);
    if (c->nccl && lb::g_nccl.handle) lb::g_nccl.CommDestroy((ncclComm_t)c->nccl);
    delete c;
}
