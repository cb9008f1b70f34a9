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

// Peer-to-peer mailbox (NVLink / NVSwitch): every rank owns one device buffer of 64-bit cells
//   cells   [2 parities][LBFGSB200_MAIL_RANKS senders][LBFGSB200_MAIL_WIDTH doubles][2 halves]
//   counter [1] (+ padding): the exchange number of this communicator
// that all peers can store into (CUDA IPC mappings between processes, cudaDeviceEnablePeerAccess inside one
// process).  A cell = { low 32 bits of the exchange number } << 32 | one 32-bit half of a double: the scalar
// kernel of rank r stores its message straight into every peer's mailbox and polls the cells of its own mailbox
// until they carry the current exchange number -- pack + exchange + scalar logic in ONE kernel, one-way latency, no
// fence + flag round trip (scalar_ops.cuh).  NCCL is used to bootstrap multi-process communicators (exchange of
// the IPC handles) and remains the exchange when peer access is unavailable or LBFGSB200_P2P=0.
#define LBFGSB200_MAIL_RANKS 16
#define LBFGSB200_MAIL_WIDTH 320
#define LBFGSB200_MAIL_DOUBLES ((size_t)2 * LBFGSB200_MAIL_RANKS * LBFGSB200_MAIL_WIDTH * 2 + 16)

struct lbfgsb200_comm {
    void *nccl; // ncclComm_t (NULL for an in-process communicator)
    int rank;
    int nranks;
    int p2p;            // 1: mailboxes mapped on every rank
    int local;          // 1: created by lbfgsb200_comm_create_local (peer access, no IPC, no NCCL)
    int device;
    double *mail;       // this rank's mailbox (device memory)
    double **peers_dev; // device array [nranks] of mailbox pointers (own entry = mail)
    void *opened[LBFGSB200_MAIL_RANKS]; // IPC mappings to close
};

namespace lb {
// all-gather `count` doubles per rank: recv[r*count + i] = send_r[i].  Stream-ordered; safe to
// capture in a CUDA graph.  Returns 0 or LBFGSB200_ERR_NCCL.
int comm_allgather(lbfgsb200_comm *c, const double *send, double *recv, int count,
                   cudaStream_t stream);
} // namespace lb
