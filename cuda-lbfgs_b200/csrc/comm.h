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

// Peer-to-peer mailbox (NVLink / NVSwitch): every rank owns one device buffer
//   data  [2 parities][kMailRanks senders][kMailWidth doubles]
//   flags [2 parities][kMailRanks senders] (64-bit sequence numbers)
// that all peers map through CUDA IPC.  The scalar kernel of rank r stores its packet straight
// into every peer's mailbox, fences, raises its flag there, and spins (bounded) on the flags in
// its own mailbox -- pack + exchange + scalar logic in ONE kernel instead of three kernels plus
// an NCCL launch.  NCCL is still used to bootstrap (exchange of the IPC handles) and remains the
// fallback exchange when peer access is unavailable or LBFGSB200_P2P=0.
#define LBFGSB200_MAIL_RANKS 16
#define LBFGSB200_MAIL_WIDTH 304

struct lbfgsb200_comm {
    void *nccl; // ncclComm_t
    int rank;
    int nranks;
    int p2p;            // 1: mailboxes mapped on every rank
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
