// state.h -- device-resident solver state shared by every kernel.
//
// The whole iteration is driven from this struct: ring-buffer slot arithmetic, the
// rho/alpha/beta scalars of the two-loop recursion, the line-search state machine and the
// exit flags all live in HBM, are written by the 1-CTA scalar kernel (scalar_ops.cuh) and
// read by the streaming kernels (kernels.cuh).  The host never needs a scalar to decide
// what to launch next except the 16-byte `ctrl` block in host-stepped mode.
#pragma once
#include <stdint.h>

#include "../../include/lbfgsb200.h"
#include "ls_logic.h"

namespace lb {

constexpr int kMaxM = LBFGSB200_MAX_M;
constexpr int kMaxSlots = kMaxM + 1; // one spare slot: the candidate pair is written there and
                                     // only committed when the profile's curvature gate passes
constexpr int kThreads = 256;        // streaming kernels: 256 threads, 4 CTAs / SM
constexpr int kCtasPerSm = 4;
constexpr int kCtasPerSmAccept = 2;  // the 7-stream accept kernel trades occupancy for registers
constexpr int kUnroll = 4;           // double2 items per thread per tile
constexpr int kTileVec = kThreads * kUnroll; // 1024 double2 = 16 KB per stream per tile
constexpr int kMaxQ = 8;             // partial sums one kernel may emit
constexpr int kPacket = 12;          // doubles per rank in the packed per-step exchange
constexpr int kMaxRanks = 16;
constexpr int kScalarThreads = 256;   // the 1-CTA scalar kernel
constexpr int kMailRanks = 16;        // == LBFGSB200_MAIL_RANKS (comm.h)
constexpr int kMailWidth = 320;       // == LBFGSB200_MAIL_WIDTH: >= kPacket and >= 3*(2*50+1) Gram rows + kRowsExtra
constexpr int kRowsExtra = 8;         // fused accept exchange: f, x_first, x_last, g_first, g_last (+3 spare) after the rows

// scalar-kernel opcodes
enum Op : int {
    OP_INIT = 0,     // after the x0 evaluation
    OP_ITER_BEGIN,   // convergence test, steepest-descent decision, first alpha
    OP_SG,           // (rare) finalise s_newest . g when the last pair was not committed
    OP_L1,           // after loop-1 pass p
    OP_L2,           // after loop-2 pass p
    OP_LS_INIT,      // descent safeguard + line-search start
    OP_LS_STEP,      // one line-search decision
    OP_ACCEPT,       // after the accept/update kernel
    OP_COMPACT,      // compact form: Gram update + coefficient recursion
    OP_COMPACT_DIR,  // after the compact direction kernel (g.d)
    // ---- fused compact flow (accept_gram.cuh): 2 + (t-1) scalar kernels per iteration ----
    OP_F_INIT,       // after k_accept_gram(init): f, g.g at x0
    OP_F_ACCEPT,     // after k_accept_gram: accept bookkeeping, curvature gate, convergence test, Gram update,
                     //   coefficient recursion for the NEXT direction, neighbours' d boundary values
    OP_F_DIR,        // after k_combine_trial: g.d, descent safeguard, line-search start + its first decision
    OP_F_FIX,        // (rare) after the stand-alone pass A: the rejected pair left rows of g against the oldest pair missing
    OP_F_BEGIN,      // graph prologue: arm the WHILE / IF conditions from the state
};

// 16-byte block the host reads back once per trial in host-stepped mode
struct Ctrl {
    int ls_active;
    int done;
    int h;
    int k;
    int need_fix; // fused flow: the stand-alone pass A has to run before the next direction
    int pad;
};

struct DevState {
    // ---- immutable after create ----
    long long n;      // local elements
    long long goff;   // global index of local element 0
    long long nglob;  // global problem size
    int m, nslots;
    int objective, profile, direction;
    int max_iterations;
    double tolerance;
    int rank, nranks;
    int grid;         // CTAs of the 4-per-SM streaming kernels (partial count is passed per launch)
    int grid_accept;  // CTAs of the accept kernel
    double *arena0;            // row 0 of the arena = x_a (tensor-map row arithmetic: x is row 0 or row 3)
    double *x, *x_alt, *g, *w; // w: two-loop work vector q/r, ends as the direction d;
                               // x_alt: the accept kernel writes the new iterate here, then x <-> x_alt
    double *S, *Y;     // ring buffers, nslots rows of `stride` doubles
    long long stride;
    double *partials;  // [kMaxQ][grid]
    double *send, *recv; // multi-GPU packet buffers: [kPacket], [nranks][kPacket]
    double *trace;
    long long trace_rows;
    LsParams lsp;

    // ---- ring ----
    int base; // physical slot of the oldest pair
    int h;    // stored pairs
    double rho[kMaxSlots]; // per physical slot: 1/(s.y)
    double sy[kMaxSlots];
    double yy[kMaxSlots];
    unsigned char skip[kMaxSlots]; // CUDA profile: pair excluded (s.y <= 1e-10)

    // ---- direction ----
    double alpha[kMaxM]; // per window position (0 = oldest)
    double coef;         // coefficient of the next streaming pass
    double gamma;
    double sg;           // s_newest . g (from the accept kernel when the pair was committed)
    int sg_valid;
    int need_sg;
    int steepest;

    // ---- iterate ----
    double f, gg, gd;
    double f0, gg0; // at x0
    int k;
    int status;
    Ctrl ctrl;
    long long trial_evals;

    // ---- line search ----
    LsState ls;

    // ---- halo (multi-GPU, neighbour-coupled objectives) ----
    // boundary values of the neighbours' shards, refreshed once per outer iteration
    double xL, xR, dL, dR, gL, gR;

    // neighbours' boundary elements of EVERY basis vector (fused compact flow): index = basis index
    // (S slot s -> s, Y slot s -> nslots + s, g -> 2 nslots); bL: LAST element of the left neighbour's shard,
    // bR: FIRST element of the right neighbour's.  Kept up to date from the accept packets (s = x_new - x_old and
    // y = g_new - g_old are formed here with the neighbour's own operands), so that every rank can form its
    // neighbours' boundary d = -sum_j delta_j b_j with the fma chain of k_combine_trial -- bit-identical to what
    // the neighbour computes -- BEFORE the combine pass runs; the pass can then evaluate the first trial itself.
    double bL[2 * kMaxSlots + 1], bR[2 * kMaxSlots + 1];

    // ---- peer-to-peer mailbox exchange (multi-GPU; see comm.h) ----
    double *mail;    // this rank's mailbox
    double **peers;  // [nranks] mailbox pointers of all ranks (own entry == mail)
    int p2p;
    unsigned long long p2p_timeout_ns; // bounded rendezvous: trap instead of hanging the GPU

    // ---- CUDA-graph mode: WHILE-node condition handles set by the scalar kernel ----
    unsigned long long cond_outer, cond_inner, cond_fix;
    int use_graph;
    // ---- fused compact flow ----
    int fused;          // 1: k_accept_gram / k_combine_trial flow
    int pend_steepest;  // the descent safeguard fired AFTER the combine pass: the next k_trial rewrites d = -g itself
    int tl_sub;         // diagnostic timeline: also record the sub-marks inside OP_F_ACCEPT (each costs ~1 us: LBFGSB200_TIMELINE_SUB)
    long long iters_left; // iteration budget of the current iterate() call

    // ---- accounting: algorithmic HBM traffic in units of one local vector (8 n bytes) ----
    double vec_streams;

    // ---- compact form (separate allocations; see compact.cuh) ----
    double *gram;      // Gram matrix of the basis [s slots, y slots, g], (2*nslots+1)^2
    double *gram_rows; // pass-A output: 3 x (2h+1) inner products, window-column order
    double *gram_recv; // multi-GPU: all-gathered gram_rows, [nranks][gram_count]
    int gram_count;    // doubles per rank in that exchange: 3 * (2m+1)
    double *delta;     // direction coefficients by window column

    // ---- diagnostic timeline (LBFGSB200_TIMELINE=rows): (op, %globaltimer at entry, at exit) of every
    //      scalar kernel, so that the gaps between the vector kernels can be read off the device ----
    unsigned long long *tl;
    int tl_cap, tl_n;
};

__host__ __device__ inline int slot_of(const DevState &st, int pos)
{
    return (st.base + pos) % st.nslots;
}
__host__ __device__ inline int spare_slot(const DevState &st)
{
    return (st.base + st.h) % st.nslots;
}

} // namespace lb
