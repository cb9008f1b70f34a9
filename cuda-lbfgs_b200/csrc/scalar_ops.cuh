// scalar_ops.cuh -- the 1-CTA "scalar kernel": everything between two streaming passes.
//
// It (1) sums the per-CTA partials of the pass that just ended in a fixed order (thread t
// adds partials t, t+256, ... then a fixed shuffle/shared-memory tree: deterministic, no
// atomics), or takes the rank-ordered sum of the all-gathered packets on multi-GPU runs, and
// (2) lets one thread do the scalar work the reference does on the host between cuBLAS calls
// (par/L-BFGS.cu:219-272) or inside its line-search loops: rho/alpha/beta of the two-loop
// recursion, gamma, the descent safeguard, the line-search state machine, the curvature gate
// and ring-buffer bookkeeping, the convergence test.  All of it stays in HBM (DevState); the
// host reads nothing back except the 16-byte Ctrl block in host-stepped mode.
#pragma once
#include <cuda_runtime.h>

#include "compact.cuh"
#include "kernels.cuh"
#include "state.h"

namespace lb {

// which boundary values a packet carries besides the sums
enum PackKind : int { PACK_NONE = 0, PACK_X0 = 1, PACK_DIR = 2, PACK_ACCEPT = 3 };
// packet layout (doubles): [0..4] sums, [5] x_first, [6] x_last, [7] g_first, [8] g_last,
// [9] d_first, [10] d_last, [11] spare

// All nq <= kMaxQ quantities at once: thread t adds partials t, t+256, ... of each quantity, one warp
// shuffle tree per quantity, then warp q adds the 8 warp sums of quantity q (the same fixed tree).
__device__ __forceinline__ void reduce_partials(const double *__restrict__ partials, int grid, int nq,
                                                double *out /* shared, >= nq */)
{
    static_assert(kMaxQ <= kScalarThreads / 32, "one finalising warp per quantity");
    __shared__ double sm[kMaxQ][kScalarThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v[kMaxQ];
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) v[q] = 0.0;
    // (the batch of loads of all quantities is issued before the first add: one round trip to L2 per 256 partials)
    for (int i = threadIdx.x; i < grid; i += kScalarThreads) {
        double t[kMaxQ];
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q) t[q] = (q < nq) ? partials[(size_t)q * grid + i] : 0.0;
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q)
            if (q < nq) v[q] += t[q];
    }
#pragma unroll
    for (int q = 0; q < kMaxQ; ++q) {
        if (q < nq) {
            const double w = warp_sum(v[q]);
            if (lane == 0) sm[q][warp] = w;
        }
    }
    __syncthreads();
    if (warp < nq) {
        double w = (lane < kScalarThreads / 32) ? sm[warp][lane] : 0.0;
#pragma unroll
        for (int o = kScalarThreads / 64; o > 0; o >>= 1)
            w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0) out[warp] = w;
    }
    __syncthreads();
}

__device__ __forceinline__ int nq_of(int op)
{
    switch (op) {
    case OP_INIT:
    case OP_ACCEPT: return 5;
    case OP_LS_STEP:
    case OP_F_DIR: return 3; // OP_F_DIR: g.d, f(x + step0 d), grad f(x + step0 d).d
    case OP_SG:
    case OP_L1:
    case OP_L2:
    case OP_COMPACT_DIR: return 1;
    default: return 0;
    }
}

// local sums + this shard's boundary values -> one packet of kPacket doubles
__device__ __forceinline__ void build_packet(const DevState *st, int op, int kind, const double *r, double *s)
{
    for (int q = 0; q < 5; ++q) s[q] = (q < nq_of(op)) ? r[q] : 0.0;
    const long long n = st->n;
    for (int q = 5; q < kPacket; ++q) s[q] = 0.0;
    if (n > 0) {
        if (kind == PACK_X0) {
            s[5] = st->x[0];
            s[6] = st->x[n - 1];
        } else if (kind == PACK_ACCEPT) {
            s[5] = st->x_alt[0]; // the accept kernel wrote the new iterate to x_alt
            s[6] = st->x_alt[n - 1];
            s[7] = st->g[0];
            s[8] = st->g[n - 1];
        } else if (kind == PACK_DIR) {
            s[9] = st->w[0];
            s[10] = st->w[n - 1];
        }
    }
}

// multi-GPU, NCCL exchange: packet -> st->send (the host then all-gathers kPacket doubles per
// rank into st->recv)
__global__ void __launch_bounds__(kScalarThreads) k_pack(DevState *st, int op, int kind, int nparts)
{
    __shared__ double r[kMaxQ];
    reduce_partials(st->partials, nparts, nq_of(op), r);
    if (threadIdx.x != 0) return;
    build_packet(st, op, kind, r, st->send);
}

// ---- multi-GPU, peer-to-peer exchange (comm.h) -------------------------------------------------
// One-way protocol: every 8-byte store carries its own validity tag.  A double travels as two 64-bit cells
// { tag (low 32 bits of the exchange number) << 32 | half of the payload }; 8-byte stores are single NVLink
// transactions, so a receiver that sees the expected tag in a cell sees the payload half with it.  No fence, no
// separate flag, no round trip: an exchange costs one NVLink store latency plus the wait for the slowest rank.
// Two parities of cells alternate; a rank can be at most one exchange ahead of a peer (it needs every peer's
// message of exchange e to finish e), so parity e & 1 is never overwritten before it was read.
__device__ __forceinline__ unsigned long long *mail_cells(double *base, int parity, int sender)
{
    return reinterpret_cast<unsigned long long *>(base) + ((size_t)parity * kMailRanks + sender) * (size_t)(2 * kMailWidth);
}
__device__ __forceinline__ unsigned long long *mail_counter(double *base)
{
    return reinterpret_cast<unsigned long long *>(base) + (size_t)2 * kMailRanks * (size_t)(2 * kMailWidth);
}

struct P2pTicket {
    int par;
    unsigned tag;
};

// Store `count` doubles of src (shared memory) into every rank's mailbox (own included).  Called by every
// thread of the 1-CTA scalar kernel.  The exchange counter lives in the mailbox (per communicator, not per
// solver), so it stays monotonic across solver handles that share a communicator and is identical on every rank.
__device__ P2pTicket p2p_send(DevState *st, const double *src, int count)
{
    __shared__ unsigned long long s_seq;
    __syncthreads(); // src is complete; s_seq of the previous exchange is no longer read
    if (threadIdx.x == 0) {
        unsigned long long *ctr = mail_counter(st->mail);
        s_seq = ++(*ctr);
    }
    __syncthreads();
    P2pTicket t;
    t.par = (int)(s_seq & 1ull);
    t.tag = (unsigned)(s_seq & 0xffffffffull);
    const int P = st->nranks, me = st->rank;
    const unsigned long long hi = (unsigned long long)t.tag << 32;
    for (int idx = threadIdx.x; idx < P * count; idx += kScalarThreads) {
        const int rk = idx / count, i = idx - rk * count;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(src[i]);
        volatile unsigned long long *c = mail_cells(st->peers[rk], t.par, me) + 2 * i;
        c[0] = hi | (bits & 0xffffffffull); // NVLink store into the peer's HBM
        c[1] = hi | (bits >> 32);
    }
    __syncthreads(); // src may be overwritten by the caller from here on
    return t;
}

// Value i of `sender`'s message of the exchange `t`.  Each rank runs on its own GPU, so the bounded spin is a real
// rendezvous; a peer that never arrives fails the launch (trap) after DevState::p2p_timeout_ns (default 120 s,
// LBFGSB200_P2P_TIMEOUT_S) instead of hanging the GPU.
__device__ __forceinline__ double p2p_recv(const DevState *st, P2pTicket t, int sender, int i)
{
    const volatile unsigned long long *c = mail_cells(st->mail, t.par, sender) + 2 * i;
    unsigned long long a = c[0], b = c[1];
    if ((unsigned)(a >> 32) != t.tag || (unsigned)(b >> 32) != t.tag) {
        const unsigned long long t0 = global_ns();
        for (unsigned spin = 1;; ++spin) {
            a = c[0];
            b = c[1];
            if ((unsigned)(a >> 32) == t.tag && (unsigned)(b >> 32) == t.tag) break;
            if ((spin & 1023u) == 0 && global_ns() - t0 > st->p2p_timeout_ns) __trap();
        }
    }
    return __longlong_as_double((long long)((b << 32) | (a & 0xffffffffull)));
}

__device__ __forceinline__ void write_trace(DevState *st)
{
    if (st->trace && st->k < st->trace_rows) {
        double *row = st->trace + (size_t)st->k * LBFGSB200_TRACE_COLS;
        row[0] = (double)st->k;
        row[1] = st->f;
        row[2] = sqrt(st->gg);
        row[3] = st->ls.alpha;
        row[4] = (double)st->ls.trials;
        row[5] = (double)st->h;
        row[6] = st->n > 0 ? st->x[0] : 0.0;
        row[7] = st->n > 0 ? st->x[st->n / 2] : 0.0;
    }
}

// Commit the candidate pair sitting in the spare slot (seq/lbfgs.cpp:182-190: pop the oldest
// when full, push the new pair).
__device__ __forceinline__ void commit_pair(DevState *st, double sy, double yy, double sg)
{
    const int sp = spare_slot(*st);
    st->rho[sp] = 1.0 / sy; // seq/lbfgs.cpp:102 (computed once here, not twice per iteration)
    st->sy[sp] = sy;
    st->yy[sp] = yy;
    st->skip[sp] = 0;
    if (st->h < st->m)
        st->h += 1;
    else
        st->base = (st->base + 1) % st->nslots;
    st->sg = sg;
    st->sg_valid = 1;
}

// One line-search decision on the trial that was just evaluated (f_new, dphi_new at st->ls.alpha).
__device__ __forceinline__ void ls_consume_trial(DevState *st, double f_new, double dphi_new)
{
    st->trial_evals += 1;
    const int cont = ls_step(st->lsp, st->ls, f_new, dphi_new);
    if (!cont) {
        st->ctrl.ls_active = 0;
        // seq/lbfgs.cpp:164-168, par/L-BFGS.cu:295: a step below 1e-10 => give up, keep the OLD x.  The inlined
        // searches give up only if they also did not succeed (par/L-BFGS-Wolfe.cu:353), backtracking never.
        const bool inl = st->lsp.flavor == FLAVOR_PAR_INLINED;
        if (st->ls.alpha < 1e-10 && !(inl && (st->ls.success || st->lsp.kind == LS_BACKTRACKING))) {
            st->status = LBFGSB200_LS_FAILED;
            st->ctrl.done = 1;
        }
    }
}

__device__ void scalar_logic(DevState *st, int op, int p, const double *r)
{
    const bool seq = st->profile == LBFGSB200_PROFILE_SEQ;
    switch (op) {
    case OP_INIT: {
        // x0 evaluation (seq/lbfgs.cpp:28-30): the accept kernel ran with alpha=0 on d=0
        double *t = st->x; st->x = st->x_alt; st->x_alt = t;
        st->f = r[0];
        st->gg = r[1];
        st->f0 = r[0];
        st->gg0 = r[1];
        st->k = 0;
        st->h = 0;
        st->base = 0;
        st->sg_valid = 0;
        st->need_sg = 0;
        st->steepest = 0;
        st->status = LBFGSB200_RUNNING;
        st->trial_evals = 0;
        st->vec_streams = 0.0;
        st->ctrl.ls_active = 0;
        st->ctrl.done = 0;
        st->ctrl.h = 0;
        st->ctrl.k = 0;
        st->ls.alpha = 0.0;
        st->ls.trials = 0;
        break;
    }
    case OP_ITER_BEGIN: {
        if (st->ctrl.done) return;
        if (seq && sqrt(st->gg) < st->tolerance) { // seq/lbfgs.cpp:80-84
            st->status = LBFGSB200_CONVERGED;
            st->ctrl.done = 1;
            return;
        }
        if (st->k >= st->max_iterations) {
            st->status = LBFGSB200_MAX_ITER;
            st->ctrl.done = 1;
            return;
        }
        st->need_sg = 0;
        const int h = st->h;
        int steepest = (st->k == 0 || h == 0); // seq/lbfgs.cpp:87
        if (!steepest) {
            const int newest = slot_of(*st, h - 1);
            if (seq) {
                for (int i = 0; i < h; ++i)
                    if (!isfinite(st->rho[slot_of(*st, i)])) steepest = 1; // :103-108
                const double gamma = st->sy[newest] / st->yy[newest];      // :117
                if (gamma <= 0 || !isfinite(gamma)) steepest = 1;          // :119-124
                st->gamma = gamma;
            } else {
                // par/L-BFGS.cu:241-255
                const double ys = st->sy[newest], yy = st->yy[newest];
                st->gamma = (yy > 0 && ys > 1e-10) ? ys / yy : 1.0;
            }
            if (!steepest && st->direction == LBFGSB200_DIR_TWO_LOOP) {
                if (st->sg_valid) {
                    const double a = st->skip[newest] ? 0.0 : st->rho[newest] * st->sg; // :109
                    st->alpha[h - 1] = a;
                    st->coef = a;
                } else {
                    st->need_sg = 1;
                }
            }
        }
        st->steepest = steepest;
        // bytes model: two-loop = 8h-1 vector streams, d=-g = 2
        if (st->direction == LBFGSB200_DIR_COMPACT)
            st->vec_streams += steepest ? 2.0 : (4.0 * h + 3.0); // pass A (2h+1) + pass B (2h+2)
        else
            st->vec_streams += steepest ? 2.0 : (8.0 * h - 1.0) + (st->need_sg ? 2.0 : 0.0);
        break;
    }
    case OP_SG: {
        if (st->ctrl.done || !st->need_sg) return;
        const int newest = slot_of(*st, st->h - 1);
        st->sg = r[0];
        const double a = st->skip[newest] ? 0.0 : st->rho[newest] * st->sg;
        st->alpha[st->h - 1] = a;
        st->coef = a;
        st->need_sg = 0;
        break;
    }
    case OP_L1: {
        const int h = st->h;
        if (st->ctrl.done || st->steepest || p >= h) return;
        if (p > 0) {
            const int sl = slot_of(*st, p - 1);
            const double a = st->skip[sl] ? 0.0 : st->rho[sl] * r[0]; // alpha_i = rho_i (s_i . q)
            st->alpha[p - 1] = a;
            st->coef = a;
        } else {
            const int sl = slot_of(*st, 0);
            const double beta = st->rho[sl] * r[0]; // seq/lbfgs.cpp:136, r = gamma q
            st->coef = st->skip[sl] ? 0.0 : st->alpha[0] - beta;
        }
        break;
    }
    case OP_L2: {
        const int h = st->h;
        if (st->ctrl.done || st->steepest || p >= h) return;
        if (p < h - 1) {
            const int sl = slot_of(*st, p + 1);
            const double beta = st->rho[sl] * r[0];
            st->coef = st->skip[sl] ? 0.0 : st->alpha[p + 1] - beta; // seq/lbfgs.cpp:139
        } else {
            st->gd = r[0]; // seq/lbfgs.cpp:146
            if (seq && st->gd >= 0) { // :147-153, resolved by k_steepest + OP_LS_INIT
                st->steepest = 1;
                st->vec_streams += 2.0;
            }
        }
        break;
    }
    case OP_LS_INIT: {
        if (st->ctrl.done) return;
        if (st->steepest) {
            // d = -g  =>  g.d = -(g.g) with the same summation order (negation is exact)
            st->gd = -st->gg;
            st->dL = -st->gL;
            st->dR = -st->gR;
        }
        ls_begin(st->lsp, st->ls, st->f, st->gd, st->f0);
        st->ctrl.ls_active = 1;
        break;
    }
    case OP_LS_STEP: {
        if (st->ctrl.done || !st->ctrl.ls_active) return;
        st->pend_steepest = 0; // the trial that just ran rewrote d = -g if that was pending
        st->vec_streams += 2.0; // trial: reads x, d
        ls_consume_trial(st, r[0], r[1]);
        break;
    }
    case OP_F_DIR: {
        // fused compact flow, after k_combine_trial: r = { g.d, f(x + step0 d), grad f(x + step0 d).d }
        if (st->ctrl.done) return;
        const int h = st->h;
        st->vec_streams += (st->steepest || h == 0) ? 3.0 : (2.0 * h + 3.0); // reads the basis (or g) and x, writes d
        st->gd = r[0];
        if (st->steepest || h == 0) {
            st->gd = -st->gg; // d = -g  =>  g.d = -(g.g) with the same summation order (negation is exact)
        } else if (seq && st->gd >= 0) {
            // seq/lbfgs.cpp:147-153: not a descent direction => d = -g.  The direction in memory and the fused
            // first trial are void: the next k_trial reads g, uses and stores d = -g (kernels.cuh, PEND).
            st->steepest = 1;
            st->pend_steepest = 1;
            st->gd = -st->gg;
            st->dL = -st->gL;
            st->dR = -st->gR;
            ls_begin(st->lsp, st->ls, st->f, st->gd, st->f0);
            st->ctrl.ls_active = 1;
            st->vec_streams += 1.0; // the trial that follows also writes d (it is counted as 2 like any other)
            break;
        }
        ls_begin(st->lsp, st->ls, st->f, st->gd, st->f0);
        st->ctrl.ls_active = 1;
        ls_consume_trial(st, r[1], r[2]); // the first trial (alpha = step0) was evaluated by the combine pass
        break;
    }
    case OP_ACCEPT: {
        if (st->ctrl.done) return;
        double *t = st->x; st->x = st->x_alt; st->x_alt = t; // x <- x_new
        st->f = r[0];
        st->gg = r[1];
        const double sy = r[2], yy = r[3], sg = r[4];
        if (seq) {
            if (sy > 0) commit_pair(st, sy, yy, sg); // seq/lbfgs.cpp:181-190
            else st->sg_valid = 0;                   // :192-195 "Skipping update"
        } else {
            commit_pair(st, sy, yy, sg);             // par/L-BFGS.cu:332-333: always overwritten
            if (sy <= 1e-10) {                       // par/L-BFGS.cu:222-223
                const int newest = slot_of(*st, st->h - 1);
                st->skip[newest] = 1;
                st->rho[newest] = 0.0;
            }
        }
        st->vec_streams += 7.0; // accept: reads x, d, g ; writes x, g, s, y
        write_trace(st);
        st->k += 1;
        st->iters_left -= 1;
        if (!seq && sqrt(st->gg) <= st->tolerance) { // par/L-BFGS.cu:353-357
            st->status = LBFGSB200_CONVERGED;
            st->ctrl.done = 1;
        } else if (st->k >= st->max_iterations) {    // seq/lbfgs.cpp:201
            st->status = LBFGSB200_MAX_ITER;
            st->ctrl.done = 1;
        }
        st->ctrl.h = st->h;
        st->ctrl.k = st->k;
        break;
    }
    case OP_COMPACT_DIR: {
        if (st->ctrl.done || st->steepest || st->h == 0) return;
        st->gd = r[0];
        if (seq && st->gd >= 0) { // seq/lbfgs.cpp:147-153
            st->steepest = 1;
            st->vec_streams += 2.0;
        }
        break;
    }
    default: break;
    }
}


// device-side control flow of graph mode: the iteration loop and the trial loop are WHILE nodes, the stand-alone
// pass A of the fused flow an IF node (thread 0 only; legal only inside the graph, hence the use_graph gate)
__device__ __forceinline__ void set_conditions(const DevState *st, int op)
{
    if (!st->use_graph) return;
    const unsigned run = (!st->ctrl.done && st->iters_left > 0) ? 1u : 0u;
    if (op == OP_LS_INIT || op == OP_LS_STEP || op == OP_F_DIR)
        cudaGraphSetConditional(st->cond_inner, (st->ctrl.ls_active && !st->ctrl.done) ? 1u : 0u);
    if (op == OP_ACCEPT || op == OP_ITER_BEGIN || op == OP_LS_STEP || op == OP_F_ACCEPT || op == OP_F_DIR ||
        op == OP_F_FIX || op == OP_F_BEGIN)
        cudaGraphSetConditional(st->cond_outer, run);
    if (st->fused && (op == OP_F_ACCEPT || op == OP_F_FIX || op == OP_F_BEGIN))
        cudaGraphSetConditional(st->cond_fix, (run && st->ctrl.need_fix) ? 1u : 0u);
}

// ---- fused compact flow: everything between k_accept_gram and the next k_combine_trial -----------------------
// op = OP_F_INIT (x0 evaluation) or OP_F_ACCEPT.  CTA-wide.  partials of k_accept_gram: rows q = j*3 + r for the
// J' = 2h'+1 columns of the anticipated new window (r = 0: s_new, 1: y_new, 2: g_new), then q = 3J': f.
//   1. sum the partials; multi-GPU: ONE exchange per accept carrying the rows, f and this shard's boundary x / g
//   2. accept bookkeeping (seq/lbfgs.cpp:159-199, par/L-BFGS.cu:309-357): x <- x_new, curvature gate, ring commit,
//      convergence / iteration-limit tests, trace row
//   3. next direction: Gram update, the two loops of seq/lbfgs.cpp:93-143 on coefficients, neighbours' boundary d
__device__ void fused_accept(DevState *st, int op, int from_comm, int nparts, double *dyn)
{
    __shared__ double rows[3 * kMaxCols + 2 * kRowsExtra];
    __shared__ double edge[4]; // new boundary values of the neighbours: x_last(left), x_first(right), g_last(left), g_first(right)
    __shared__ int s_flags[3]; // fresh (pair committed), prepare the next direction here, remap columns
    const bool init = (op == OP_F_INIT);
    if (!init && st->ctrl.done) return;
    const bool seq = st->profile == LBFGSB200_PROFILE_SEQ;
    const int h_old = init ? 0 : st->h;
    const int ks = (h_old == st->m) ? 1 : 0, hk = h_old - ks, hp = hk + 1, Jp = 2 * hp + 1;
    const int nrows = 3 * Jp + 1;
    const int P = st->nranks, left = st->rank - 1, right = st->rank + 1;
    if (from_comm == 1) {
        // NCCL exchange: k_pack_rows + all-gather ran before this kernel
        const int cnt = st->gram_count;
        for (int q = threadIdx.x; q < nrows; q += kScalarThreads) {
            double v = 0.0;
            for (int k = 0; k < P; ++k) v += st->gram_recv[(size_t)k * cnt + q];
            rows[q] = v;
        }
        if (threadIdx.x == 0) {
            edge[0] = left >= 0 ? st->gram_recv[(size_t)left * cnt + nrows + 1] : 0.0;
            edge[1] = right < P ? st->gram_recv[(size_t)right * cnt + nrows + 0] : 0.0;
            edge[2] = left >= 0 ? st->gram_recv[(size_t)left * cnt + nrows + 3] : 0.0;
            edge[3] = right < P ? st->gram_recv[(size_t)right * cnt + nrows + 2] : 0.0;
        }
        __syncthreads();
    } else {
        gram_rows_from_partials(st->partials, nparts, nrows, nrows + 4, rows);
        tl_mark(st, 101, global_ns()); // (diagnostic timeline: partial sums done)
        if (from_comm == 2) {
            if (threadIdx.x == 0 && st->n > 0) { // the accept kernel wrote the new iterate to x_alt
                rows[nrows + 0] = st->x_alt[0];
                rows[nrows + 1] = st->x_alt[st->n - 1];
                rows[nrows + 2] = st->g[0];
                rows[nrows + 3] = st->g[st->n - 1];
            }
            const P2pTicket t = p2p_send(st, rows, nrows + 4);
            for (int q = threadIdx.x; q < nrows; q += kScalarThreads) {
                double v = 0.0;
                for (int k = 0; k < P; ++k) v += p2p_recv(st, t, k, q); // rank order: identical bits on every rank
                rows[q] = v;
            }
            if (threadIdx.x == 32) {
                edge[0] = left >= 0 ? p2p_recv(st, t, left, nrows + 1) : 0.0;
                edge[1] = right < P ? p2p_recv(st, t, right, nrows + 0) : 0.0;
                edge[2] = left >= 0 ? p2p_recv(st, t, left, nrows + 3) : 0.0;
                edge[3] = right < P ? p2p_recv(st, t, right, nrows + 2) : 0.0;
            }
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        double *t = st->x; st->x = st->x_alt; st->x_alt = t; // x <- x_new
        const double f = rows[3 * Jp], gg = rows[(2 * hp) * 3 + 2];
        const double sy = rows[(2 * hp - 1) * 3 + 0], yy = rows[(2 * hp - 1) * 3 + 1], sg = rows[(2 * hp) * 3 + 0];
        st->f = f;
        st->gg = gg;
        const int ns = st->nslots, sp = init ? 0 : spare_slot(*st);
        if (P > 1) { // halo: the neighbours' boundary x, g, and (as differences of their own operands) s, y
            if (!init) {
                st->bL[sp] = edge[0] - st->xL;
                st->bR[sp] = edge[1] - st->xR;
                st->bL[ns + sp] = edge[2] - st->gL;
                st->bR[ns + sp] = edge[3] - st->gR;
            }
            st->xL = edge[0]; st->xR = edge[1]; st->gL = edge[2]; st->gR = edge[3];
            st->bL[2 * ns] = edge[2];
            st->bR[2 * ns] = edge[3];
        }
        int committed = 0;
        if (init) { // seq/lbfgs.cpp:28-30
            st->f0 = f;
            st->gg0 = gg;
            st->k = 0;
            st->h = 0;
            st->base = 0;
            st->sg_valid = 0;
            st->need_sg = 0;
            st->status = LBFGSB200_RUNNING;
            st->trial_evals = 0;
            st->vec_streams = 0.0;
            st->ctrl.ls_active = 0;
            st->ctrl.done = 0;
            st->ls.alpha = 0.0;
            st->ls.trials = 0;
            st->pend_steepest = 0;
            if (st->max_iterations <= 0) {
                st->status = LBFGSB200_MAX_ITER;
                st->ctrl.done = 1;
            }
        } else {
            if (seq) {
                if (sy > 0) { commit_pair(st, sy, yy, sg); committed = 1; } // seq/lbfgs.cpp:181-190
                else st->sg_valid = 0;                                       // :192-195 "Skipping update"
            } else {
                commit_pair(st, sy, yy, sg); // par/L-BFGS.cu:332-333: always overwritten
                committed = 1;
                if (sy <= 1e-10) {           // par/L-BFGS.cu:222-223
                    const int newest = slot_of(*st, st->h - 1);
                    st->skip[newest] = 1;
                    st->rho[newest] = 0.0;
                }
            }
            st->vec_streams += 2.0 * hk + 7.0; // reads the kept history, x, d, g ; writes x, g, s, y
            write_trace(st);
            st->k += 1;
            st->iters_left -= 1;
            if (!seq && sqrt(gg) <= st->tolerance) { // par/L-BFGS.cu:353-357
                st->status = LBFGSB200_CONVERGED;
                st->ctrl.done = 1;
            } else if (st->k >= st->max_iterations) { // seq/lbfgs.cpp:201
                st->status = LBFGSB200_MAX_ITER;
                st->ctrl.done = 1;
            }
        }
        // the test at the top of the next iteration (seq/lbfgs.cpp:80-84), made as soon as |g| is known
        if (!st->ctrl.done && seq && sqrt(gg) < st->tolerance) {
            st->status = LBFGSB200_CONVERGED;
            st->ctrl.done = 1;
        }
        const int h = st->h;
        st->steepest = (h == 0); // seq/lbfgs.cpp:87
        // a rejected pair: with a full ring the rows of g against the oldest pair (evicted in anticipation) are
        // missing -> stand-alone pass A; otherwise only the columns of s_new / y_new have to be dropped
        const int need_fix = (!init && !committed && ks == 1 && !st->ctrl.done) ? 1 : 0;
        st->ctrl.need_fix = need_fix;
        st->ctrl.h = h;
        st->ctrl.k = st->k;
        s_flags[0] = committed;
        s_flags[1] = (!st->ctrl.done && h > 0 && !need_fix) ? 1 : 0;
        s_flags[2] = (!init && !committed && ks == 0) ? 1 : 0;
        if (s_flags[1] && s_flags[2]) {
            // columns of the unchanged window (h = hk pairs): S j <- fused column j, Y j <- hk+1+j, g <- 2hk+2
            for (int j = h; j < 2 * h + 1; ++j) {
                const int src = (j < 2 * h) ? j + 1 : 2 * h + 2;
                for (int r = 0; r < 3; ++r) rows[j * 3 + r] = rows[src * 3 + r];
            }
        }
    }
    __syncthreads();
    tl_mark(st, 102, global_ns()); // (bookkeeping done)
    if (s_flags[1]) {
        compact_recursion(st, rows, dyn, s_flags[0], true);
        __syncthreads();
    }
    tl_mark(st, 103, global_ns()); // (Gram update + recursion done)
    if (threadIdx.x == 0) {
        if (st->steepest) { // d = -g: the neighbours' boundary d follows from their boundary g
            st->dL = -st->gL;
            st->dR = -st->gR;
        }
        set_conditions(st, OP_F_ACCEPT);
    }
}

// (rare) the stand-alone pass A re-computed the rows of the CURRENT window after a rejected pair: op = OP_F_FIX
__device__ void fused_fix(DevState *st, int from_comm, int nparts, double *dyn)
{
    __shared__ double rows[3 * kMaxCols + 2 * kRowsExtra];
    if (st->ctrl.done || !st->ctrl.need_fix) return;
    const int h = st->h, nrows = 3 * (2 * h + 1), P = st->nranks;
    if (from_comm == 1) {
        const int cnt = st->gram_count;
        for (int q = threadIdx.x; q < nrows; q += kScalarThreads) {
            double v = 0.0;
            for (int k = 0; k < P; ++k) v += st->gram_recv[(size_t)k * cnt + q];
            rows[q] = v;
        }
        __syncthreads();
    } else {
        gram_rows_from_partials(st->partials, nparts, nrows, nrows, rows);
        if (from_comm == 2) {
            const P2pTicket t = p2p_send(st, rows, nrows);
            for (int q = threadIdx.x; q < nrows; q += kScalarThreads) {
                double v = 0.0;
                for (int k = 0; k < P; ++k) v += p2p_recv(st, t, k, q);
                rows[q] = v;
            }
            __syncthreads();
        }
    }
    compact_recursion(st, rows, dyn, 0, true);
    __syncthreads();
    if (threadIdx.x == 0) {
        st->vec_streams += 2.0 * h + 1.0;
        st->ctrl.need_fix = 0;
        if (st->steepest) {
            st->dL = -st->gL;
            st->dR = -st->gR;
        }
        set_conditions(st, OP_F_FIX);
    }
}

// multi-GPU, NCCL exchange of the fused flow: the rows / f / boundary values of k_accept_gram (or the rows of the
// stand-alone pass A) -> st->gram_rows, which the host all-gathers into st->gram_recv
__global__ void __launch_bounds__(kScalarThreads) k_pack_rows(DevState *st, int op, int nparts)
{
    __shared__ double rows[3 * kMaxCols + 2 * kRowsExtra];
    const bool init = (op == OP_F_INIT);
    if (!init && st->ctrl.done) return;
    int nrows, extra = 0;
    if (op == OP_F_FIX) {
        if (!st->ctrl.need_fix) return;
        nrows = 3 * (2 * st->h + 1);
    } else {
        const int h_old = init ? 0 : st->h;
        const int hp = h_old - ((h_old == st->m) ? 1 : 0) + 1;
        nrows = 3 * (2 * hp + 1) + 1;
        extra = 4;
    }
    gram_rows_from_partials(st->partials, nparts, nrows, st->gram_count, rows);
    if (extra && threadIdx.x == 0 && st->n > 0) {
        rows[nrows + 0] = st->x_alt[0];
        rows[nrows + 1] = st->x_alt[st->n - 1];
        rows[nrows + 2] = st->g[0];
        rows[nrows + 3] = st->g[st->n - 1];
    }
    __syncthreads();
    for (int q = threadIdx.x; q < st->gram_count; q += kScalarThreads) st->gram_rows[q] = rows[q];
}

// st is the SHARED-MEMORY copy of the solver state (see k_scalar); nparts is the partial count of the pass
// that just ended; dyn is the kernel's dynamic shared memory (compact ops: the window Gram matrix).
__device__ void scalar_body(DevState *st, int op, int p, int from_comm, int pack_kind, int nparts,
                            unsigned long long t_in, double *dyn)
{
    __shared__ double r[kMaxQ];
    __shared__ double pk[kPacket];
    __shared__ double rbuf[kMaxRanks * kPacket];
    if (op == OP_F_INIT || op == OP_F_ACCEPT) {
        fused_accept(st, op, from_comm, nparts, dyn);
        tl_mark(st, op, t_in);
        return;
    }
    if (op == OP_F_FIX) {
        fused_fix(st, from_comm, nparts, dyn);
        tl_mark(st, op, t_in);
        return;
    }
    if (op == OP_F_BEGIN) { // graph prologue
        if (threadIdx.x == 0) set_conditions(st, op);
        return;
    }
    if (op == OP_COMPACT) { // CTA-wide: pass-A sums (+ exchange) -> Gram update + coefficient recursion (compact.cuh)
        if (st->ctrl.done || st->h == 0 || (st->steepest && !st->sg_valid)) return;
        __shared__ double rows[3 * kMaxCols + 2 * kRowsExtra];
        const int cnt = st->gram_count, nrows = 3 * (2 * st->h + 1);
        if (from_comm == 1) {
            // NCCL path: k_gram_finalize + all-gather ran before this kernel; rank-ordered sum
            for (int q = threadIdx.x; q < nrows; q += kScalarThreads) {
                double v = 0.0;
                for (int k = 0; k < st->nranks; ++k) v += st->gram_recv[(size_t)k * cnt + q];
                rows[q] = v;
            }
            __syncthreads();
        } else {
            gram_rows_from_partials(st->partials, nparts, nrows, nrows, rows);
            if (from_comm == 2) {
                // peer-to-peer: all-gather the pass-A rows and add them in rank order
                const P2pTicket t = p2p_send(st, rows, nrows);
                for (int q = threadIdx.x; q < nrows; q += kScalarThreads) {
                    double v = 0.0;
                    for (int k = 0; k < st->nranks; ++k) v += p2p_recv(st, t, k, q);
                    rows[q] = v;
                }
                __syncthreads();
            }
        }
        // a direction that was forced to d = -g still needs the rows of the freshly committed pair stored
        compact_recursion(st, rows, dyn, st->sg_valid, !st->steepest);
        __syncthreads();
        tl_mark(st, op, t_in);
        return;
    }
    const int nq = nq_of(op);
    const double *rv = nullptr; // gathered packets, [rank][stride]
    int rv_stride = kPacket;
    if (!from_comm) {
        reduce_partials(st->partials, nparts, nq_of(op), r);
    } else if (from_comm == 2) {
        // peer-to-peer: pack + exchange inside this kernel
        reduce_partials(st->partials, nparts, nq_of(op), r);
        if (threadIdx.x == 0) build_packet(st, op, pack_kind, r, pk);
        const int count = (pack_kind == PACK_NONE) ? (nq > 0 ? nq : 1) : kPacket; // sums only, or sums + boundary values
        const P2pTicket t = p2p_send(st, pk, count);
        for (int idx = threadIdx.x; idx < st->nranks * count; idx += kScalarThreads) {
            const int rk = idx / count, i = idx - rk * count;
            rbuf[rk * kPacket + i] = p2p_recv(st, t, rk, i);
        }
        __syncthreads();
        rv = rbuf;
    } else {
        rv = st->recv;
    }
    if (rv && threadIdx.x == 0) {
        for (int q = 0; q < nq; ++q) {
            double t = 0.0;
            for (int k = 0; k < st->nranks; ++k) t += rv[(size_t)k * rv_stride + q];
            r[q] = t;
        }
        const int left = st->rank - 1, right = st->rank + 1;
        // (a finished solver keeps its halo: an idle accept segment must not replace it with stale boundary values)
        if ((pack_kind == PACK_X0 || pack_kind == PACK_ACCEPT) && !(st->ctrl.done && pack_kind == PACK_ACCEPT)) {
            st->xL = left >= 0 ? rv[(size_t)left * rv_stride + 6] : 0.0;
            st->xR = right < st->nranks ? rv[(size_t)right * rv_stride + 5] : 0.0;
            st->gL = left >= 0 ? rv[(size_t)left * rv_stride + 8] : 0.0;
            st->gR = right < st->nranks ? rv[(size_t)right * rv_stride + 7] : 0.0;
        } else if (pack_kind == PACK_DIR && !st->ctrl.done) {
            st->dL = left >= 0 ? rv[(size_t)left * rv_stride + 10] : 0.0;
            st->dR = right < st->nranks ? rv[(size_t)right * rv_stride + 9] : 0.0;
        }
    }
    if (threadIdx.x == 0) {
        scalar_logic(st, op, p, r);
        set_conditions(st, op);
        tl_mark(st, op, t_in);
    }
}

// from_comm = 0: single GPU, sums come straight from the partials of the last pass.
// from_comm = 1: sums are the rank-ordered totals of the all-gathered packets (identical
//                bits on every rank); neighbours' boundary values are picked up as halo.
// from_comm = 2: as 1, with the exchange done inside this kernel through the peer mailboxes.
// The solver state (~5 KB) is staged in shared memory for the duration of the kernel and written back at
// the end: the scalar logic is a long chain of dependent reads and writes of that state by ONE thread,
// ~0.7 us per link in HBM/L2, ~30 ns in shared memory.  No other kernel runs concurrently on the state
// (stream order), so the copy is exclusive.
__global__ void __launch_bounds__(kScalarThreads)
k_scalar(DevState *gst, int op, int p, int from_comm, int pack_kind, int nparts)
{
    extern __shared__ __align__(16) double dyn[];
    __shared__ __align__(16) unsigned long long sbuf[(sizeof(DevState) + 7) / 8];
    static_assert(sizeof(DevState) % 8 == 0, "DevState is copied in 8-byte words");
    const unsigned long long t_in = global_ns();
    const unsigned long long *src = reinterpret_cast<const unsigned long long *>(gst);
    for (int i = threadIdx.x; i < (int)(sizeof(DevState) / 8); i += kScalarThreads) sbuf[i] = src[i];
    __syncthreads();
    scalar_body(reinterpret_cast<DevState *>(sbuf), op, p, from_comm, pack_kind, nparts, t_in, dyn);
    __syncthreads();
    unsigned long long *dst = reinterpret_cast<unsigned long long *>(gst);
    for (int i = threadIdx.x; i < (int)(sizeof(DevState) / 8); i += kScalarThreads) dst[i] = sbuf[i];
}

// unit-test surface: finalise nq partial sums into d_out[0..nq)
__global__ void __launch_bounds__(kScalarThreads)
k_finalize(const double *partials, int grid, int nq, double *d_out, int take_sqrt)
{
    __shared__ double r[kMaxQ];
    reduce_partials(partials, grid, nq, r);
    if (threadIdx.x == 0)
        for (int q = 0; q < nq; ++q) d_out[q] = take_sqrt ? sqrt(r[q]) : r[q];
}

} // namespace lb
