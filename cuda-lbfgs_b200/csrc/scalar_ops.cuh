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
    for (int i = threadIdx.x; i < grid; i += kScalarThreads) {
#pragma unroll
        for (int q = 0; q < kMaxQ; ++q)
            if (q < nq) v[q] += partials[(size_t)q * grid + i];
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
    case OP_LS_STEP: return 3;
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
__device__ __forceinline__ double *mail_slot(double *base, int parity, int sender)
{
    return base + ((size_t)parity * kMailRanks + sender) * kMailWidth;
}
__device__ __forceinline__ volatile unsigned long long *mail_flag(double *base, int parity, int sender)
{
    return reinterpret_cast<volatile unsigned long long *>(base + (size_t)2 * kMailRanks * kMailWidth) +
           parity * kMailRanks + sender;
}
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// All-gather `count` doubles per rank through the mailboxes.  Called by every thread of the
// 1-CTA scalar kernel.  On return rank r's payload is at mail_slot(st->mail, parity, r) (read it
// with volatile loads); returns the parity.  Each rank runs on its own GPU, so the bounded spin
// on the peers' flags is a real rendezvous; a peer that never arrives fails the launch (trap)
// after DevState::p2p_timeout_ns (default 120 s, LBFGSB200_P2P_TIMEOUT_S) instead of hanging the GPU.
__device__ int p2p_allgather(DevState *st, const double *src, int count)
{
    __shared__ unsigned long long s_seq;
    // the exchange counter lives in the mailbox (per communicator, not per solver): it stays
    // monotonic across solver handles that share a communicator and is identical on every rank
    if (threadIdx.x == 0) {
        unsigned long long *ctr = const_cast<unsigned long long *>(mail_flag(st->mail, 2, 0));
        s_seq = ++(*ctr);
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const int par = (int)(seq & 1ull), P = st->nranks, me = st->rank;
    for (int idx = threadIdx.x; idx < P * count; idx += kScalarThreads) {
        const int rk = idx / count, i = idx - rk * count;
        mail_slot(st->peers[rk], par, me)[i] = src[i]; // NVLink store into the peer's HBM
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < P) {
        *mail_flag(st->peers[threadIdx.x], par, me) = seq;
        volatile unsigned long long *f = mail_flag(st->mail, par, threadIdx.x);
        const unsigned long long t0 = global_ns();
        while (*f < seq) {
            if (global_ns() - t0 > st->p2p_timeout_ns) __trap();
        }
    }
    __threadfence_system();
    __syncthreads();
    return par;
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
        st->trial_evals += 1;
        st->vec_streams += 2.0; // trial: reads x, d
        const int cont = ls_step(st->lsp, st->ls, r[0], r[1]);
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

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_mark(DevState *st, int op, unsigned long long t_in)
{
    if (st->tl && threadIdx.x == 0 && st->tl_n < st->tl_cap) {
        unsigned long long *row = st->tl + 3 * (size_t)st->tl_n++;
        row[0] = (unsigned long long)(long long)op;
        row[1] = t_in;
        row[2] = globaltimer_ns();
    }
}

// st is the SHARED-MEMORY copy of the solver state (see k_scalar); nparts is the partial count of the pass
// that just ended; dyn is the kernel's dynamic shared memory (op == OP_COMPACT: the window Gram matrix).
__device__ void scalar_body(DevState *st, int op, int p, int from_comm, int pack_kind, int nparts,
                            unsigned long long t_in, double *dyn)
{
    __shared__ double r[kMaxQ];
    __shared__ double pk[kPacket];
    if (op == OP_COMPACT) { // CTA-wide: pass-A sums (+ exchange) -> Gram update + coefficient recursion (compact.cuh)
        if (st->ctrl.done || st->steepest || st->h == 0) return;
        __shared__ double rows[3 * kMaxCols];
        const int cnt = st->gram_count, nrows = 3 * (2 * st->h + 1);
        if (from_comm == 1) {
            // NCCL path: k_gram_finalize + all-gather ran before this kernel; rank-ordered sum
            for (int q = threadIdx.x; q < cnt; q += kScalarThreads) {
                double v = 0.0;
                for (int k = 0; k < st->nranks; ++k) v += st->gram_recv[(size_t)k * cnt + q];
                rows[q] = v;
            }
            __syncthreads();
        } else {
            gram_rows_from_partials(st->partials, nparts, nrows, cnt, rows);
            if (from_comm == 2) {
                // peer-to-peer: all-gather the pass-A rows and add them in rank order
                const int par = p2p_allgather(st, rows, cnt);
                for (int q = threadIdx.x; q < cnt; q += kScalarThreads) {
                    double v = 0.0;
                    for (int k = 0; k < st->nranks; ++k)
                        v += const_cast<const volatile double *>(mail_slot(st->mail, par, k))[q];
                    rows[q] = v;
                }
                __syncthreads();
            }
        }
        compact_recursion(st, rows, dyn);
        __syncthreads();
        tl_mark(st, op, t_in);
        return;
    }
    const int nq = nq_of(op);
    const double *rv = nullptr; // gathered packets, [rank][stride]
    int rv_stride = kPacket;
    if (!from_comm) {
        reduce_partials(st->partials, nparts, nq, r);
    } else if (from_comm == 2) {
        // peer-to-peer: pack + exchange inside this kernel
        reduce_partials(st->partials, nparts, nq, r);
        if (threadIdx.x == 0) build_packet(st, op, pack_kind, r, pk);
        __syncthreads();
        const int par = p2p_allgather(st, pk, kPacket);
        rv = mail_slot(st->mail, par, 0);
        rv_stride = kMailWidth;
    } else {
        rv = st->recv;
    }
    if (rv && threadIdx.x == 0) {
        const volatile double *v = rv;
        for (int q = 0; q < nq; ++q) {
            double t = 0.0;
            for (int k = 0; k < st->nranks; ++k) t += v[(size_t)k * rv_stride + q];
            r[q] = t;
        }
        const int left = st->rank - 1, right = st->rank + 1;
        if (pack_kind == PACK_X0 || pack_kind == PACK_ACCEPT) {
            st->xL = left >= 0 ? v[(size_t)left * rv_stride + 6] : 0.0;
            st->xR = right < st->nranks ? v[(size_t)right * rv_stride + 5] : 0.0;
            st->gL = left >= 0 ? v[(size_t)left * rv_stride + 8] : 0.0;
            st->gR = right < st->nranks ? v[(size_t)right * rv_stride + 7] : 0.0;
        } else if (pack_kind == PACK_DIR) {
            st->dL = left >= 0 ? v[(size_t)left * rv_stride + 10] : 0.0;
            st->dR = right < st->nranks ? v[(size_t)right * rv_stride + 9] : 0.0;
        }
    }
    if (threadIdx.x == 0) {
        scalar_logic(st, op, p, r);
        if (st->use_graph) {
            // device-side control flow: the trial loop and the iteration loop are graph WHILE nodes
            if (op == OP_LS_INIT || op == OP_LS_STEP)
                cudaGraphSetConditional(st->cond_inner, (st->ctrl.ls_active && !st->ctrl.done) ? 1u : 0u);
            if (op == OP_ACCEPT || op == OP_ITER_BEGIN || op == OP_LS_STEP)
                cudaGraphSetConditional(st->cond_outer, (!st->ctrl.done && st->iters_left > 0) ? 1u : 0u);
        }
        tl_mark(st, op, t_in);
    }
}

// from_comm = 0: single GPU, sums come straight from the partials of the last pass.
// from_comm = 1: sums are the rank-ordered totals of the all-gathered packets (identical
//                bits on every rank); neighbours' boundary values are picked up as halo.
// from_comm = 2: as 1, with the exchange done inside this kernel through the peer mailboxes.
// The solver state (~3 KB) is staged in shared memory for the duration of the kernel and written back at
// the end: the scalar logic is a long chain of dependent reads and writes of that state by ONE thread,
// ~0.7 us per link in HBM/L2, ~30 ns in shared memory.  No other kernel runs concurrently on the state
// (stream order), so the copy is exclusive.
__global__ void __launch_bounds__(kScalarThreads)
k_scalar(DevState *gst, int op, int p, int from_comm, int pack_kind, int nparts)
{
    extern __shared__ __align__(16) double dyn[];
    __shared__ __align__(16) unsigned long long sbuf[(sizeof(DevState) + 7) / 8];
    static_assert(sizeof(DevState) % 8 == 0, "DevState is copied in 8-byte words");
    const unsigned long long t_in = globaltimer_ns();
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
