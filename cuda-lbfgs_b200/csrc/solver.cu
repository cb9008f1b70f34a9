// solver.cu -- host side of the B200-native L-BFGS hot path and its C ABI (include/lbfgsb200.h).
//
// The host does no arithmetic on problem data.  It owns the HBM arena, launches the streaming
// kernels (kernels.cuh) and the 1-CTA scalar kernel (scalar_ops.cuh) in the fixed order of one
// L-BFGS iteration, and -- in host-stepped mode -- reads back a 16-byte control block once per
// line-search trial to learn whether the device-side state machine wants another trial.  In
// graph mode (build_graph below) even that disappears: the iteration loop and the trial loop are CUDA
// graph WHILE nodes whose conditions the scalar kernel sets on the device.
//
// Replaces: LBFGS() seq/lbfgs.cpp:17-203 ; LBFGS_CUDA() par/L-BFGS.cu:105-382 and the four
// inlined-line-search variants.  There is no CPU fallback: without a CUDA device every compute
// entry point returns LBFGSB200_ERR_CUDA.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/lbfgsb200.h"
#include "comm.h"
#include "compact.cuh"
#include "kernels.cuh"
#include "scalar_ops.cuh"
#include "state.h"

namespace lb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

#define CUDA_TRY(expr)                                                                        \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            lb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr,                  \
                          cudaGetErrorString(e_));                                            \
            return LBFGSB200_ERR_CUDA;                                                        \
        }                                                                                     \
    } while (0)

#define LB_TRY(expr)                                                                          \
    do {                                                                                      \
        int rc_ = (expr);                                                                     \
        if (rc_ < 0) return rc_;                                                              \
    } while (0)

static int sm_count(int *out)
{
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    *out = sms;
    return 0;
}

// streaming-kernel grid: a multiple of the SM count (kCtasPerSm resident CTAs per SM), shrunk
// for tiny problems so idle CTAs do not pad the partial arrays
static int pick_grid(long long n, int sms, int forced, int ctas_per_sm = kCtasPerSm)
{
    if (forced > 0) return forced;
    const long long nvec = n >> 1;
    long long tiles = (nvec + kTileVec - 1) / kTileVec;
    if (tiles < 1) tiles = 1;
    const long long full = (long long)sms * ctas_per_sm;
    return (int)(tiles < full ? tiles : full);
}

enum KClass { KC_PASS = 0, KC_TRIAL = 1, KC_ACCEPT = 2, KC_OTHER = 3, KC_GRAM = 4, KC_COMBINE = 5, KC_COUNT = 6 };

} // namespace lb

using namespace lb;

struct lbfgsb200_solver {
    int sms = 0;
    int objective = 0;
    size_t n_global = 0, n_local = 0, offset = 0, stride = 0;
    int nslots = 0;
    int grid = 1, grid_accept = 1, grid_gram = 1, grid_combine = 1;
    int gram_T = 512;           // compact form: elements per vector per shared-memory tile
    int gram_tma = 0, gram_NG = 2, gram_NS = 2, gram_G = 1; // pass A: cp.async pipeline (default) or TMA bulk copies
    size_t gram_smem = 0;
    GramMaps *gram_maps = nullptr; // device: tensor maps of the tensor-map TMA variant (gram_tma == 2)
    double *gram = nullptr;     // compact form: Gram matrix + pass-A rows + delta + all-gather buffer
    lbfgsb200_params_t params;
    lbfgsb200_comm *comm = nullptr;
    trial_kernel_t trial_kernel = nullptr;   // objective-specific instantiations
    lbfgsb200_fg_device_fn cb = nullptr;     // user device objective (lbfgsb200_create_callback)
    void *cb_user = nullptr;
    double *cb_buf = nullptr;                // g_trial [stride] + {f, g.d, g.g} + a device zero
    double *x_cur = nullptr;                 // host mirror of DevState::x (the accept step swaps x / x_alt)
    accept_kernel_t accept_kernel = nullptr;

    double *arena = nullptr;    // x, x_alt, g, w, S[nslots], Y[nslots]
    double *partials = nullptr; // [kMaxQ][grid]
    double *pkt = nullptr;      // send [kPacket] + recv [nranks][kPacket]
    double *trace = nullptr;
    unsigned long long *timeline = nullptr; // LBFGSB200_TIMELINE diagnostic (DevState::tl)
    size_t trace_rows = 0;
    DevState *d_st = nullptr;
    Ctrl *h_ctrl = nullptr;     // pinned
    DevState h_snapshot;        // last state copied back

    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    bool x0_set = false;
    int64_t k_host = 0; // accepts launched so far: an upper bound on the device's h
    int64_t launches = 0;
    double last_ms = 0.0;
    double streams_at_start = 0.0; // vec_streams when the last timed region began
    double streams_last = 0.0;

    // graph mode
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    int64_t graph_fixed_launches = 0; // kernel nodes per iteration outside the trial loop

    // per-class event pairs (iterate_profiled)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev[LBFGSB200_PROFILE_CLASSES];
};

namespace lb {

struct ClassTimer {
    lbfgsb200_solver *s;
    int cls;
    ClassTimer(lbfgsb200_solver *s_, int cls_) : s(s_), cls(cls_)
    {
        if (s->profiling) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, s->stream);
            s->prof_ev[cls].push_back(e);
        }
    }
    ~ClassTimer()
    {
        if (s->profiling) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            cudaEventRecord(e, s->stream);
            s->prof_ev[cls].push_back(e);
        }
    }
};

// scalar step: on one GPU the scalar kernel sums the partials itself; on several the local sums
// and halo values are packed, all-gathered (one small NCCL call) and summed in rank order.
static int scalar_step(lbfgsb200_solver *s, int op, int p, int pack_kind, int nparts_override = -1)
{
    const int nparts = nparts_override >= 0 ? nparts_override : (op == OP_ACCEPT || op == OP_INIT) ? s->grid_accept : (op == OP_COMPACT_DIR ? s->grid_combine : s->grid);
    const bool needs_data = (op != OP_ITER_BEGIN && op != OP_LS_INIT);
    if (s->comm && s->comm->nranks > 1 && needs_data) {
        if (s->comm->p2p) {
            // pack + NVLink mailbox exchange + scalar logic fused in ONE kernel (scalar_ops.cuh)
            k_scalar<<<1, kScalarThreads, 0, s->stream>>>(s->d_st, op, p, 2, pack_kind, nparts);
            s->launches += 1;
        } else {
            k_pack<<<1, kScalarThreads, 0, s->stream>>>(s->d_st, op, pack_kind, nparts);
            LB_TRY(comm_allgather(s->comm, s->pkt, s->pkt + kPacket, kPacket, s->stream));
            k_scalar<<<1, kScalarThreads, 0, s->stream>>>(s->d_st, op, p, 1, pack_kind, nparts);
            s->launches += 2;
        }
    } else {
        k_scalar<<<1, kScalarThreads, 0, s->stream>>>(s->d_st, op, p, 0, PACK_NONE, nparts);
        s->launches += 1;
    }
    return 0;
}

// search direction: seq/lbfgs.cpp:86-153
static int launch_direction(lbfgsb200_solver *s)
{
    const int m = s->params.m;
    const int h_upper = (int)(s->k_host < m ? s->k_host : m);
    LB_TRY(scalar_step(s, OP_ITER_BEGIN, 0, PACK_NONE));
    if (h_upper > 0 && s->params.direction == LBFGSB200_DIR_COMPACT) {
        // compact form: pass A (Gram rows) -> coefficient recursion -> pass B (combine)
        const int J = 2 * h_upper + 1;
        {
            ClassTimer t(s, KC_GRAM);
            const size_t smem = s->gram_smem;
            if (s->gram_tma == 2) {
                if ((2 * m + 1 + s->gram_NG - 1) / s->gram_NG <= 7)
                    k_gram_tma2d<7><<<s->grid_gram, kWsThreads, smem, s->stream>>>(s->d_st, s->gram_maps, s->gram_T, s->gram_NG);
                else
                    k_gram_tma2d<kMaxCW><<<s->grid_gram, kWsThreads, smem, s->stream>>>(s->d_st, s->gram_maps, s->gram_T, s->gram_NG);
            } else if (s->gram_tma) {
                if ((2 * m + 1 + s->gram_NG - 1) / s->gram_NG <= 7)
                    k_gram_tma<7><<<s->grid_gram, kWsThreads, smem, s->stream>>>(s->d_st, s->gram_T, s->gram_NG);
                else
                    k_gram_tma<kMaxCW><<<s->grid_gram, kWsThreads, smem, s->stream>>>(s->d_st, s->gram_T, s->gram_NG);
            } else {
                const int G = s->gram_G, per = (2 * m + 1 + G - 1) / G, cwg = (per + kGramWarps - 1) / kGramWarps;
                const dim3 grid(s->grid_gram, G);
                if (cwg <= 3) k_gram<3><<<grid, kThreads, smem, s->stream>>>(s->d_st, s->gram_T, s->gram_NS, G);
                else if (cwg <= 6) k_gram<6><<<grid, kThreads, smem, s->stream>>>(s->d_st, s->gram_T, s->gram_NS, G);
                else k_gram<kMaxCW><<<grid, kThreads, smem, s->stream>>>(s->d_st, s->gram_T, s->gram_NS, G);
            }
            s->launches += 1;
        }
        const bool multi = s->comm && s->comm->nranks > 1;
        const bool p2p = multi && s->comm->p2p;
        if (multi && !p2p) {
            // NCCL path: the rows must be in HBM for the all-gather (otherwise the scalar kernel sums them itself)
            k_gram_finalize<<<3 * J, kScalarThreads, 0, s->stream>>>(s->d_st, s->grid_gram);
            s->launches += 1;
            const int cnt = 3 * (2 * m + 1);
            double *rows = s->h_snapshot.gram_rows;
            LB_TRY(comm_allgather(s->comm, rows, s->h_snapshot.gram_recv, cnt, s->stream));
        }
        // dynamic shared memory: the (2h+1)^2 window Gram matrix of the coefficient recursion
        k_scalar<<<1, kScalarThreads, sizeof(double) * (size_t)J * J, s->stream>>>(s->d_st, OP_COMPACT, 0, p2p ? 2 : (multi ? 1 : 0),
                                                                                   PACK_NONE, s->grid_gram);
        s->launches += 1;
        {
            ClassTimer t(s, KC_COMBINE);
            k_combine<<<s->grid_combine, kThreads, 0, s->stream>>>(s->d_st);
            s->launches += 1;
        }
        LB_TRY(scalar_step(s, OP_COMPACT_DIR, 0, PACK_DIR));
    } else if (h_upper > 0) {
        {
            ClassTimer t(s, KC_OTHER);
            k_dot_sg<<<s->grid, kThreads, 0, s->stream>>>(s->d_st);
            s->launches += 1;
        }
        LB_TRY(scalar_step(s, OP_SG, 0, PACK_NONE));
        for (int p = h_upper - 1; p >= 0; --p) {
            {
                ClassTimer t(s, KC_PASS);
                k_two_loop_pass<<<s->grid, kThreads, 0, s->stream>>>(s->d_st, 1, p);
                s->launches += 1;
            }
            LB_TRY(scalar_step(s, OP_L1, p, PACK_NONE));
        }
        for (int p = 0; p < h_upper; ++p) {
            {
                ClassTimer t(s, KC_PASS);
                k_two_loop_pass<<<s->grid, kThreads, 0, s->stream>>>(s->d_st, 2, p);
                s->launches += 1;
            }
            LB_TRY(scalar_step(s, OP_L2, p, PACK_DIR));
        }
    }
    {
        ClassTimer t(s, KC_OTHER);
        k_steepest<<<s->grid, kThreads, 0, s->stream>>>(s->d_st);
        s->launches += 1;
    }
    LB_TRY(scalar_step(s, OP_LS_INIT, 0, PACK_NONE));
    return 0;
}

static int read_ctrl(lbfgsb200_solver *s)
{
    CUDA_TRY(cudaMemcpyAsync(s->h_ctrl, &s->d_st->ctrl, sizeof(Ctrl), cudaMemcpyDeviceToHost,
                             s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

static int snapshot(lbfgsb200_solver *s)
{
    CUDA_TRY(cudaMemcpyAsync(&s->h_snapshot, s->d_st, sizeof(DevState), cudaMemcpyDeviceToHost,
                             s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

// host-stepped iteration loop
static int run_stepped(lbfgsb200_solver *s, int64_t iterations)
{
    for (int64_t it = 0; it < iterations; ++it) {
        LB_TRY(launch_direction(s));
        // line search: one fused evaluation + one device-side decision per trial
        do {
            {
                ClassTimer t(s, KC_TRIAL);
                s->trial_kernel<<<s->grid, kThreads, 0, s->stream>>>(s->d_st);
                s->launches += 1;
            }
            LB_TRY(scalar_step(s, OP_LS_STEP, 0, PACK_NONE));
            LB_TRY(read_ctrl(s));
        } while (s->h_ctrl->ls_active && !s->h_ctrl->done);
        if (s->h_ctrl->done) break;
        {
            ClassTimer t(s, KC_ACCEPT);
            s->accept_kernel<<<s->grid_accept, kThreads, 0, s->stream>>>(s->d_st, 0);
            s->launches += 1;
        }
        LB_TRY(scalar_step(s, OP_ACCEPT, 0, PACK_ACCEPT));
        s->k_host += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- graph mode -------------------------------------------------------------------------------
// One CUDA graph = the whole solve.  Its single top-level node is a WHILE node (condition: not
// done and iteration budget left) whose body is one L-BFGS iteration: the direction phase, a
// nested WHILE node around {fused trial evaluation, line-search decision}, and the accept step.
// Both conditions are set on the device by k_scalar (cudaGraphSetConditional), so the host
// launches ONE graph per iterate() call and reads nothing back until it ends.
#define GRAPH_TRY(expr)                                                                       \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            lb::set_error("graph build: %s failed: %s", #expr, cudaGetErrorString(e_));       \
            return LBFGSB200_ERR_CUDA;                                                        \
        }                                                                                     \
    } while (0)

template <class F>
static int capture_segment(lbfgsb200_solver *s, cudaGraph_t g, std::vector<cudaGraphNode_t> &tail, F &&body)
{
    GRAPH_TRY(cudaStreamBeginCaptureToGraph(s->stream, g, tail.empty() ? nullptr : tail.data(), nullptr, tail.size(),
                                            cudaStreamCaptureModeThreadLocal));
    int rc = body();
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    cudaError_t e = cudaStreamGetCaptureInfo_v2(s->stream, &status, nullptr, nullptr, &deps, &ndeps);
    if (e == cudaSuccess) tail.assign(deps, deps + ndeps);
    cudaGraph_t out = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(s->stream, &out);
    if (rc < 0) return rc;
    GRAPH_TRY(e);
    GRAPH_TRY(e2);
    return 0;
}

static int add_while(cudaGraph_t parent, std::vector<cudaGraphNode_t> &tail, cudaGraphConditionalHandle h,
                     cudaGraph_t *body)
{
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = cudaGraphCondTypeWhile;
    p.conditional.size = 1;
    cudaGraphNode_t node;
    GRAPH_TRY(cudaGraphAddNode(&node, parent, tail.empty() ? nullptr : tail.data(), tail.size(), &p));
    *body = p.conditional.phGraph_out[0];
    tail.assign(1, node);
    return 0;
}

static int build_graph(lbfgsb200_solver *s)
{
    GRAPH_TRY(cudaGraphCreate(&s->graph, 0));
    cudaGraphConditionalHandle h_outer, h_inner;
    GRAPH_TRY(cudaGraphConditionalHandleCreate(&h_outer, s->graph, 1, cudaGraphCondAssignDefault));
    GRAPH_TRY(cudaGraphConditionalHandleCreate(&h_inner, s->graph, 0, 0));
    s->h_snapshot.cond_outer = h_outer;
    s->h_snapshot.cond_inner = h_inner;
    s->h_snapshot.use_graph = 1;
    CUDA_TRY(cudaMemcpyAsync(&s->d_st->cond_outer, &s->h_snapshot.cond_outer, sizeof h_outer, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(&s->d_st->cond_inner, &s->h_snapshot.cond_inner, sizeof h_inner, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaMemcpyAsync(&s->d_st->use_graph, &s->h_snapshot.use_graph, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));

    std::vector<cudaGraphNode_t> top_tail, tail;
    cudaGraph_t iter_body = nullptr, trial_body = nullptr;
    LB_TRY(add_while(s->graph, top_tail, h_outer, &iter_body));
    const int64_t k_saved = s->k_host, l_saved = s->launches;
    s->k_host = s->params.m; // capture the passes of all m window positions; unused ones exit at once
    LB_TRY(capture_segment(s, iter_body, tail, [&]() { return launch_direction(s); }));
    LB_TRY(add_while(iter_body, tail, h_inner, &trial_body));
    std::vector<cudaGraphNode_t> inner_tail;
    LB_TRY(capture_segment(s, trial_body, inner_tail, [&]() {
        s->trial_kernel<<<s->grid, kThreads, 0, s->stream>>>(s->d_st);
        return scalar_step(s, OP_LS_STEP, 0, PACK_NONE);
    }));
    const int64_t after_inner = s->launches;
    LB_TRY(capture_segment(s, iter_body, tail, [&]() {
        s->accept_kernel<<<s->grid_accept, kThreads, 0, s->stream>>>(s->d_st, 0);
        s->launches += 1;
        return scalar_step(s, OP_ACCEPT, 0, PACK_ACCEPT);
    }));
    (void)after_inner;
    // fixed part = everything captured except the two nodes of the trial loop body
    s->graph_fixed_launches = (s->launches - l_saved) - 1; // scalar_step of the trial body counted once; k_trial not counted
    s->k_host = k_saved;
    s->launches = l_saved;
    GRAPH_TRY(cudaGraphInstantiate(&s->graph_exec, s->graph, 0));
    return 0;
}

static int run_graph(lbfgsb200_solver *s, int64_t iterations)
{
    if (iterations <= 0) return 0;
    if (!s->graph_exec) LB_TRY(build_graph(s));
    long long budget = iterations;
    CUDA_TRY(cudaMemcpyAsync(&s->d_st->iters_left, &budget, sizeof budget, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaGraphLaunch(s->graph_exec, s->stream));
    return 0;
}

// host-stepped loop for user (callback) objectives: the callback is launched by the host, so the
// host has to know after every decision whether another evaluation is wanted
static int run_stepped_callback(lbfgsb200_solver *s, int64_t iterations)
{
    double *g_trial = s->cb_buf, *scal = s->cb_buf + s->stride;
    const double *d_alpha = &s->d_st->ls.alpha;
    for (int64_t it = 0; it < iterations; ++it) {
        LB_TRY(launch_direction(s));
        LB_TRY(read_ctrl(s));
        while (s->h_ctrl->ls_active && !s->h_ctrl->done) {
            if (s->cb(s->x_cur, s->h_snapshot.w, d_alpha, g_trial, s->partials, s->n_local, 0, s->cb_user, s->stream)) {
                set_error("the objective callback failed");
                return LBFGSB200_ERR_INVALID;
            }
            LB_TRY(scalar_step(s, OP_LS_STEP, 0, PACK_NONE, 1)); // the 3 sums are already final: one "partial" each
            LB_TRY(read_ctrl(s));
        }
        if (s->h_ctrl->done) break;
        // gradient and f at the accepted step (the last trial may have been at another alpha)
        if (s->cb(s->x_cur, s->h_snapshot.w, d_alpha, g_trial, scal, s->n_local, 0, s->cb_user, s->stream)) {
            set_error("the objective callback failed");
            return LBFGSB200_ERR_INVALID;
        }
        k_accept_generic<<<s->grid, kThreads, 0, s->stream>>>(s->d_st, g_trial, scal, 0);
        s->launches += 1;
        LB_TRY(scalar_step(s, OP_ACCEPT, 0, PACK_ACCEPT, s->grid));
        s->x_cur = (s->x_cur == s->arena) ? s->arena + s->stride : s->arena;
        s->k_host += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

static int do_iterate(lbfgsb200_solver *s, int64_t iterations)
{
    if (!s->x0_set) {
        set_error("iterate: set_x0 has not been called");
        return LBFGSB200_ERR_INVALID;
    }
    s->streams_at_start = s->h_snapshot.vec_streams;
    // graph mode: single GPU, not instrumented (NCCL calls and event pairs stay on the stepped path)
    // (multi-GPU: only with the peer-to-peer exchange, whose kernels are ordinary graph nodes)
    const bool graph = s->params.use_graph && !s->profiling && !s->cb && !(s->comm && s->comm->nranks > 1 && !s->comm->p2p);
    if (graph && !s->graph_exec) LB_TRY(build_graph(s));
    if (s->graph_exec) { // cudaGraphSetConditional is only legal inside the graph: gate it per run
        const int flag = graph ? 1 : 0;
        CUDA_TRY(cudaMemcpyAsync(&s->d_st->use_graph, &flag, sizeof flag, cudaMemcpyHostToDevice, s->stream));
    }
    const long long k0 = s->h_snapshot.k, t0 = s->h_snapshot.trial_evals;
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    int rc = graph ? run_graph(s, iterations) : (s->cb ? run_stepped_callback(s, iterations) : run_stepped(s, iterations));
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    if (rc < 0) return rc;
    LB_TRY(snapshot(s));
    if (graph) { // kernel nodes executed: fixed part per iteration + 2 per trial
        const long long its = s->h_snapshot.k - k0, trials = s->h_snapshot.trial_evals - t0;
        // an exit at the top of / inside an iteration (converged, line search failed) still ran that
        // iteration's nodes as no-ops; "maximum iterations" is raised by the last accept itself
        const bool extra = s->h_snapshot.ctrl.done && s->h_snapshot.status != LBFGSB200_MAX_ITER;
        s->launches += (its + (extra ? 1 : 0)) * s->graph_fixed_launches + 2 * trials;
        s->k_host += its;
    }
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
    s->last_ms = ms;
    s->streams_last = s->h_snapshot.vec_streams - s->streams_at_start;
    return s->h_snapshot.status;
}

} // namespace lb

// ====================================================================
// C ABI
// ====================================================================
extern "C" {

int lbfgsb200_version(void) { return LBFGSB200_VERSION; }

const char *lbfgsb200_last_error(void) { return lb::g_err; }

const char *lbfgsb200_strerror(int status)
{
    switch (status) {
    case LBFGSB200_CONVERGED: return "converged";
    case LBFGSB200_MAX_ITER: return "maximum iterations reached";
    case LBFGSB200_LS_FAILED: return "line search failed";
    case LBFGSB200_RUNNING: return "running";
    case LBFGSB200_ERR_INVALID: return "invalid argument";
    case LBFGSB200_ERR_CUDA: return "CUDA error";
    case LBFGSB200_ERR_NCCL: return "NCCL error";
    case LBFGSB200_ERR_NOMEM: return "out of device memory";
    default: return "unknown status";
    }
}

int lbfgsb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int lbfgsb200_params_default(lbfgsb200_params_t *p, int flavor)
{
    if (!p) return LBFGSB200_ERR_INVALID;
    memset(p, 0, sizeof *p);
    p->m = 10;                 // seq/lbfgs.h:23
    p->max_iterations = 1000;  // seq/lbfgs.h:22
    p->tolerance = 1e-5;       // seq/lbfgs.h:24
    p->line_search = LBFGSB200_LS_BACKTRACKING;
    p->flavor = flavor;
    p->profile = LBFGSB200_PROFILE_SEQ;
    p->direction = LBFGSB200_DIR_TWO_LOOP;
    p->c1 = 1e-4;                                         // seq/config.h:5, par/constants.h:5
    p->c2 = (flavor != LBFGSB200_FLAVOR_SEQ) ? 0.7 : 0.9; // par/constants.h:6 / seq/config.h:6
    p->step0 = 1.0;                                       // INITIAL_STEP_SIZE
    p->shrink = 0.5;                                      // BACKTRACKING_ALPHA
    p->backtracking_tol = (flavor == LBFGSB200_FLAVOR_PAR_INLINED) ? 1e-10 : 1e-8; // BACKTRACKING_TOL (par/L-BFGS-Backtracking.cu:155)
    p->wolfe_min = 1e-10;                                 // WOLFE_INTERP_MIN
    p->ls_max_trials = 20;
    p->use_graph = 0;
    p->verbose = 0;
    p->grid_ctas = 0;
    return 0;
}

void lbfgsb200_shard_range(size_t n_global, int rank, int nranks, size_t *offset, size_t *n_local)
{
    // even-sized shards so every shard starts on a 16-byte boundary of the global vector and
    // the double2 path never straddles ranks; the last rank takes the remainder
    if (nranks < 1) nranks = 1;
    size_t chunk = (n_global / (size_t)nranks) & ~(size_t)1;
    size_t off = chunk * (size_t)rank;
    size_t len = (rank == nranks - 1) ? n_global - off : chunk;
    if (offset) *offset = off;
    if (n_local) *n_local = len;
}

// Tensor maps for the tensor-map TMA variant of pass A: S and Y are [nslots][stride] row-major FP64
// tensors; one map per run length r (box = T columns x r rows), plus a 1-row map over g.
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int build_gram_maps(lbfgsb200_solver *s, const DevState &st)
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess) {
        set_error("cuTensorMapEncodeTiled is not available");
        cudaGetLastError();
        return LBFGSB200_ERR_CUDA;
    }
    encode_tiled_fn encode = (encode_tiled_fn)fn;
    GramMaps *host = new GramMaps;
    memset(host, 0, sizeof *host);
    auto make = [&](CUtensorMap *out, double *base, int rows_total, int box_rows) -> bool {
        const cuuint64_t gdim[2] = {(cuuint64_t)s->stride, (cuuint64_t)rows_total};
        const cuuint64_t gstride[1] = {(cuuint64_t)s->stride * sizeof(double)};
        const cuuint32_t box[2] = {(cuuint32_t)s->gram_T, (cuuint32_t)box_rows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) set_error("cuTensorMapEncodeTiled failed with %d (rows %d, box %d x %d)", (int)r, rows_total, s->gram_T, box_rows);
        return r == CUDA_SUCCESS;
    };
    bool ok = true;
    for (int r = 1; r <= s->nslots && ok; ++r) ok = make(&host->s[r], st.S, s->nslots, r) && make(&host->y[r], st.Y, s->nslots, r);
    ok = ok && make(&host->g, st.g, 1, 1);
    int rc = 0;
    if (!ok) rc = LBFGSB200_ERR_CUDA;
    if (!rc && cudaMalloc(&s->gram_maps, sizeof(GramMaps)) != cudaSuccess) rc = LBFGSB200_ERR_NOMEM;
    if (!rc && cudaMemcpy(s->gram_maps, host, sizeof(GramMaps), cudaMemcpyHostToDevice) != cudaSuccess) rc = LBFGSB200_ERR_CUDA;
    delete host;
    return rc;
}

static int check_params(const lbfgsb200_params_t *p)
{
    if (!p) { set_error("params is NULL"); return LBFGSB200_ERR_INVALID; }
    if (p->m < 1 || p->m > LBFGSB200_MAX_M) { set_error("m=%d out of range 1..%d", p->m, LBFGSB200_MAX_M); return LBFGSB200_ERR_INVALID; }
    if (p->line_search < 0 || p->line_search > 3) {
        // the reference throws std::invalid_argument("Unknown line search method") seq/lbfgs.cpp:69
        set_error("Unknown line search method: %d", p->line_search);
        return LBFGSB200_ERR_INVALID;
    }
    if (p->flavor < 0 || p->flavor > 2 || p->profile < 0 || p->profile > 1 || p->direction < 0 ||
        p->direction > 1) {
        set_error("bad flavor/profile/direction");
        return LBFGSB200_ERR_INVALID;
    }
    if (p->direction == LBFGSB200_DIR_COMPACT && p->m > kMaxCompactM) { set_error("compact direction supports m <= %d", kMaxCompactM); return LBFGSB200_ERR_INVALID; }
    if (p->max_iterations < 0 || p->ls_max_trials < 1) { set_error("bad iteration limits"); return LBFGSB200_ERR_INVALID; }
    return 0;
}

int lbfgsb200_create(lbfgsb200_solver_t **out, int objective, size_t n_global,
                     const lbfgsb200_params_t *params, lbfgsb200_comm_t *comm, size_t trace_rows)
{
    if (!out) return LBFGSB200_ERR_INVALID;
    *out = nullptr;
    LB_TRY(check_params(params));
    if ((objective < 0 || objective > LBFGSB200_OBJ_TRIDIAG) && objective != LBFGSB200_OBJ_DEVICE_CALLBACK) { set_error("unknown objective %d", objective); return LBFGSB200_ERR_INVALID; }
    if (n_global == 0) { set_error("n must be > 0"); return LBFGSB200_ERR_INVALID; }
    if (lbfgsb200_device_count() < 1) {
        set_error("no usable CUDA device: this library has no CPU fallback");
        return LBFGSB200_ERR_CUDA;
    }
    lbfgsb200_solver *s = new (std::nothrow) lbfgsb200_solver;
    if (!s) return LBFGSB200_ERR_NOMEM;
    s->params = *params;
    s->objective = objective;
    s->trial_kernel = trial_kernel_for(objective);
    s->accept_kernel = accept_kernel_for(objective);
    s->comm = comm;
    s->n_global = n_global;
    const int rank = comm ? comm->rank : 0, nranks = comm ? comm->nranks : 1;
    if (nranks > kMaxRanks) { set_error("at most %d ranks", kMaxRanks); delete s; return LBFGSB200_ERR_INVALID; }
    lbfgsb200_shard_range(n_global, rank, nranks, &s->offset, &s->n_local);
    if (s->n_local == 0) { set_error("rank %d owns no elements (n=%zu, ranks=%d)", rank, n_global, nranks); delete s; return LBFGSB200_ERR_INVALID; }
    s->nslots = params->m + 1;
    s->stride = (s->n_local + 31) / 32 * 32; // 256-byte rows
    int rc = sm_count(&s->sms);
    if (rc < 0) { delete s; return rc; }
    s->grid = pick_grid((long long)s->n_local, s->sms, params->grid_ctas);
    s->grid_accept = pick_grid((long long)s->n_local, s->sms, params->grid_ctas, kCtasPerSmAccept);
    if (params->direction == LBFGSB200_DIR_COMPACT) {
        // shared-memory tile of all 2m+1 basis vectors; keep two CTAs per SM resident
        const int J = 2 * params->m + 1;
        // Three implementations of pass A, picked by history size, LBFGSB200_GRAM_TMA forces one:
        //   2  tensor-map TMA (UTMALDG.2D): <= 5 tiled loads per tile of the whole history.  Default for
        //      m > 6: 7.0-7.3 TB/s at m = 10..50 (1.07-1.11x the measured copy peak)
        //   1  one-row bulk copies (UBLKCP), tiles of 512 elements.  Default for m <= 6: 6.8-7.2 TB/s
        //   0  cp.async (LDGSTS) pipeline with column groups: 6.1-6.9 TB/s; fallback when the tensor maps
        //      cannot be built (shards of 2^31 elements or more: TMA coordinates are 32-bit)
        const char *env = getenv("LBFGSB200_GRAM_TMA");
        s->gram_tma = env ? atoi(env) : (params->m <= 6 ? 1 : 2);
        if (s->gram_tma == 2 && s->stride >= ((size_t)1 << 31)) s->gram_tma = 0;
        int Jt = J;
        if (s->gram_tma) {
            // TMA variant: ONE CTA per SM owning (almost) all of shared memory: kGramStages stages of the
            // largest tile that fits ~200 KB, so kGramStages-1 whole tiles per SM are in flight
            const char *eb = getenv("LBFGSB200_GRAM_TMA_KB");
            const size_t budget = (size_t)(eb ? atoi(eb) : 216) * 1024;
            s->gram_NS = kGramStages;
            s->gram_T = s->gram_tma == 2 ? 256 : 512; // a TMA box dimension is at most 256 elements
            while (s->gram_T > 32 && (size_t)s->gram_NS * J * s->gram_T * sizeof(double) > budget) s->gram_T >>= 1;
        } else {
            // cp.async pipeline.  Large tiles matter (per-tile barrier/issue overhead): split the basis
            // into G column groups of <= ~40 columns (+3 row vectors when G > 1), take the largest T
            // whose double-buffered tile fits ~100 KB (two CTAs per SM), then add stages if room is left.
            const char *et = getenv("LBFGSB200_GRAM_T"), *es = getenv("LBFGSB200_GRAM_NS"), *eg = getenv("LBFGSB200_GRAM_G");
            s->gram_G = eg ? atoi(eg) : (J + 40) / 41;
            const int per = (J + s->gram_G - 1) / s->gram_G;
            Jt = s->gram_G == 1 ? J : per + 3;
            s->gram_T = 512;
            while (s->gram_T > 32 && 2 * (size_t)Jt * s->gram_T * sizeof(double) > 100 * 1024) s->gram_T >>= 1;
            if (et) s->gram_T = atoi(et);
            int ns = (int)((100 * 1024) / ((size_t)Jt * s->gram_T * sizeof(double)));
            s->gram_NS = ns < 2 ? 2 : (ns > 4 ? 4 : ns);
            if (es) s->gram_NS = atoi(es);
        }
        s->gram_smem = (size_t)s->gram_NS * Jt * s->gram_T * sizeof(double);
        // TMA variant: 16 consumer warps = NG column groups x NE element groups.  Each element group
        // should span >= 32 double2 items (all lanes busy) and no warp may own more than kMaxCW columns.
        int NE = s->gram_T / 2 / 32;
        if (NE < 1) NE = 1;
        if (NE > kWsConsumerWarps / 2) NE = kWsConsumerWarps / 2;
        s->gram_NG = kWsConsumerWarps / NE;
        while (s->gram_NG < kWsConsumerWarps && (J + s->gram_NG - 1) / s->gram_NG > kMaxCW) s->gram_NG <<= 1;
        long long tiles = ((long long)s->n_local + s->gram_T - 1) / s->gram_T;
        const long long full = (long long)s->sms * kGramCtasPerSm;
        s->grid_combine = pick_grid((long long)s->n_local, s->sms, params->grid_ctas, kCombineCtasPerSm);
        long long gx = full / s->gram_G; // the G column groups share the resident-CTA slots
        if (s->gram_tma) gx = s->sms; // warp-specialised: one CTA per SM
        if (gx < 1) gx = 1;
        s->grid_gram = params->grid_ctas > 0 ? params->grid_ctas : (int)(tiles < gx ? (tiles < 1 ? 1 : tiles) : gx);
    }

    const size_t nvecs = 4 + 2 * (size_t)s->nslots;
    const size_t arena_bytes = nvecs * s->stride * sizeof(double);
#define CREATE_TRY(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e2_ = (expr);                                                             \
        if (e2_ != cudaSuccess) {                                                             \
            set_error("%s failed: %s", #expr, cudaGetErrorString(e2_));                       \
            lbfgsb200_destroy(s);                                                             \
            return LBFGSB200_ERR_CUDA;                                                        \
        }                                                                                     \
    } while (0)
    CREATE_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreate(&s->ev0));
    CREATE_TRY(cudaEventCreate(&s->ev1));
    // The arena ((2m+6) vectors, tens of GB) comes from the device's stream-ordered memory pool with the
    // release threshold raised, so that destroying a solver and creating the next one of similar size
    // re-uses the mapping instead of paying cudaMalloc/cudaFree of tens of GB (5-130 ms each way) again.
    // lbfgsb200_trim_memory() hands the cached memory back to the driver.
    cudaMemPool_t pool;
    {
        int dev = 0;
        CREATE_TRY(cudaGetDevice(&dev));
        CREATE_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
        uint64_t keep = UINT64_MAX;
        CREATE_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    cudaError_t e = cudaMallocAsync(&s->arena, arena_bytes, s->stream);
    if (e == cudaErrorMemoryAllocation) {
        // an arena cached from a destroyed solver of another size may be what is in the way: give the
        // cache back to the driver and try once more
        cudaGetLastError();
        s->arena = nullptr;
        cudaStreamSynchronize(s->stream);
        cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocAsync(&s->arena, arena_bytes, s->stream);
    }
    if (e != cudaSuccess) {
        set_error("allocation of %.2f GB for %zu vectors failed: %s", arena_bytes / 1e9, nvecs, cudaGetErrorString(e));
        cudaGetLastError();
        s->arena = nullptr;
        lbfgsb200_destroy(s);
        return LBFGSB200_ERR_NOMEM;
    }
    // Zero only what must be zero: the work vector w (d = 0 for the x0 evaluation) and the <= 31
    // pad doubles at the end of every row (the compact kernels read rows up to the padded length).
    // Everything else is written before it is read; clearing the whole arena would cost a full
    // HBM pass over (2m+6) vectors on every create().
    {
        const size_t pad = s->stride - s->n_local;
        for (size_t v = 0; v < nvecs; ++v) {
            if (v == 3) CREATE_TRY(cudaMemsetAsync(s->arena + v * s->stride, 0, s->stride * sizeof(double), s->stream));
            else if (pad) CREATE_TRY(cudaMemsetAsync(s->arena + v * s->stride + s->n_local, 0, pad * sizeof(double), s->stream));
        }
    }
    size_t npart = (size_t)kMaxQ * (size_t)s->grid;
    if (params->direction == LBFGSB200_DIR_COMPACT) {
        const size_t need = (size_t)3 * (2 * params->m + 1) * (size_t)s->grid_gram;
        if (need > npart) npart = need;
    }
    CREATE_TRY(cudaMalloc(&s->partials, sizeof(double) * npart));
    CREATE_TRY(cudaMemsetAsync(s->partials, 0, sizeof(double) * npart, s->stream));
    size_t gram_nb = 0, gram_cnt = 0;
    if (params->direction == LBFGSB200_DIR_COMPACT) {
        gram_nb = (size_t)(2 * s->nslots + 1);
        gram_cnt = (size_t)3 * (2 * params->m + 1);
        const size_t total = gram_nb * gram_nb + gram_cnt * (size_t)(nranks + 1) + (size_t)(2 * params->m + 1);
        CREATE_TRY(cudaMalloc(&s->gram, sizeof(double) * total));
        CREATE_TRY(cudaMemsetAsync(s->gram, 0, sizeof(double) * total, s->stream));
        // opt every pass-A variant in to the device's full dynamic shared memory once (the limit is per
        // function and process-wide; occupancy still follows the size actually passed at launch)
        int dev = 0, optin = 0;
        CREATE_TRY(cudaGetDevice(&dev));
        CREATE_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        if ((size_t)optin < s->gram_smem + 4096) {
            set_error("pass A needs %zu bytes of shared memory, the device offers %d", s->gram_smem, optin);
            lbfgsb200_destroy(s);
            return LBFGSB200_ERR_INVALID;
        }
        CREATE_TRY(cudaFuncSetAttribute((const void *)k_scalar, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)(sizeof(double) * kMaxCols * kMaxCols)));
        const void *variants[] = {(const void *)k_gram<3>, (const void *)k_gram<6>, (const void *)k_gram<kMaxCW>,
                                  (const void *)k_gram_tma<7>, (const void *)k_gram_tma<kMaxCW>,
                                  (const void *)k_gram_tma2d<7>, (const void *)k_gram_tma2d<kMaxCW>};
        for (const void *fn : variants) {
            cudaFuncAttributes fa;
            CREATE_TRY(cudaFuncGetAttributes(&fa, fn));
            CREATE_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
        }
    }
    CREATE_TRY(cudaMalloc(&s->pkt, sizeof(double) * kPacket * (size_t)(nranks + 1)));
    CREATE_TRY(cudaMemsetAsync(s->pkt, 0, sizeof(double) * kPacket * (size_t)(nranks + 1), s->stream));
    s->trace_rows = trace_rows;
    if (trace_rows) {
        CREATE_TRY(cudaMalloc(&s->trace, sizeof(double) * LBFGSB200_TRACE_COLS * trace_rows));
        CREATE_TRY(cudaMemsetAsync(s->trace, 0, sizeof(double) * LBFGSB200_TRACE_COLS * trace_rows, s->stream));
    }
    CREATE_TRY(cudaMalloc(&s->d_st, sizeof(DevState)));
    CREATE_TRY(cudaHostAlloc(&s->h_ctrl, sizeof(Ctrl), cudaHostAllocDefault));
    memset(s->h_ctrl, 0, sizeof(Ctrl));

    DevState &st = s->h_snapshot;
    memset(&st, 0, sizeof st);
    st.n = (long long)s->n_local;
    st.goff = (long long)s->offset;
    st.nglob = (long long)n_global;
    st.m = params->m;
    st.nslots = s->nslots;
    st.objective = objective;
    st.profile = params->profile;
    st.direction = params->direction;
    st.max_iterations = params->max_iterations;
    st.tolerance = params->tolerance;
    st.rank = rank;
    st.nranks = nranks;
    st.grid = s->grid;
    st.grid_accept = s->grid_accept;
    st.x = s->arena;
    st.x_alt = s->arena + s->stride;
    st.g = s->arena + 2 * s->stride;
    st.w = s->arena + 3 * s->stride;
    st.S = s->arena + 4 * s->stride;
    st.Y = st.S + (size_t)s->nslots * s->stride;
    st.stride = (long long)s->stride;
    if (comm && comm->p2p) {
        st.p2p = 1;
        const char *et = getenv("LBFGSB200_P2P_TIMEOUT_S");
        st.p2p_timeout_ns = (unsigned long long)(et ? atoi(et) : 120) * 1000000000ull;
        st.mail = comm->mail;
        st.peers = comm->peers_dev;
    }
    st.partials = s->partials;
    st.send = s->pkt;
    st.recv = s->pkt + kPacket;
    st.trace = s->trace;
    st.trace_rows = (long long)trace_rows;
    if (s->gram) {
        st.gram = s->gram;
        st.gram_rows = s->gram + gram_nb * gram_nb;
        st.gram_recv = st.gram_rows + gram_cnt;
        st.delta = st.gram_recv + gram_cnt * (size_t)nranks;
        st.gram_count = (int)gram_cnt;
    }
    st.lsp.kind = params->line_search;
    st.lsp.flavor = params->flavor;
    st.lsp.max_trials = params->ls_max_trials;
    st.lsp.c1 = params->c1;
    st.lsp.c2 = params->c2;
    st.lsp.step0 = params->step0;
    st.lsp.shrink = params->shrink;
    st.lsp.bt_tol = params->backtracking_tol;
    st.lsp.wolfe_min = params->wolfe_min;
    st.status = LBFGSB200_RUNNING;
    if (const char *et = getenv("LBFGSB200_TIMELINE")) {
        st.tl_cap = atoi(et);
        if (st.tl_cap > 0) {
            CREATE_TRY(cudaMalloc(&s->timeline, sizeof(unsigned long long) * 3 * (size_t)st.tl_cap));
            st.tl = s->timeline;
        }
    }
    if (s->gram && s->gram_tma == 2) {
        int rc_maps = build_gram_maps(s, st);
        if (rc_maps < 0) { lbfgsb200_destroy(s); return rc_maps; }
    }
    CREATE_TRY(cudaMemcpyAsync(s->d_st, &st, sizeof st, cudaMemcpyHostToDevice, s->stream));
    CREATE_TRY(cudaStreamSynchronize(s->stream));
#undef CREATE_TRY
    *out = s;
    return 0;
}

void lbfgsb200_destroy(lbfgsb200_solver_t *s)
{
    if (!s) return;
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int c = 0; c < LBFGSB200_PROFILE_CLASSES; ++c)
        for (cudaEvent_t e : s->prof_ev[c]) cudaEventDestroy(e);
    if (s->arena && s->stream) {
        cudaFreeAsync(s->arena, s->stream); // back to the pool (see lbfgsb200_create)
        cudaStreamSynchronize(s->stream);
    }
    if (s->partials) cudaFree(s->partials);
    if (s->gram) cudaFree(s->gram);
    if (s->gram_maps) cudaFree(s->gram_maps);
    if (s->cb_buf) cudaFree(s->cb_buf);
    if (s->pkt) cudaFree(s->pkt);
    if (s->trace) cudaFree(s->trace);
    if (s->timeline) cudaFree(s->timeline);
    if (s->d_st) cudaFree(s->d_st);
    if (s->h_ctrl) cudaFreeHost(s->h_ctrl);
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int lbfgsb200_create_callback(lbfgsb200_solver_t **out, lbfgsb200_fg_device_fn fn, void *user, size_t n,
                              const lbfgsb200_params_t *params, size_t trace_rows)
{
    if (!fn) { set_error("create_callback: fn is NULL"); return LBFGSB200_ERR_INVALID; }
    int rc = lbfgsb200_create(out, LBFGSB200_OBJ_DEVICE_CALLBACK, n, params, nullptr, trace_rows);
    if (rc < 0) return rc;
    lbfgsb200_solver *s = *out;
    s->cb = fn;
    s->cb_user = user;
    cudaError_t e = cudaMalloc(&s->cb_buf, sizeof(double) * (s->stride + 8));
    if (e == cudaSuccess) e = cudaMemset(s->cb_buf, 0, sizeof(double) * (s->stride + 8));
    if (e != cudaSuccess) {
        set_error("create_callback: %s", cudaGetErrorString(e));
        cudaGetLastError();
        lbfgsb200_destroy(s);
        *out = nullptr;
        return LBFGSB200_ERR_NOMEM;
    }
    return 0;
}

size_t lbfgsb200_local_size(const lbfgsb200_solver_t *s) { return s ? s->n_local : 0; }

// ---- checkpoint / resume ------------------------------------------------------------------------
struct CkptHeader {
    char magic[8]; // "LBB200C1"
    uint64_t n_local, stride, nvecs, gram_doubles, trace_rows, devstate_bytes;
    int32_t m, objective, direction, profile, rank, nranks, x_is_alt, pad;
    int64_t k_host;
};

static size_t gram_doubles_of(const lbfgsb200_solver *s)
{
    if (!s->gram) return 0;
    const size_t nb = (size_t)(2 * s->nslots + 1), cnt = (size_t)3 * (2 * s->params.m + 1);
    const int nranks = s->comm ? s->comm->nranks : 1;
    return nb * nb + cnt * (size_t)(nranks + 1) + (size_t)(2 * s->params.m + 1);
}

static int stream_dev_to_file(FILE *f, const double *dev, size_t count, cudaStream_t st)
{
    std::vector<double> buf((size_t)1 << 22); // 32 MB staging
    for (size_t off = 0; off < count; off += buf.size()) {
        const size_t c = count - off < buf.size() ? count - off : buf.size();
        CUDA_TRY(cudaMemcpyAsync(buf.data(), dev + off, c * sizeof(double), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (fwrite(buf.data(), sizeof(double), c, f) != c) { set_error("checkpoint: short write"); return LBFGSB200_ERR_INVALID; }
    }
    return 0;
}

static int stream_file_to_dev(FILE *f, double *dev, size_t count, cudaStream_t st)
{
    std::vector<double> buf((size_t)1 << 22);
    for (size_t off = 0; off < count; off += buf.size()) {
        const size_t c = count - off < buf.size() ? count - off : buf.size();
        if (fread(buf.data(), sizeof(double), c, f) != c) { set_error("checkpoint: short read"); return LBFGSB200_ERR_INVALID; }
        CUDA_TRY(cudaMemcpyAsync(dev + off, buf.data(), c * sizeof(double), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}

int lbfgsb200_checkpoint_save(lbfgsb200_solver_t *s, const char *path)
{
    if (!s || !path || !s->x0_set) { set_error("checkpoint_save: no state to save"); return LBFGSB200_ERR_INVALID; }
    LB_TRY(snapshot(s));
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("checkpoint_save: cannot open %s", path); return LBFGSB200_ERR_INVALID; }
    CkptHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "LBB200C1", 8);
    h.n_local = s->n_local; h.stride = s->stride; h.nvecs = 4 + 2 * (size_t)s->nslots;
    h.gram_doubles = gram_doubles_of(s); h.trace_rows = s->trace_rows; h.devstate_bytes = sizeof(DevState);
    h.m = s->params.m; h.objective = s->objective; h.direction = s->params.direction; h.profile = s->params.profile;
    h.rank = s->comm ? s->comm->rank : 0; h.nranks = s->comm ? s->comm->nranks : 1;
    h.x_is_alt = (s->h_snapshot.x != s->arena);
    h.k_host = s->k_host;
    int rc = 0;
    if (fwrite(&h, sizeof h, 1, f) != 1 || fwrite(&s->h_snapshot, sizeof(DevState), 1, f) != 1) {
        set_error("checkpoint_save: short write");
        rc = LBFGSB200_ERR_INVALID;
    }
    if (!rc) rc = stream_dev_to_file(f, s->arena, h.nvecs * h.stride, s->stream);
    if (!rc && h.gram_doubles) rc = stream_dev_to_file(f, s->gram, h.gram_doubles, s->stream);
    if (!rc && s->trace_rows) rc = stream_dev_to_file(f, s->trace, s->trace_rows * LBFGSB200_TRACE_COLS, s->stream);
    fclose(f);
    return rc;
}

int lbfgsb200_checkpoint_load(lbfgsb200_solver_t *s, const char *path)
{
    if (!s || !path) return LBFGSB200_ERR_INVALID;
    FILE *f = fopen(path, "rb");
    if (!f) { set_error("checkpoint_load: cannot open %s", path); return LBFGSB200_ERR_INVALID; }
    CkptHeader h;
    DevState st;
    int rc = 0;
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "LBB200C1", 8) != 0 || h.devstate_bytes != sizeof(DevState) ||
        fread(&st, sizeof st, 1, f) != 1) {
        set_error("checkpoint_load: %s is not a checkpoint of this library version", path);
        rc = LBFGSB200_ERR_INVALID;
    }
    const int rank = s->comm ? s->comm->rank : 0, nranks = s->comm ? s->comm->nranks : 1;
    if (!rc && (h.n_local != s->n_local || h.stride != s->stride || h.m != s->params.m || h.objective != s->objective ||
                h.direction != s->params.direction || h.profile != s->params.profile || h.rank != rank || h.nranks != nranks ||
                h.gram_doubles != gram_doubles_of(s) || h.trace_rows != s->trace_rows)) {
        set_error("checkpoint_load: the checkpoint was written by a solver of a different shape");
        rc = LBFGSB200_ERR_INVALID;
    }
    if (!rc) rc = stream_file_to_dev(f, s->arena, h.nvecs * h.stride, s->stream);
    if (!rc && h.gram_doubles) rc = stream_file_to_dev(f, s->gram, h.gram_doubles, s->stream);
    if (!rc && s->trace_rows) rc = stream_file_to_dev(f, s->trace, s->trace_rows * LBFGSB200_TRACE_COLS, s->stream);
    fclose(f);
    if (rc) return rc;
    // scalars come from the file; every pointer and handle from this solver
    const DevState &cur = s->h_snapshot;
    st.x = h.x_is_alt ? s->arena + s->stride : s->arena;
    st.x_alt = h.x_is_alt ? s->arena : s->arena + s->stride;
    st.g = cur.g; st.w = cur.w; st.S = cur.S; st.Y = cur.Y;
    st.partials = cur.partials; st.send = cur.send; st.recv = cur.recv; st.trace = cur.trace;
    st.gram = cur.gram; st.gram_rows = cur.gram_rows; st.gram_recv = cur.gram_recv; st.delta = cur.delta;
    st.mail = cur.mail; st.peers = cur.peers; st.p2p = cur.p2p; st.p2p_timeout_ns = cur.p2p_timeout_ns;
    st.cond_outer = cur.cond_outer; st.cond_inner = cur.cond_inner; st.use_graph = cur.use_graph;
    st.tl = cur.tl; st.tl_cap = cur.tl_cap; st.tl_n = cur.tl_n;
    st.max_iterations = cur.max_iterations; st.tolerance = cur.tolerance; st.lsp = cur.lsp; // the new handle's limits apply
    if (st.status != LBFGSB200_CONVERGED && st.status != LBFGSB200_LS_FAILED && st.k < st.max_iterations) {
        st.status = LBFGSB200_RUNNING; // a run that only ran out of iterations may continue
        st.ctrl.done = 0;
    }
    s->h_snapshot = st;
    CUDA_TRY(cudaMemcpyAsync(s->d_st, &s->h_snapshot, sizeof(DevState), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->k_host = h.k_host;
    s->x_cur = st.x;
    s->x0_set = true;
    return 0;
}

int lbfgsb200_set_x0(lbfgsb200_solver_t *s, const double *x0_local)
{
    if (!s || !x0_local) { set_error("set_x0: NULL argument"); return LBFGSB200_ERR_INVALID; }
    // restore the pristine state (ring empty, pointers un-swapped), then evaluate f(x0), g(x0)
    DevState st = s->h_snapshot;
    st.x = s->arena;
    st.x_alt = s->arena + s->stride;
    st.base = 0;
    st.h = 0;
    st.k = 0;
    st.ctrl.ls_active = 0;
    st.ctrl.done = 0;
    st.ctrl.h = 0;
    st.ctrl.k = 0;
    st.status = LBFGSB200_RUNNING;
    st.use_graph = 0; // re-armed per run by do_iterate
    memset(&st.ls, 0, sizeof st.ls); // FLAVOR_PAR_INLINED carries state from one search to the next
    st.tl_n = 0;
    st.xL = st.xR = st.dL = st.dR = st.gL = st.gR = 0.0;
    CUDA_TRY(cudaMemcpyAsync(s->d_st, &st, sizeof st, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream)); // st is a stack object
    CUDA_TRY(cudaMemcpyAsync(st.x, x0_local, s->n_local * sizeof(double), cudaMemcpyDefault, s->stream));
    CUDA_TRY(cudaMemsetAsync(st.w, 0, s->stride * sizeof(double), s->stream)); // d = 0
    if (s->comm && s->comm->nranks > 1) {
        // neighbours' boundary x before the first evaluation (one-element halo)
        if (s->comm->p2p) {
            k_scalar<<<1, kScalarThreads, 0, s->stream>>>(s->d_st, -1, 0, 2, PACK_X0, 0);
            s->launches += 1;
        } else {
            k_pack<<<1, kScalarThreads, 0, s->stream>>>(s->d_st, -1, PACK_X0, 0);
            LB_TRY(comm_allgather(s->comm, s->pkt, s->pkt + kPacket, kPacket, s->stream));
            k_scalar<<<1, kScalarThreads, 0, s->stream>>>(s->d_st, -1, 0, 1, PACK_X0, 0);
            s->launches += 2;
        }
    }
    if (s->cb) {
        double *g_trial = s->cb_buf, *scal = s->cb_buf + s->stride, *d_zero = scal + 4;
        if (s->cb(st.x, st.w, d_zero, g_trial, scal, s->n_local, 0, s->cb_user, s->stream)) {
            set_error("the objective callback failed");
            return LBFGSB200_ERR_INVALID;
        }
        k_accept_generic<<<s->grid, kThreads, 0, s->stream>>>(s->d_st, g_trial, scal, 1);
        s->launches += 1;
        LB_TRY(scalar_step(s, OP_INIT, 0, PACK_ACCEPT, s->grid));
        s->x_cur = s->arena + s->stride; // OP_INIT swapped x and x_alt
    } else {
        s->accept_kernel<<<s->grid_accept, kThreads, 0, s->stream>>>(s->d_st, 1);
        s->launches += 1;
        LB_TRY(scalar_step(s, OP_INIT, 0, PACK_ACCEPT));
    }
    CUDA_TRY(cudaGetLastError());
    LB_TRY(snapshot(s));
    s->k_host = 0;
    s->x0_set = true;
    return 0;
}

int lbfgsb200_iterate(lbfgsb200_solver_t *s, int64_t iterations)
{
    if (!s) return LBFGSB200_ERR_INVALID;
    s->profiling = false;
    return do_iterate(s, iterations);
}

int lbfgsb200_iterate_profiled(lbfgsb200_solver_t *s, int64_t iterations,
                               double class_ms[LBFGSB200_PROFILE_CLASSES],
                               int64_t class_launches[LBFGSB200_PROFILE_CLASSES])
{
    if (!s) return LBFGSB200_ERR_INVALID;
    for (int c = 0; c < LBFGSB200_PROFILE_CLASSES; ++c) {
        for (cudaEvent_t e : s->prof_ev[c]) cudaEventDestroy(e);
        s->prof_ev[c].clear();
    }
    s->profiling = true;
    int rc = do_iterate(s, iterations);
    s->profiling = false;
    if (rc < 0) return rc;
    for (int c = 0; c < LBFGSB200_PROFILE_CLASSES; ++c) {
        double total = 0.0;
        const std::vector<cudaEvent_t> &v = s->prof_ev[c];
        for (size_t i = 0; i + 1 < v.size(); i += 2) {
            float ms = 0.f;
            CUDA_TRY(cudaEventElapsedTime(&ms, v[i], v[i + 1]));
            total += ms;
        }
        if (class_ms) class_ms[c] = total;
        if (class_launches) class_launches[c] = (int64_t)(v.size() / 2);
    }
    return rc;
}

int lbfgsb200_get_x(lbfgsb200_solver_t *s, double *x_local_out)
{
    if (!s || !x_local_out) return LBFGSB200_ERR_INVALID;
    CUDA_TRY(cudaMemcpyAsync(x_local_out, s->h_snapshot.x, s->n_local * sizeof(double), cudaMemcpyDefault, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    return 0;
}

int lbfgsb200_get_result(lbfgsb200_solver_t *s, lbfgsb200_result_t *r)
{
    if (!s || !r) return LBFGSB200_ERR_INVALID;
    const DevState &st = s->h_snapshot;
    r->status = st.status;
    r->iterations = st.k;
    r->trial_evals = st.trial_evals;
    r->kernel_launches = s->launches;
    r->f = st.f;
    r->gnorm = sqrt(st.gg);
    r->device_ms = s->last_ms;
    r->bytes_moved = s->streams_last * 8.0 * (double)s->n_local;
    r->f0 = st.f0;
    r->gnorm0 = sqrt(st.gg0);
    return 0;
}

int64_t lbfgsb200_get_trace(lbfgsb200_solver_t *s, double *rows, size_t max_rows)
{
    if (!s || !rows || !s->trace) return 0;
    size_t have = (size_t)s->h_snapshot.k;
    if (have > s->trace_rows) have = s->trace_rows;
    if (have > max_rows) have = max_rows;
    if (have == 0) return 0;
    if (cudaMemcpy(rows, s->trace, have * LBFGSB200_TRACE_COLS * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        return LBFGSB200_ERR_CUDA;
    return (int64_t)have;
}

int lbfgsb200_solve(int objective, size_t n, const double *x0_host, double *x_out_host,
                    const lbfgsb200_params_t *params, lbfgsb200_result_t *result, double *trace,
                    size_t trace_rows)
{
    if (!x0_host || !x_out_host || !params) { set_error("solve: NULL argument"); return LBFGSB200_ERR_INVALID; }
    lbfgsb200_solver *s = nullptr;
    int rc = lbfgsb200_create(&s, objective, n, params, nullptr, trace ? trace_rows : 0);
    if (rc < 0) return rc;
    rc = lbfgsb200_set_x0(s, x0_host);
    if (rc >= 0) rc = lbfgsb200_iterate(s, (int64_t)params->max_iterations + 1);
    if (rc >= 0) {
        int rc2 = lbfgsb200_get_x(s, x_out_host);
        if (rc2 < 0) rc = rc2;
    }
    if (rc >= 0 && result) lbfgsb200_get_result(s, result);
    if (rc >= 0 && trace) lbfgsb200_get_trace(s, trace, trace_rows);
    if (rc >= 0 && params->verbose) {
        // the reference's per-iteration line (seq/lbfgs.cpp:77-78: printed at the top of iteration k
        // with the current f and |grad|), reproduced from the device trace
        lbfgsb200_result_t r;
        lbfgsb200_get_result(s, &r);
        std::vector<double> rows((size_t)LBFGSB200_TRACE_COLS * (trace_rows ? trace_rows : 1));
        const int64_t got = trace ? lbfgsb200_get_trace(s, rows.data(), trace_rows) : 0;
        const int64_t lines = r.status == LBFGSB200_MAX_ITER ? r.iterations : r.iterations + 1;
        for (int64_t k = 0; k < lines; ++k) {
            if (k == 0) printf("Iteration 0, f = %g, |grad| = %g\n", r.f0, r.gnorm0);
            else if (k - 1 < got)
                printf("Iteration %lld, f = %g, |grad| = %g\n", (long long)k, rows[(k - 1) * LBFGSB200_TRACE_COLS + 1],
                       rows[(k - 1) * LBFGSB200_TRACE_COLS + 2]);
        }
    }
    lbfgsb200_destroy(s);
    return rc;
}

// --------------------------------------------------------------------
// unit-test surface
// --------------------------------------------------------------------
struct Scratch {
    double *p = nullptr;
    cudaStream_t st;
    int grid = 1;
    int init(long long n, cudaStream_t stream, size_t extra_doubles = 0, int ctas_per_sm = kCtasPerSm)
    {
        st = stream;
        int sms = 0;
        LB_TRY(sm_count(&sms));
        grid = pick_grid(n, sms, 0, ctas_per_sm);
        CUDA_TRY(cudaMallocAsync(&p, sizeof(double) * (kMaxQ * (size_t)grid + extra_doubles), st));
        return 0;
    }
    double *extra() { return p + kMaxQ * (size_t)grid; }
    ~Scratch() { if (p) cudaFreeAsync(p, st); }
};

int lbfgsb200_dot(const double *a, const double *b, size_t n, double *d_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc;
    LB_TRY(sc.init((long long)n, st));
    k_dot<<<sc.grid, kThreads, 0, st>>>(a, b, (long long)n, sc.p);
    k_finalize<<<1, kScalarThreads, 0, st>>>(sc.p, sc.grid, 1, d_out, 0);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int lbfgsb200_nrm2(const double *a, size_t n, double *d_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc;
    LB_TRY(sc.init((long long)n, st));
    k_dot<<<sc.grid, kThreads, 0, st>>>(a, a, (long long)n, sc.p);
    k_finalize<<<1, kScalarThreads, 0, st>>>(sc.p, sc.grid, 1, d_out, 1);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int lbfgsb200_axpy(const double *d_alpha, const double *x, double *y, size_t n, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0;
    LB_TRY(sm_count(&sms));
    k_axpy<true><<<pick_grid((long long)n, sms, 0), kThreads, 0, st>>>(d_alpha, x, y, y, (long long)n);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int lbfgsb200_scal(const double *d_alpha, const double *x, double *out, size_t n, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    int sms = 0;
    LB_TRY(sm_count(&sms));
    k_axpy<false><<<pick_grid((long long)n, sms, 0), kThreads, 0, st>>>(d_alpha, x, nullptr, out, (long long)n);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int lbfgsb200_eval_trial(int objective, const double *x, const double *d, const double *d_alpha,
                         size_t n, double *g_out, double *d_out3, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc;
    LB_TRY(sc.init((long long)n, st, 0, kCtasPerSmAccept));
    k_eval_explicit<<<sc.grid, kThreads, 0, st>>>(MODE_TRIAL, objective, x, d, d_alpha, (long long)n,
                                                   g_out, nullptr, nullptr, nullptr, nullptr, sc.p);
    k_finalize<<<1, kScalarThreads, 0, st>>>(sc.p, sc.grid, 3, d_out3, 0);
    CUDA_TRY(cudaGetLastError());
    return 0;
}

int lbfgsb200_accept(int objective, double *x, const double *d, double *g, const double *d_alpha,
                     size_t n, double *s_out, double *y_out, double *d_out5, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    Scratch sc;
    const size_t stride = (n + 31) / 32 * 32;
    LB_TRY(sc.init((long long)n, st, stride, kCtasPerSmAccept));
    double *x_new = sc.extra(); // the kernel never writes the new iterate over x (see kernels.cuh)
    k_eval_explicit<<<sc.grid, kThreads, 0, st>>>(MODE_ACCEPT, objective, x, d, d_alpha, (long long)n,
                                                   nullptr, x_new, g, s_out, y_out, sc.p);
    k_finalize<<<1, kScalarThreads, 0, st>>>(sc.p, sc.grid, 5, d_out5, 0);
    CUDA_TRY(cudaMemcpyAsync(x, x_new, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// helper kernels of lbfgsb200_two_loop: fill the per-slot scalars of a scratch DevState from
// two dot products, then run the production direction phase on it
__global__ void k_set_pair(DevState *st, int slot)
{
    __shared__ double r[kMaxQ];
    reduce_partials(st->partials, st->grid, 2, r);
    if (threadIdx.x == 0) {
        st->sy[slot] = r[0];
        st->yy[slot] = r[1];
        st->rho[slot] = 1.0 / r[0];
        st->skip[slot] = 0;
    }
}
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
k_pair_dots(const double *__restrict__ s, const double *__restrict__ y, long long n, double *partials)
{
    double a0 = 0.0, a1 = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x) {
        a0 += s[i] * y[i];
        a1 += y[i] * y[i];
    }
    double v[2] = {a0, a1};
    block_emit<2>(v, partials);
}
__global__ void k_two_loop_out(const DevState *st, double *d_out2)
{
    d_out2[0] = st->gd;
    d_out2[1] = (double)st->steepest;
}

__global__ void k_set_gg(DevState *st)
{
    __shared__ double r[kMaxQ];
    reduce_partials(st->partials, st->grid, 1, r);
    if (threadIdx.x == 0) st->gg = r[0];
}

int lbfgsb200_two_loop(const double *g, const double *S, const double *Y, int h, size_t n,
                       size_t stride, double *d, double *d_out2, void *stream)
{
    if (h < 0 || h > LBFGSB200_MAX_M) { set_error("two_loop: h out of range"); return LBFGSB200_ERR_INVALID; }
    cudaStream_t stream_ = (cudaStream_t)stream;
    lbfgsb200_solver tmp; // borrowed launch context: owns nothing
    memset(&tmp.params, 0, sizeof tmp.params);
    tmp.stream = stream_;
    LB_TRY(sm_count(&tmp.sms));
    tmp.grid = pick_grid((long long)n, tmp.sms, 0);
    tmp.params.m = h > 0 ? h : 1;
    tmp.k_host = h;
    double *partials = nullptr;
    DevState *d_st = nullptr;
    CUDA_TRY(cudaMallocAsync(&partials, sizeof(double) * kMaxQ * (size_t)tmp.grid, stream_));
    CUDA_TRY(cudaMallocAsync(&d_st, sizeof(DevState), stream_));
    DevState st;
    memset(&st, 0, sizeof st);
    st.n = (long long)n;
    st.nglob = (long long)n;
    st.m = tmp.params.m;
    st.nslots = st.m + 1;
    st.profile = LBFGSB200_PROFILE_SEQ;
    st.max_iterations = 1 << 30;
    st.tolerance = 0.0;
    st.nranks = 1;
    st.grid = tmp.grid;
    st.g = const_cast<double *>(g);
    st.w = d;
    st.S = const_cast<double *>(S);
    st.Y = const_cast<double *>(Y);
    st.stride = (long long)stride;
    st.partials = partials;
    st.h = h;
    st.k = 1;        // not the first iteration: use the history
    st.sg_valid = 0; // exercises the stand-alone s.g path
    CUDA_TRY(cudaMemcpyAsync(d_st, &st, sizeof st, cudaMemcpyHostToDevice, stream_));
    CUDA_TRY(cudaStreamSynchronize(stream_)); // st is a stack object
    k_dot<<<tmp.grid, kThreads, 0, stream_>>>(g, g, (long long)n, partials);
    k_set_gg<<<1, kScalarThreads, 0, stream_>>>(d_st);
    for (int i = 0; i < h; ++i) {
        k_pair_dots<<<tmp.grid, kThreads, 0, stream_>>>(S + (size_t)i * stride, Y + (size_t)i * stride, (long long)n, partials);
        k_set_pair<<<1, kScalarThreads, 0, stream_>>>(d_st, i);
    }
    tmp.d_st = d_st;
    int rc = launch_direction(&tmp);
    k_two_loop_out<<<1, 1, 0, stream_>>>(d_st, d_out2);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(partials, stream_);
    cudaFreeAsync(d_st, stream_);
    if (rc < 0) return rc;
    if (e != cudaSuccess) { set_error("two_loop launch: %s", cudaGetErrorString(e)); return LBFGSB200_ERR_CUDA; }
    return 0;
}

// pinned host buffers for callers without their own CUDA runtime access (the Python harness)
void *lbfgsb200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        set_error("cudaHostAlloc(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void lbfgsb200_host_free(void *p) { if (p) cudaFreeHost(p); }
void *lbfgsb200_device_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        set_error("cudaMalloc(%zu) failed", bytes);
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void lbfgsb200_device_free(void *p) { if (p) cudaFree(p); }
int lbfgsb200_memcpy(void *dst, const void *src, size_t bytes)
{
    CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return 0;
}
long lbfgsb200_debug_timeline(lbfgsb200_solver_t *s, unsigned long long *rows, size_t cap_rows, int reset)
{
    if (!s || !s->timeline) { set_error("debug_timeline: create the solver with LBFGSB200_TIMELINE=<rows> set"); return LBFGSB200_ERR_INVALID; }
    LB_TRY(snapshot(s));
    size_t nrows = (size_t)s->h_snapshot.tl_n;
    if (nrows > cap_rows) nrows = cap_rows;
    if (rows && nrows) CUDA_TRY(cudaMemcpy(rows, s->timeline, sizeof(unsigned long long) * 3 * nrows, cudaMemcpyDeviceToHost));
    if (reset) {
        const int zero = 0;
        CUDA_TRY(cudaMemcpy((char *)s->d_st + offsetof(DevState, tl_n), &zero, sizeof zero, cudaMemcpyHostToDevice));
    }
    return (long)nrows;
}
int lbfgsb200_trim_memory(void)
{
    int dev = 0;
    cudaMemPool_t pool;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, dev));
    CUDA_TRY(cudaMemPoolTrimTo(pool, 0));
    return 0;
}
int lbfgsb200_mem_info(size_t *free_bytes, size_t *total_bytes)
{
    size_t f = 0, t = 0;
    CUDA_TRY(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return 0;
}
int lbfgsb200_set_device(int ordinal)
{
    CUDA_TRY(cudaSetDevice(ordinal));
    return 0;
}
int lbfgsb200_device_sync(void)
{
    CUDA_TRY(cudaDeviceSynchronize());
    return 0;
}

} // extern "C"
