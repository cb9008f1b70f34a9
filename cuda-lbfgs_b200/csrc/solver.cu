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
#include <nvtx3/nvToolsExt.h> // header-only NVTX v3: no link dependency, a no-op unless a profiler is attached
#include <math.h>
#include <stdarg.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lbfgsb200.h"
#include "accept_gram.cuh"
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

// ---- device memory: a PRIVATE stream-ordered pool per device --------------------------------------------
// Solver arenas ((2m+6) vectors, tens of GB) are served from a pool this library creates itself, with the
// release threshold raised on THAT pool only, so destroying a solver and creating the next one of similar
// size re-uses the mapping instead of paying cudaMalloc/cudaFree of tens of GB (5-130 ms each way).  The
// device's default pool -- which torch, NCCL or any other cudaMallocAsync user of the host process shares --
// is never touched.  Every allocation of the library goes through pool_alloc(), which on an out-of-memory
// answer hands the cached blocks back to the driver (cudaMemPoolTrimTo) and tries once more;
// lbfgsb200_trim_memory() does the same on request.
constexpr int kMaxDevices = 64;
static cudaMemPool_t g_pools[kMaxDevices] = {};
static std::mutex g_pool_mutex;
static int s_optin[kMaxDevices] = {}; // cudaDevAttrMaxSharedMemoryPerBlockOptin per device

static int private_pool(cudaMemPool_t *out)
{
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) { set_error("device ordinal %d out of range", dev); return LBFGSB200_ERR_INVALID; }
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    if (!g_pools[dev]) {
        cudaMemPoolProps props;
        memset(&props, 0, sizeof props);
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool;
        CUDA_TRY(cudaMemPoolCreate(&pool, &props));
        uint64_t keep = UINT64_MAX;
        CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        g_pools[dev] = pool;
    }
    *out = g_pools[dev];
    return 0;
}

// stream-ordered allocation from the private pool; trims the pool and retries once when memory is short
static int pool_alloc(void **ptr, size_t bytes, cudaStream_t stream)
{
    cudaMemPool_t pool;
    LB_TRY(private_pool(&pool));
    cudaError_t e = cudaMallocFromPoolAsync(ptr, bytes, pool, stream);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        cudaStreamSynchronize(stream);
        cudaMemPoolTrimTo(pool, 0);
        e = cudaMallocFromPoolAsync(ptr, bytes, pool, stream);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        *ptr = nullptr;
        set_error("device allocation of %.3f GB failed: %s", bytes / 1e9, cudaGetErrorString(e));
        return e == cudaErrorMemoryAllocation ? LBFGSB200_ERR_NOMEM : LBFGSB200_ERR_CUDA;
    }
    return 0;
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
    int gram_tma = 0, gram_NG = 2, gram_NS = 2, gram_G = 1; // stand-alone pass A: tensor-map TMA (2, default) or cp.async pipeline (0)
    size_t gram_smem = 0;
    ArenaMaps *arena_maps = nullptr;    // device: tensor maps over the arena, boxes gram_T wide (pass A, fused accept + pass A)
    ArenaMaps *arena_maps_ct = nullptr; // device: the same with boxes ct_T wide (k_combine_trial)
    ArenaMaps arena_maps_host[2];       // staging copies (the uploads are stream-ordered)
    int ct_T = 256, ct_NS = 4, ct_halo = kCtHaloItems; // k_combine_trial: tile width, stages, overlap
    double *gram = nullptr;     // compact form: Gram matrix + pass-A rows + delta + all-gather buffer
    // fused compact flow (accept_gram.cuh): k_accept_gram + k_combine_trial
    bool fused = false;
    accept_gram_kernel_t ag_kernel = nullptr;
    combine_trial_kernel_t ct_kernel = nullptr;
    size_t ag_smem = 0, ct_smem = 0;
    int grid_ag = 1;
    int device = 0;
    void *aux = nullptr;        // ONE allocation for every small device buffer below (partials ... d_st)
    size_t gram_doubles = 0;
    lbfgsb200_params_t params;
    lbfgsb200_comm *comm = nullptr;
    trial_kernel_t trial_kernel = nullptr;   // objective-specific instantiations
    lbfgsb200_fg_device_fn cb = nullptr;     // user device objective (lbfgsb200_create_callback)
    void *cb_user = nullptr;
    double *cb_buf = nullptr;                // g_trial [stride] + {f, g.d, g.g} + a device zero
    int ag_boxes = 3;                        // k_accept_gram box layout: bit 0 merged halo boxes, bit 1 merged {x, g, d} box
    bool cb_graph_failed = false;            // the callback could not be stream-captured: host-stepped loop instead
    accept_kernel_t accept_kernel = nullptr;

    double *arena = nullptr;    // rows x_a, g, w, x_b, S[nslots], Y[nslots] (compact.cuh: kArenaRow*)
    double *partials = nullptr; // [kMaxQ][grid]
    double *pkt = nullptr;      // send [kPacket] + recv [nranks][kPacket]
    double *trace = nullptr;
    unsigned long long *timeline = nullptr; // LBFGSB200_TIMELINE diagnostic (DevState::tl)
    size_t trace_rows = 0;
    DevState *d_st = nullptr;
    Ctrl *h_ctrl = nullptr;     // pinned
    DevState h_snapshot;        // last state copied back

    cudaStream_t stream = nullptr;
    cudaStream_t capture_stream = nullptr; // graph recording only
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long cond_handles[3] = {0, 0, 0}; // WHILE / WHILE / IF handles (DevState::cond_outer, _inner, _fix)
    int cond_flag = 0;

    bool x0_set = false;
    int64_t k_host = 0; // accepts launched so far: an upper bound on the device's h
    int64_t launches = 0;
    bool last_graph = false;
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

static bool is_multi(const lbfgsb200_solver *s) { return s->comm && s->comm->nranks > 1; }

// NVTX ranges (nsys / ncu timelines): one per API call, and in the host-stepped loops one per phase of an iteration.
// Inside a CUDA graph the phases are kernel nodes and show up by kernel name.
struct NvtxRange {
    explicit NvtxRange(const char *name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

// scalar step: on one GPU the scalar kernel sums the partials itself; on several the local sums
// and halo values are packed, exchanged (NVLink mailboxes inside the scalar kernel, or one small NCCL
// all-gather) and summed in rank order.
// LBFGSB200_DEBUG_SYNC=1 (host-stepped loop only): synchronise after every launch and name the kernel that failed
static int debug_sync(lbfgsb200_solver *s, const char *what, int op = -100)
{
    static const bool on = getenv("LBFGSB200_DEBUG_SYNC") != nullptr;
    if (!on) return 0;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s->stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return 0;
    const cudaError_t e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        set_error("%s (op %d) failed: %s [k=%lld]", what, op, cudaGetErrorString(e), (long long)s->k_host);
        fprintf(stderr, "lbfgsb200: %s\n", lbfgsb200_last_error());
        return LBFGSB200_ERR_CUDA;
    }
    return 0;
}

static int scalar_step(lbfgsb200_solver *s, int op, int p, int pack_kind, int nparts_override = -1)
{
    const int nparts = nparts_override >= 0 ? nparts_override
                       : (op == OP_ACCEPT || op == OP_INIT) ? s->grid_accept
                       : (op == OP_COMPACT_DIR || op == OP_F_DIR) ? s->grid_combine : s->grid;
    const bool needs_data = (op != OP_ITER_BEGIN && op != OP_LS_INIT && op != OP_F_BEGIN);
    if (is_multi(s) && needs_data) {
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
    return debug_sync(s, "k_scalar", op);
}

// the stand-alone pass A (old compact flow; fused flow: only after a rejected pair with a full ring)
static void launch_gram(lbfgsb200_solver *s)
{
    const int m = s->params.m;
    const size_t smem = s->gram_smem;
    if (s->gram_tma == 2) {
        if ((2 * m + 1 + s->gram_NG - 1) / s->gram_NG <= 7)
            k_gram_tma2d<7><<<s->grid_gram, kWsThreads, smem, s->stream>>>(s->d_st, s->arena_maps, s->gram_T, s->gram_NG, s->gram_NS);
        else
            k_gram_tma2d<kMaxCW><<<s->grid_gram, kWsThreads, smem, s->stream>>>(s->d_st, s->arena_maps, s->gram_T, s->gram_NG, s->gram_NS);
    } else {
        const int G = s->gram_G, per = (2 * m + 1 + G - 1) / G, cwg = (per + kGramWarps - 1) / kGramWarps;
        const dim3 grid(s->grid_gram, G);
        if (cwg <= 3) k_gram<3><<<grid, kThreads, smem, s->stream>>>(s->d_st, s->gram_T, s->gram_NS, G);
        else if (cwg <= 6) k_gram<6><<<grid, kThreads, smem, s->stream>>>(s->d_st, s->gram_T, s->gram_NS, G);
        else k_gram<kMaxCW><<<grid, kThreads, smem, s->stream>>>(s->d_st, s->gram_T, s->gram_NS, G);
    }
    s->launches += 1;
}

// scalar kernel of the compact ops that carry pass-A rows (OP_COMPACT, OP_F_INIT, OP_F_ACCEPT, OP_F_FIX): dynamic
// shared memory = the (2m+1)^2 window Gram matrix of the coefficient recursion
static int rows_step(lbfgsb200_solver *s, int op, int nparts)
{
    const int J = 2 * s->params.m + 1;
    const size_t dyn = sizeof(double) * (size_t)J * J;
    const bool multi = is_multi(s), p2p = multi && s->comm->p2p;
    if (multi && !p2p) {
        // NCCL path: the rows must be in HBM for the all-gather (otherwise the scalar kernel sums them itself)
        if (op == OP_COMPACT) k_gram_finalize<<<3 * J, kScalarThreads, 0, s->stream>>>(s->d_st, nparts);
        else k_pack_rows<<<1, kScalarThreads, 0, s->stream>>>(s->d_st, op, nparts);
        s->launches += 1;
        LB_TRY(comm_allgather(s->comm, s->h_snapshot.gram_rows, s->h_snapshot.gram_recv, s->h_snapshot.gram_count, s->stream));
    }
    k_scalar<<<1, kScalarThreads, dyn, s->stream>>>(s->d_st, op, 0, p2p ? 2 : (multi ? 1 : 0), PACK_NONE, nparts);
    s->launches += 1;
    return debug_sync(s, "k_scalar(rows)", op);
}

// search direction: seq/lbfgs.cpp:86-153 (explicit two-loop recursion, or the UNFUSED compact form, which user
// objectives still run on; built-in objectives with direction = compact take the fused flow below)
static int launch_direction(lbfgsb200_solver *s)
{
    const int m = s->params.m;
    const int h_upper = (int)(s->k_host < m ? s->k_host : m);
    LB_TRY(scalar_step(s, OP_ITER_BEGIN, 0, PACK_NONE));
    if (h_upper > 0 && s->params.direction == LBFGSB200_DIR_COMPACT) {
        // compact form: pass A (Gram rows) -> coefficient recursion -> pass B (combine)
        {
            ClassTimer t(s, KC_GRAM);
            launch_gram(s);
        }
        LB_TRY(rows_step(s, OP_COMPACT, s->grid_gram));
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
    *s->h_ctrl = s->h_snapshot.ctrl;
    return 0;
}

// host-stepped iteration loop
static int run_stepped(lbfgsb200_solver *s, int64_t iterations)
{
    for (int64_t it = 0; it < iterations; ++it) {
        {
            NvtxRange r("lbfgsb200:direction");
            LB_TRY(launch_direction(s));
        }
        NvtxRange r("lbfgsb200:line_search+accept");
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

// ---- user (callback) objectives ----------------------------------------------------------------
// The callback replaces k_trial (one call per line-search trial) and the evaluation half of the accept kernel (one
// more call at the accepted step, followed by k_accept_generic).  x is updated IN PLACE for these solvers
// (DevState::x_alt == x: the generic accept kernel is element-wise), so every pointer the callback receives is the
// same in every iteration and the calls can be recorded into the CUDA graph like the built-in kernels.
static int callback_eval(lbfgsb200_solver *s, const double *d_alpha, double *d_out3)
{
    if (s->cb(s->arena, s->h_snapshot.w, d_alpha, s->cb_buf, d_out3, s->n_local, s->offset, s->cb_user, s->stream)) {
        if (getenv("LBFGSB200_VERBOSE")) fprintf(stderr, "lbfgsb200: the objective callback returned non-zero (last CUDA error: %s)\n", cudaGetErrorName(cudaPeekAtLastError()));
        set_error("the objective callback failed");
        return LBFGSB200_ERR_INVALID;
    }
    return 0;
}

static int callback_trial_segment(lbfgsb200_solver *s)
{
    {
        ClassTimer t(s, KC_TRIAL);
        LB_TRY(callback_eval(s, &s->d_st->ls.alpha, s->partials));
    }
    return scalar_step(s, OP_LS_STEP, 0, PACK_NONE, 1); // the 3 sums of this shard are final: one "partial" each
}

static int callback_accept_segment(lbfgsb200_solver *s, int init)
{
    double *scal = s->cb_buf + s->stride, *d_zero = scal + 4;
    {
        // gradient and f at the accepted step (the last trial may have been at another alpha)
        ClassTimer t(s, KC_ACCEPT);
        LB_TRY(callback_eval(s, init ? d_zero : &s->d_st->ls.alpha, scal));
        k_accept_generic<<<s->grid, kThreads, 0, s->stream>>>(s->d_st, s->cb_buf, scal, init);
        s->launches += 1;
    }
    return scalar_step(s, init ? OP_INIT : OP_ACCEPT, 0, PACK_ACCEPT, s->grid);
}

// host-stepped loop for user objectives that cannot be captured (or use_graph = 0)
static int run_stepped_callback(lbfgsb200_solver *s, int64_t iterations)
{
    for (int64_t it = 0; it < iterations; ++it) {
        LB_TRY(launch_direction(s));
        LB_TRY(read_ctrl(s));
        while (s->h_ctrl->ls_active && !s->h_ctrl->done) {
            LB_TRY(callback_trial_segment(s));
            LB_TRY(read_ctrl(s));
        }
        if (s->h_ctrl->done) break;
        LB_TRY(callback_accept_segment(s, 0));
        s->k_host += 1;
    }
    CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---- fused compact flow (accept_gram.cuh) ---------------------------------------------------------
//   [ stand-alone pass A + OP_F_FIX, only after a rejected pair with a full ring ]
//   k_combine_trial (d, g.d, first trial)      -> OP_F_DIR    (safeguard, search start, first decision)
//   { k_trial -> OP_LS_STEP }                   further trials, if the search wants them
//   k_accept_gram (accept + next pass A)       -> OP_F_ACCEPT (bookkeeping, Gram update, next coefficients)
static void launch_accept_gram(lbfgsb200_solver *s, int init)
{
    ClassTimer t(s, KC_GRAM);
    s->ag_kernel<<<s->grid_ag, kAgThreads, s->ag_smem, s->stream>>>(s->d_st, s->arena_maps, s->gram_T, s->gram_NS, init | (s->ag_boxes << 1));
    s->launches += 1;
    debug_sync(s, "k_accept_gram", init);
}

static int fused_fix_segment(lbfgsb200_solver *s)
{
    launch_gram(s);
    return rows_step(s, OP_F_FIX, s->grid_gram);
}

static int fused_direction_segment(lbfgsb200_solver *s)
{
    {
        ClassTimer t(s, KC_COMBINE);
        s->ct_kernel<<<s->grid_combine, kCtThreads, s->ct_smem, s->stream>>>(s->d_st, s->arena_maps_ct, s->ct_T, s->ct_NS, s->ct_halo);
        s->launches += 1;
        debug_sync(s, "k_combine_trial");
    }
    return scalar_step(s, OP_F_DIR, 0, PACK_NONE);
}

static int fused_trial_segment(lbfgsb200_solver *s)
{
    {
        ClassTimer t(s, KC_TRIAL);
        s->trial_kernel<<<s->grid, kThreads, 0, s->stream>>>(s->d_st);
        s->launches += 1;
    }
    return scalar_step(s, OP_LS_STEP, 0, PACK_NONE);
}

static int fused_accept_segment(lbfgsb200_solver *s, int init)
{
    launch_accept_gram(s, init);
    return rows_step(s, init ? OP_F_INIT : OP_F_ACCEPT, s->grid_ag);
}

static int run_stepped_fused(lbfgsb200_solver *s, int64_t iterations)
{
    // s->h_ctrl is current: set_x0 / the previous run ended with a snapshot
    for (int64_t it = 0; it < iterations && !s->h_ctrl->done; ++it) {
        {
            NvtxRange r("lbfgsb200:direction+first_trial");
            if (s->h_ctrl->need_fix) LB_TRY(fused_fix_segment(s));
            LB_TRY(fused_direction_segment(s));
            LB_TRY(read_ctrl(s));
        }
        {
            NvtxRange r("lbfgsb200:line_search");
            while (s->h_ctrl->ls_active && !s->h_ctrl->done) {
                LB_TRY(fused_trial_segment(s));
                LB_TRY(read_ctrl(s));
            }
        }
        if (s->h_ctrl->done) break;
        NvtxRange r("lbfgsb200:accept+gram");
        LB_TRY(fused_accept_segment(s, 0));
        LB_TRY(read_ctrl(s));
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

// The graph is recorded on a stream of its own: the solver stream may still be busy (the upload of x0), and
// recording does not execute anything.
template <class F>
static int capture_segment(lbfgsb200_solver *s, cudaGraph_t g, std::vector<cudaGraphNode_t> &tail, F &&body)
{
    cudaStream_t run_stream = s->stream;
    s->stream = s->capture_stream;
    cudaError_t e0 = cudaStreamBeginCaptureToGraph(s->stream, g, tail.empty() ? nullptr : tail.data(), nullptr, tail.size(),
                                                   cudaStreamCaptureModeThreadLocal);
    if (e0 != cudaSuccess) {
        s->stream = run_stream;
        GRAPH_TRY(e0);
    }
    int rc = body();
    cudaStreamCaptureStatus status;
    const cudaGraphNode_t *deps = nullptr;
    size_t ndeps = 0;
    cudaError_t e = cudaStreamGetCaptureInfo_v2(s->stream, &status, nullptr, nullptr, &deps, &ndeps);
    if (e == cudaSuccess && status == cudaStreamCaptureStatusActive) tail.assign(deps, deps + ndeps);
    cudaGraph_t out = nullptr;
    cudaError_t e2 = cudaStreamEndCapture(s->stream, &out);
    s->stream = run_stream;
    if (rc < 0) return rc;
    GRAPH_TRY(e);
    GRAPH_TRY(e2);
    return 0;
}

static int add_conditional(cudaGraph_t parent, std::vector<cudaGraphNode_t> &tail, cudaGraphConditionalHandle h,
                           cudaGraphConditionalNodeType type, cudaGraph_t *body)
{
    cudaGraphNodeParams p = {};
    p.type = cudaGraphNodeTypeConditional;
    p.conditional.handle = h;
    p.conditional.type = type;
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
    cudaGraphConditionalHandle h_outer, h_inner, h_fix = 0;
    if (s->fused) {
        // every condition is armed by the prologue kernel / the scalar kernels from the device state
        GRAPH_TRY(cudaGraphConditionalHandleCreate(&h_outer, s->graph, 0, 0));
        GRAPH_TRY(cudaGraphConditionalHandleCreate(&h_inner, s->graph, 0, 0));
        GRAPH_TRY(cudaGraphConditionalHandleCreate(&h_fix, s->graph, 0, 0));
    } else {
        GRAPH_TRY(cudaGraphConditionalHandleCreate(&h_outer, s->graph, 1, cudaGraphCondAssignDefault));
        GRAPH_TRY(cudaGraphConditionalHandleCreate(&h_inner, s->graph, 0, 0));
    }
    s->h_snapshot.cond_outer = h_outer;
    s->h_snapshot.cond_inner = h_inner;
    s->h_snapshot.cond_fix = h_fix;
    s->cond_handles[0] = h_outer;
    s->cond_handles[1] = h_inner;
    s->cond_handles[2] = h_fix; // uploaded (stream-ordered) by do_iterate before every launch

    std::vector<cudaGraphNode_t> top_tail, tail;
    cudaGraph_t iter_body = nullptr, trial_body = nullptr;
    const int64_t k_saved = s->k_host, l_saved = s->launches;
    struct Restore { // recording counts launches and pretends the history is full: undo both on every exit
        lbfgsb200_solver *s;
        int64_t k, l;
        ~Restore() { s->k_host = k; s->launches = l; }
    } restore{s, k_saved, l_saved};
    if (s->fused) {
        LB_TRY(capture_segment(s, s->graph, top_tail, [&]() { return scalar_step(s, OP_F_BEGIN, 0, PACK_NONE); }));
        LB_TRY(add_conditional(s->graph, top_tail, h_outer, cudaGraphCondTypeWhile, &iter_body));
        cudaGraph_t fix_body = nullptr;
        LB_TRY(add_conditional(iter_body, tail, h_fix, cudaGraphCondTypeIf, &fix_body));
        std::vector<cudaGraphNode_t> fix_tail;
        LB_TRY(capture_segment(s, fix_body, fix_tail, [&]() { return fused_fix_segment(s); }));
        LB_TRY(capture_segment(s, iter_body, tail, [&]() { return fused_direction_segment(s); }));
        LB_TRY(add_conditional(iter_body, tail, h_inner, cudaGraphCondTypeWhile, &trial_body));
        std::vector<cudaGraphNode_t> inner_tail;
        LB_TRY(capture_segment(s, trial_body, inner_tail, [&]() { return fused_trial_segment(s); }));
        LB_TRY(capture_segment(s, iter_body, tail, [&]() { return fused_accept_segment(s, 0); }));
        s->graph_fixed_launches = 4; // combine_trial + OP_F_DIR + accept_gram + OP_F_ACCEPT ; + 2 per further trial
    } else {
        LB_TRY(add_conditional(s->graph, top_tail, h_outer, cudaGraphCondTypeWhile, &iter_body));
        s->k_host = s->params.m; // capture the passes of all m window positions; unused ones exit at once
        LB_TRY(capture_segment(s, iter_body, tail, [&]() { return launch_direction(s); }));
        LB_TRY(add_conditional(iter_body, tail, h_inner, cudaGraphCondTypeWhile, &trial_body));
        std::vector<cudaGraphNode_t> inner_tail;
        LB_TRY(capture_segment(s, trial_body, inner_tail, [&]() {
            if (s->cb) return callback_trial_segment(s); // the user's kernels become nodes of the loop body
            s->trial_kernel<<<s->grid, kThreads, 0, s->stream>>>(s->d_st);
            return scalar_step(s, OP_LS_STEP, 0, PACK_NONE);
        }));
        LB_TRY(capture_segment(s, iter_body, tail, [&]() {
            if (s->cb) return callback_accept_segment(s, 0);
            s->accept_kernel<<<s->grid_accept, kThreads, 0, s->stream>>>(s->d_st, 0);
            s->launches += 1;
            return scalar_step(s, OP_ACCEPT, 0, PACK_ACCEPT);
        }));
        // fixed part = everything captured except the two nodes of the trial loop body
        s->graph_fixed_launches = (s->launches - l_saved) - 1; // scalar_step of the trial body counted once; k_trial not counted
    }
    GRAPH_TRY(cudaGraphInstantiate(&s->graph_exec, s->graph, 0));
    return 0;
}

static bool wants_graph(const lbfgsb200_solver *s)
{
    // graph mode: not instrumented, and on several GPUs only with the peer-to-peer exchange, whose kernels are
    // ordinary graph nodes (NCCL calls and event pairs stay on the stepped path).  A user objective runs in the graph
    // when its callback could be stream-captured (kernel launches / async copies on the stream it is given).
    return s->params.use_graph && !s->profiling && !s->cb_graph_failed && !(is_multi(s) && !s->comm->p2p);
}

// Can the user's callback be recorded?  One evaluation is captured into a throw-away graph: the capture must
// survive, no CUDA call of the callback may have failed (synchronising, cudaMalloc, the legacy stream: all illegal
// while capturing), and it may only have produced what a conditional-node body accepts -- kernel, memcpy, memset and empty nodes (no
// stream-ordered allocations, host functions or event nodes).
static bool callback_is_capturable(lbfgsb200_solver *s)
{
    cudaStream_t run_stream = s->stream;
    if (cudaStreamBeginCapture(s->capture_stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    s->stream = s->capture_stream;
    cudaGetLastError();
    const int rc = callback_eval(s, &s->d_st->ls.alpha, s->partials);
    // a CUDA call of the callback that is illegal during capture (a synchronisation, say) may fail without
    // invalidating the capture: the error it left behind is the tell
    const cudaError_t pending = cudaGetLastError();
    s->stream = run_stream;
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(s->capture_stream, &g);
    bool ok = rc == 0 && pending == cudaSuccess && e == cudaSuccess && g != nullptr;
    if (ok) {
        size_t count = 0;
        ok = cudaGraphGetNodes(g, nullptr, &count) == cudaSuccess && count > 0;
        std::vector<cudaGraphNode_t> nodes(count);
        if (ok) ok = cudaGraphGetNodes(g, nodes.data(), &count) == cudaSuccess;
        for (size_t i = 0; ok && i < count; ++i) {
            cudaGraphNodeType type;
            ok = cudaGraphNodeGetType(nodes[i], &type) == cudaSuccess &&
                 (type == cudaGraphNodeTypeKernel || type == cudaGraphNodeTypeMemcpy || type == cudaGraphNodeTypeMemset ||
                  type == cudaGraphNodeTypeEmpty);
        }
    }
    if (g) cudaGraphDestroy(g);
    const cudaError_t left = cudaGetLastError();
    if (getenv("LBFGSB200_VERBOSE"))
        fprintf(stderr, "lbfgsb200: callback probe capture: rc=%d pending=%s end=%s graph=%p left=%s -> %s\n", rc, cudaGetErrorName(pending),
                cudaGetErrorName(e), (void *)g, cudaGetErrorName(left), ok ? "capturable" : "not capturable");
    return ok;
}

// Builds the graph if this run wants one.  A user callback that cannot be recorded makes the solver fall back, once
// and for good, to the host-stepped loop, which has no such restriction.
static int ensure_graph(lbfgsb200_solver *s)
{
    if (!wants_graph(s) || s->graph_exec) return 0;
    if (s->cb && !callback_is_capturable(s)) {
        s->cb_graph_failed = true;
        if (getenv("LBFGSB200_VERBOSE")) fprintf(stderr, "lbfgsb200: the objective callback cannot be stream-captured: host-stepped loop\n");
        return 0;
    }
    return build_graph(s);
}

static int run_graph(lbfgsb200_solver *s, int64_t iterations)
{
    if (iterations <= 0) return 0;
    long long budget = iterations;
    CUDA_TRY(cudaMemcpyAsync(&s->d_st->iters_left, &budget, sizeof budget, cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaGraphLaunch(s->graph_exec, s->stream));
    return 0;
}

static int do_iterate(lbfgsb200_solver *s, int64_t iterations)
{
    NvtxRange nvtx_range("lbfgsb200:iterate");
    if (!s->x0_set) {
        set_error("iterate: set_x0 has not been called");
        return LBFGSB200_ERR_INVALID;
    }
    s->streams_at_start = s->h_snapshot.vec_streams;
    LB_TRY(ensure_graph(s));
    const bool graph = wants_graph(s);
    s->last_graph = graph;
    if (s->graph_exec) { // cudaGraphSetConditional is only legal inside the graph: gate it per run
        s->cond_flag = graph ? 1 : 0;
        CUDA_TRY(cudaMemcpyAsync(&s->d_st->cond_outer, s->cond_handles, sizeof s->cond_handles, cudaMemcpyHostToDevice, s->stream));
        CUDA_TRY(cudaMemcpyAsync(&s->d_st->use_graph, &s->cond_flag, sizeof(int), cudaMemcpyHostToDevice, s->stream));
    }
    const long long k0 = s->h_snapshot.k, t0 = s->h_snapshot.trial_evals;
    const bool was_done = s->h_snapshot.ctrl.done != 0;
    CUDA_TRY(cudaEventRecord(s->ev0, s->stream));
    int rc = graph ? run_graph(s, iterations)
                   : (s->cb ? run_stepped_callback(s, iterations) : (s->fused ? run_stepped_fused(s, iterations) : run_stepped(s, iterations)));
    CUDA_TRY(cudaEventRecord(s->ev1, s->stream));
    if (rc < 0) return rc;
    LB_TRY(snapshot(s));
    if (graph && iterations > 0) { // kernel nodes executed: fixed part per iteration + 2 per stand-alone trial
        const long long its = s->h_snapshot.k - k0, trials = s->h_snapshot.trial_evals - t0;
        if (s->fused) {
            // prologue + per iteration {combine_trial, OP_F_DIR, accept_gram, OP_F_ACCEPT} + 2 per further trial;
            // an iteration whose search failed ran its direction segment but no accept
            const bool ls_failed_now = !was_done && s->h_snapshot.status == LBFGSB200_LS_FAILED;
            s->launches += 1 + its * s->graph_fixed_launches + 2 * (trials - its - (ls_failed_now ? 1 : 0)) + (ls_failed_now ? 2 : 0);
        } else {
            // an exit at the top of / inside an iteration (converged, line search failed) still ran that
            // iteration's nodes as no-ops; "maximum iterations" is raised by the last accept itself
            const bool extra = s->h_snapshot.ctrl.done && s->h_snapshot.status != LBFGSB200_MAX_ITER;
            s->launches += (its + (extra ? 1 : 0)) * s->graph_fixed_launches + (s->cb ? 1 : 2) * trials; // (the user's own kernels are not counted)
        }
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
    // the fast path is the default: compact (Gram) direction in the fused two-kernel flow, whole solve as one
    // CUDA graph.  Parity with the reference is pinned for it at every size the reference is quoted on
    // (tests/test_gpu_large.py); m > 50 needs direction = LBFGSB200_DIR_TWO_LOOP.
    p->direction = LBFGSB200_DIR_AUTO;
    p->c1 = 1e-4;                                         // seq/config.h:5, par/constants.h:5
    p->c2 = (flavor != LBFGSB200_FLAVOR_SEQ) ? 0.7 : 0.9; // par/constants.h:6 / seq/config.h:6
    p->step0 = 1.0;                                       // INITIAL_STEP_SIZE
    p->shrink = 0.5;                                      // BACKTRACKING_ALPHA
    p->backtracking_tol = (flavor == LBFGSB200_FLAVOR_PAR_INLINED) ? 1e-10 : 1e-8; // BACKTRACKING_TOL (par/L-BFGS-Backtracking.cu:155)
    p->wolfe_min = 1e-10;                                 // WOLFE_INTERP_MIN
    p->ls_max_trials = 20;
    p->use_graph = 1;
    p->verbose = 0;
    p->grid_ctas = 0;
    p->num_gpus = 0; // lbfgsb200_solve: as many visible GPUs as the problem size warrants (see lbfgsb200.h)
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

// Tensor maps over the arena (compact.cuh, ArenaMaps): a row-major [4 + 2 nslots][stride] FP64 tensor; one map per
// run length r (box = T columns x r rows) plus the 2-column halo box of the fused accept kernel.
typedef CUresult (*encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int build_arena_maps(lbfgsb200_solver *s, ArenaMaps *dev_dst, int box_T, int staging)
{
    static encode_tiled_fn encode = nullptr; // the entry point is process-wide
    if (!encode) {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
            qres != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled is not available");
            cudaGetLastError();
            return LBFGSB200_ERR_CUDA;
        }
        encode = (encode_tiled_fn)fn;
    }
    const int rows_total = 4 + 2 * s->nslots;
    ArenaMaps *host = &s->arena_maps_host[staging];
    memset(host, 0, sizeof *host);
    auto make = [&](CUtensorMap *out, int box_cols, int box_rows) -> bool {
        const cuuint64_t gdim[2] = {(cuuint64_t)s->stride, (cuuint64_t)rows_total};
        const cuuint64_t gstride[1] = {(cuuint64_t)s->stride * sizeof(double)};
        const cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, s->arena, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) set_error("cuTensorMapEncodeTiled failed with %d (rows %d, box %d x %d)", (int)r, rows_total, box_cols, box_rows);
        return r == CUDA_SUCCESS;
    };
    bool ok = true;
    for (int r = 1; r <= (s->nslots > 3 ? s->nslots : 3) && ok; ++r) ok = make(&host->run[r], box_T, r);
    ok = ok && make(&host->halo, 2, 4) && make(&host->halo1, 2, 1);
    if (!ok) return LBFGSB200_ERR_CUDA;
    // (the host copy lives in the solver: the upload is stream-ordered)
    CUDA_TRY(cudaMemcpyAsync(dev_dst, host, sizeof(ArenaMaps), cudaMemcpyHostToDevice, s->stream));
    return 0;
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
        p->direction > 2) {
        set_error("bad flavor/profile/direction");
        return LBFGSB200_ERR_INVALID;
    }
    if (p->direction == LBFGSB200_DIR_COMPACT && p->m > kMaxCompactM) { set_error("compact direction supports m <= %d", kMaxCompactM); return LBFGSB200_ERR_INVALID; }
    if (p->max_iterations < 0 || p->ls_max_trials < 1) { set_error("bad iteration limits"); return LBFGSB200_ERR_INVALID; }
    // The trial loops run on the device (in graph mode the host cannot interrupt them): refuse constants with which
    // a search would never end (the backtracking loops stop when alpha * shrink^k drops below backtracking_tol).
    if (!(p->shrink > 0.0 && p->shrink < 1.0)) { set_error("shrink=%g must lie in (0, 1)", p->shrink); return LBFGSB200_ERR_INVALID; }
    if (!(p->step0 > 0.0) || !isfinite(p->step0)) { set_error("step0=%g must be positive and finite", p->step0); return LBFGSB200_ERR_INVALID; }
    if (!(p->backtracking_tol > 0.0) || !(p->wolfe_min > 0.0)) { set_error("backtracking_tol and wolfe_min must be positive"); return LBFGSB200_ERR_INVALID; }
    if (!(p->c1 > 0.0) || !(p->c2 > 0.0) || !isfinite(p->c1) || !isfinite(p->c2)) { set_error("c1 and c2 must be positive and finite"); return LBFGSB200_ERR_INVALID; }
    if (!(p->tolerance >= 0.0)) { set_error("tolerance=%g must be >= 0", p->tolerance); return LBFGSB200_ERR_INVALID; }
    if (log(p->backtracking_tol / p->step0) / log(p->shrink) > 1e5) { set_error("shrink=%g needs more than 1e5 backtracking trials to reach backtracking_tol", p->shrink); return LBFGSB200_ERR_INVALID; }
    if (p->num_gpus < 0 || p->num_gpus > kMaxRanks) { set_error("num_gpus=%d out of range 0..%d", p->num_gpus, kMaxRanks); return LBFGSB200_ERR_INVALID; }
    return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// one pinned 64-byte control block per solver, recycled through a process-wide free list (cudaHostAlloc /
// cudaFreeHost cost ~0.5 ms each and synchronise the device)
static std::vector<void *> g_pinned_free;
static void *pinned_block_get()
{
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (!g_pinned_free.empty()) {
            void *p = g_pinned_free.back();
            g_pinned_free.pop_back();
            return p;
        }
    }
    void *p = nullptr;
    if (cudaHostAlloc(&p, 64, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
static void pinned_block_put(void *p)
{
    std::lock_guard<std::mutex> lock(g_pool_mutex);
    g_pinned_free.push_back(p);
}

int lbfgsb200_create(lbfgsb200_solver_t **out, int objective, size_t n_global,
                     const lbfgsb200_params_t *params, lbfgsb200_comm_t *comm, size_t trace_rows)
{
    if (!out) return LBFGSB200_ERR_INVALID;
    *out = nullptr;
    NvtxRange nvtx_range("lbfgsb200:create");
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
    if (s->params.direction == LBFGSB200_DIR_AUTO)
        s->params.direction = params->m <= kMaxCompactM ? LBFGSB200_DIR_COMPACT : LBFGSB200_DIR_TWO_LOOP;
    params = &s->params;
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
    cudaGetDevice(&s->device);
    s->grid = pick_grid((long long)s->n_local, s->sms, params->grid_ctas);
    s->grid_accept = pick_grid((long long)s->n_local, s->sms, params->grid_ctas, kCtasPerSmAccept);
    const bool compact = params->direction == LBFGSB200_DIR_COMPACT;
    const int J = 2 * params->m + 1;
    size_t budget_bytes = (size_t)216 * 1024; // shared memory of the TMA kernels (LBFGSB200_GRAM_TMA_KB)
    if (compact) {
        // Pass A, stand-alone (user objectives; fused flow: only after a rejected pair), LBFGSB200_GRAM_TMA forces one:
        //   2  tensor-map TMA (UTMALDG.2D): <= 5 tiled loads per tile of the whole history (default)
        //   0  cp.async (LDGSTS) pipeline with column groups: the fallback when the tensor maps cannot be used
        //      (shards of 2^31 elements or more: TMA coordinates are 32-bit)
        // The fused flow (k_accept_gram / k_combine_trial, LBFGSB200_FUSED=0 disables it) needs the tensor maps.
        const char *env = getenv("LBFGSB200_GRAM_TMA");
        s->gram_tma = env ? (atoi(env) ? 2 : 0) : 2;
        if (s->stride >= ((size_t)1 << 31)) s->gram_tma = 0;
        const char *ef = getenv("LBFGSB200_FUSED");
        s->fused = s->gram_tma == 2 && objective != LBFGSB200_OBJ_DEVICE_CALLBACK && !(ef && atoi(ef) == 0);
        int Jt = J;
        if (s->gram_tma) {
            // ONE CTA per SM owning (almost) all of shared memory: NS stages of T-wide tiles of every row, so NS-1
            // whole tiles per SM are in flight.  Wide tiles amortise the per-tile costs (barriers, TMA issue, the
            // overlap of k_combine_trial), deep rings hide latency: the widest tile that still leaves 3 stages is
            // taken, then 2 stages of 256 (a TMA box dimension is at most 256 elements).  LBFGSB200_AG_TILE /
            // LBFGSB200_CT_TILE = "T,NS" force a configuration (tuning knobs).
            const char *eb = getenv("LBFGSB200_GRAM_TMA_KB");
            const size_t budget = (size_t)(eb ? atoi(eb) : 216) * 1024;
            budget_bytes = budget;
            auto pick_tile = [&](size_t rows, size_t budget, const char *envname, int *T_out, int *NS_out) {
                // measured on B200 (benchmarks/tile_sweep.sh, n = 1e8): 2 stages of 256 beat 3 of 192 beat 4 of 128
                // (m = 20: 86.6 / 83.3 / 79.8 it/s), so the widest tile with >= 2 stages wins
                static const int widths[] = {256, 192, 128, 64};
                int T = 64, NS = 2;
                for (int w : widths) {
                    const int ns = (int)(budget / (rows * (size_t)w * sizeof(double)));
                    if (ns >= 2) { T = w; NS = ns > kMaxStages ? kMaxStages : ns; break; }
                }
                if (const char *e = getenv(envname)) {
                    int t = 0, n2 = 0;
                    if (sscanf(e, "%d,%d", &t, &n2) == 2 && t >= 64 && t <= 256 && t % 64 == 0 && n2 >= 2 && n2 <= kMaxStages &&
                        rows * (size_t)t * sizeof(double) * (size_t)n2 <= budget + 8192) { T = t; NS = n2; }
                }
                *T_out = T;
                *NS_out = NS;
            };
            pick_tile((size_t)J, budget, "LBFGSB200_AG_TILE", &s->gram_T, &s->gram_NS);
            // k_accept_gram box layout (accept_gram.cuh).  Measured on B200 at n = 1e8 (boxes = 0 / 1 / 2 / 3):
            //   m = 5 : 230.0 / 234.6 / 235.8 / 242.4 it/s   (box-issue-bound: fewer boxes win)
            //   m = 10: 167.0 / 164.9 / 164.9 / 162.6 it/s   (DRAM-bound: independent one-row boxes win)
            s->ag_boxes = params->m <= 7 ? 3 : 0;
            if (const char *e = getenv("LBFGSB200_AG_BOXES")) s->ag_boxes = atoi(e) & 3;
            if (kAgGroups > 1 && s->gram_NS >= 2 * kAgGroups) s->gram_NS -= s->gram_NS % kAgGroups; // (accept_gram.cuh: groups must not share a stage)
            // one row more (x), and ~9 KB of the CTA's shared memory are static (the boundary exchange of its two groups)
            pick_tile((size_t)J + 1, budget - 8192, "LBFGSB200_CT_TILE", &s->ct_T, &s->ct_NS);
            if (const char *e = getenv("LBFGSB200_CT_HALO")) s->ct_halo = atoi(e) >= 1 && 2 * atoi(e) < s->ct_T / 4 ? atoi(e) : kCtHaloItems;
            s->ag_smem = (size_t)s->gram_NS * accept_gram_stage_doubles(J, s->gram_T) * sizeof(double);
            s->ct_smem = (size_t)s->ct_NS * combine_trial_stage_doubles(params->m, s->ct_T) * sizeof(double);
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
        // TMA variants: 16 consumer warps = NG column groups x NE element groups.  Each element group
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
        s->grid_ag = s->grid_gram;
        if (s->fused) {
            const int cw = accept_gram_cw(params->m, s->gram_T);
            // The fused kernels pay per-tile costs (the overlap of k_combine_trial, the accept warps' latency) that only
            // 256-wide tiles amortise, and the 12 gram warps of k_accept_gram hold at most kAgMaxCW columns each: the
            // fused flow runs for m <= 13 (J = 2m+1 <= 27 columns in 3 column groups).  Above that the unfused flow --
            // stand-alone pass A with 16 consumer warps + register-streaming k_combine -- is the faster one (measured at
            // n = 1e8: m = 20: 93.9 vs 86.6 it/s with 128-wide fused tiles; m = 50: 41.3 vs 26.9).
            // LBFGSB200_FUSED=1 forces the fused flow where it can run at all.
            const bool force = ef && atoi(ef) == 1;
            if (s->gram_T < 64 || cw > kAgMaxCW || (!force && (s->gram_T < 256 || s->ct_T < 256))) s->fused = false;
        }
        if (s->gram_tma && !s->fused) {
            // the stand-alone pass A keeps its own tile policy: 4 stages of the widest power-of-two tile that fits
            s->gram_NS = kGramStages;
            s->gram_T = 256;
            while (s->gram_T > 32 && (size_t)s->gram_NS * J * s->gram_T * sizeof(double) > budget_bytes) s->gram_T >>= 1;
            s->gram_smem = (size_t)s->gram_NS * J * s->gram_T * sizeof(double);
            int NE2 = s->gram_T / 2 / 32;
            if (NE2 < 1) NE2 = 1;
            if (NE2 > kWsConsumerWarps / 2) NE2 = kWsConsumerWarps / 2;
            s->gram_NG = kWsConsumerWarps / NE2;
            while (s->gram_NG < kWsConsumerWarps && (J + s->gram_NG - 1) / s->gram_NG > kMaxCW) s->gram_NG <<= 1;
            const long long tiles2 = ((long long)s->n_local + s->gram_T - 1) / s->gram_T;
            s->grid_gram = params->grid_ctas > 0 ? params->grid_ctas : (int)(tiles2 < s->sms ? (tiles2 < 1 ? 1 : tiles2) : s->sms);
        }
        if (s->fused) {
            {
                const int cw = accept_gram_cw(params->m, s->gram_T);
                s->ag_kernel = cw <= 7 ? accept_gram_kernel_for<7>(objective) : accept_gram_kernel_for<kAgMaxCW>(objective);
                s->ct_kernel = combine_trial_kernel_for(objective);
                // k_combine_trial: one CTA per SM over tiles that own T/2 - 2 halo double2 items each
                const long long own2 = s->ct_T / 2 - 2 * s->ct_halo;
                const long long ct_tiles = (((long long)s->n_local + 1) / 2 + own2 - 1) / own2;
                s->grid_combine = params->grid_ctas > 0 ? params->grid_ctas : (int)(ct_tiles < s->sms ? ct_tiles : s->sms);
            }
        }
    }

#define CREATE_TRY(expr)                                                                      \
    do {                                                                                      \
        cudaError_t e2_ = (expr);                                                             \
        if (e2_ != cudaSuccess) {                                                             \
            set_error("%s failed: %s", #expr, cudaGetErrorString(e2_));                       \
            lbfgsb200_destroy(s);                                                             \
            return LBFGSB200_ERR_CUDA;                                                        \
        }                                                                                     \
    } while (0)
#define CREATE_RC(expr)                                                                       \
    do {                                                                                      \
        int rc2_ = (expr);                                                                    \
        if (rc2_ < 0) {                                                                       \
            lbfgsb200_destroy(s);                                                             \
            return rc2_;                                                                      \
        }                                                                                     \
    } while (0)
    CREATE_TRY(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaStreamCreateWithFlags(&s->capture_stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev0, cudaEventDefault));
    CREATE_TRY(cudaEventCreateWithFlags(&s->ev1, cudaEventDefault));
    // The arena ((2m+6) vectors, tens of GB) and the small buffers come from the library's PRIVATE stream-ordered
    // pool (pool_alloc above): a destroyed solver's memory is re-used by the next create without a driver round trip.
    const size_t nvecs = 4 + 2 * (size_t)s->nslots;
    const size_t arena_bytes = nvecs * s->stride * sizeof(double);
    {
        int rc_a = pool_alloc((void **)&s->arena, arena_bytes, s->stream);
        if (rc_a < 0) {
            const std::string why = lbfgsb200_last_error();
            set_error("allocation of %.2f GB for %zu vectors failed (%s)", arena_bytes / 1e9, nvecs, why.c_str());
            lbfgsb200_destroy(s);
            return LBFGSB200_ERR_NOMEM;
        }
    }
    // Zero only what must be zero: the work vector w (d = 0 for the x0 evaluation) and the <= 31
    // pad doubles at the end of every row (the compact kernels read rows up to the padded length).
    // Everything else is written before it is read; clearing the whole arena would cost a full
    // HBM pass over (2m+6) vectors on every create().
    {
        const size_t pad = s->stride - s->n_local;
        CREATE_TRY(cudaMemsetAsync(s->arena + kArenaRowW * s->stride, 0, s->stride * sizeof(double), s->stream));
        if (pad) // one strided memset for the pad of every row
            CREATE_TRY(cudaMemset2DAsync(s->arena + s->n_local, s->stride * sizeof(double), 0, pad * sizeof(double), nvecs, s->stream));
    }
    // ---- ONE allocation for all small device buffers ----
    size_t npart = (size_t)kMaxQ * (size_t)s->grid;
    size_t gram_nb = 0, gram_cnt = 0;
    if (compact) {
        const size_t need = ((size_t)3 * J + 1) * (size_t)s->grid_gram;
        if (need > npart) npart = need;
        gram_nb = (size_t)(2 * s->nslots + 1);
        gram_cnt = (size_t)3 * J + kRowsExtra;
        s->gram_doubles = gram_nb * gram_nb + gram_cnt * (size_t)(nranks + 1) + (size_t)J;
    }
    s->trace_rows = trace_rows;
    size_t off = 0;
    const size_t off_st = off;        off = align_up(off + sizeof(DevState), 256);
    const size_t off_part = off;      off = align_up(off + sizeof(double) * npart, 256);
    const size_t off_pkt = off;       off = align_up(off + sizeof(double) * kPacket * (size_t)(nranks + 1), 256);
    const size_t off_gram = off;      off = align_up(off + sizeof(double) * s->gram_doubles, 256);
    const size_t off_maps = off;      off = align_up(off + (compact && s->gram_tma ? 2 * sizeof(ArenaMaps) : 0), 256);
    const size_t off_trace = off;     off = align_up(off + sizeof(double) * LBFGSB200_TRACE_COLS * trace_rows, 256);
    int tl_cap = 0;
    if (const char *et = getenv("LBFGSB200_TIMELINE")) tl_cap = atoi(et) > 0 ? atoi(et) : 0;
    const size_t off_tl = off;        off = align_up(off + sizeof(unsigned long long) * 3 * (size_t)tl_cap, 256);
    CREATE_RC(pool_alloc(&s->aux, off, s->stream));
    char *aux = (char *)s->aux;
    CREATE_TRY(cudaMemsetAsync(aux, 0, off_trace, s->stream)); // state, partials, packets, Gram block, maps
    if (trace_rows) CREATE_TRY(cudaMemsetAsync(aux + off_trace, 0, off_tl - off_trace, s->stream));
    s->d_st = (DevState *)(aux + off_st);
    s->partials = (double *)(aux + off_part);
    s->pkt = (double *)(aux + off_pkt);
    if (s->gram_doubles) s->gram = (double *)(aux + off_gram);
    if (compact && s->gram_tma) {
        s->arena_maps = (ArenaMaps *)(aux + off_maps);
        s->arena_maps_ct = s->arena_maps + 1;
    }
    if (trace_rows) s->trace = (double *)(aux + off_trace);
    if (tl_cap) s->timeline = (unsigned long long *)(aux + off_tl);
    if (compact) {
        // opt every pass-A variant in to the device's full dynamic shared memory (the limit is per function and
        // per device; occupancy still follows the size actually passed at launch).  Done once per device.
        static bool opted[kMaxDevices] = {};
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (!opted[s->device]) {
            int optin = 0;
            CREATE_TRY(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s->device));
            CREATE_TRY(cudaFuncSetAttribute((const void *)k_scalar, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)(sizeof(double) * kMaxCols * kMaxCols)));
            const void *variants[] = {(const void *)k_gram<3>, (const void *)k_gram<6>, (const void *)k_gram<kMaxCW>,
                                      (const void *)k_gram_tma2d<7>, (const void *)k_gram_tma2d<kMaxCW>,
                                      (const void *)k_accept_gram<ObjQuadratic, 7>, (const void *)k_accept_gram<ObjRosenbrock, 7>,
                                      (const void *)k_accept_gram<ObjTridiag, 7>, (const void *)k_accept_gram<ObjQuadratic, kAgMaxCW>,
                                      (const void *)k_accept_gram<ObjRosenbrock, kAgMaxCW>, (const void *)k_accept_gram<ObjTridiag, kAgMaxCW>,
                                      (const void *)k_combine_trial<ObjQuadratic>, (const void *)k_combine_trial<ObjRosenbrock>,
                                      (const void *)k_combine_trial<ObjTridiag>};
            for (const void *fn : variants) {
                cudaFuncAttributes fa;
                CREATE_TRY(cudaFuncGetAttributes(&fa, fn));
                CREATE_TRY(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
            }
            s_optin[s->device] = optin;
            opted[s->device] = true;
        }
        const size_t need = (s->fused ? (s->ag_smem + 1024 > s->ct_smem + 10240 ? s->ag_smem + 1024 : s->ct_smem + 10240) : s->gram_smem + 1024) + 1024;
        if ((size_t)s_optin[s->device] < need) {
            set_error("pass A needs %zu bytes of shared memory, the device offers %d", need, s_optin[s->device]);
            lbfgsb200_destroy(s);
            return LBFGSB200_ERR_INVALID;
        }
    }
    s->h_ctrl = (Ctrl *)pinned_block_get();
    if (!s->h_ctrl) { set_error("cudaHostAlloc of the control block failed"); lbfgsb200_destroy(s); return LBFGSB200_ERR_NOMEM; }
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
    st.arena0 = s->arena;
    st.x = s->arena;
    st.x_alt = s->arena + kArenaRowXb * s->stride;
    st.g = s->arena + kArenaRowG * s->stride;
    st.w = s->arena + kArenaRowW * s->stride;
    st.S = s->arena + kArenaRowS * s->stride;
    st.Y = st.S + (size_t)s->nslots * s->stride;
    st.stride = (long long)s->stride;
    st.fused = s->fused ? 1 : 0;
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
    st.tl_cap = tl_cap;
    st.tl_sub = getenv("LBFGSB200_TIMELINE_SUB") ? 1 : 0;
    st.tl = s->timeline;
    if (s->arena_maps) {
        CREATE_RC(build_arena_maps(s, s->arena_maps, s->gram_T, 0));
        CREATE_RC(build_arena_maps(s, s->arena_maps_ct, s->ct_T, 1));
    }
    CREATE_TRY(cudaMemcpyAsync(s->d_st, &st, sizeof st, cudaMemcpyHostToDevice, s->stream));
    // no synchronisation here: everything above is ordered on the solver stream, which every later call uses
#undef CREATE_TRY
#undef CREATE_RC
    *out = s;
    return 0;
}

void lbfgsb200_destroy(lbfgsb200_solver_t *s)
{
    if (!s) return;
    if (s->device >= 0) cudaSetDevice(s->device);
    if (s->stream) cudaStreamSynchronize(s->stream);
    for (int c = 0; c < LBFGSB200_PROFILE_CLASSES; ++c)
        for (cudaEvent_t e : s->prof_ev[c]) cudaEventDestroy(e);
    if (s->stream) { // back to the private pool (see pool_alloc): stream-ordered, no device-wide synchronisation
        if (s->arena) cudaFreeAsync(s->arena, s->stream);
        if (s->aux) cudaFreeAsync(s->aux, s->stream);
        if (s->cb_buf) cudaFreeAsync(s->cb_buf, s->stream);
        cudaStreamSynchronize(s->stream);
    }
    if (s->h_ctrl) pinned_block_put(s->h_ctrl);
    if (s->graph_exec) cudaGraphExecDestroy(s->graph_exec);
    if (s->graph) cudaGraphDestroy(s->graph);
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    if (s->capture_stream) cudaStreamDestroy(s->capture_stream);
    if (s->stream) cudaStreamDestroy(s->stream);
    delete s;
}

int lbfgsb200_create_callback_sharded(lbfgsb200_solver_t **out, lbfgsb200_fg_device_fn fn, void *user, size_t n_global,
                                      const lbfgsb200_params_t *params, lbfgsb200_comm_t *comm, size_t trace_rows)
{
    if (!fn) { set_error("create_callback: fn is NULL"); return LBFGSB200_ERR_INVALID; }
    int rc = lbfgsb200_create(out, LBFGSB200_OBJ_DEVICE_CALLBACK, n_global, params, comm, trace_rows);
    if (rc < 0) return rc;
    lbfgsb200_solver *s = *out;
    s->cb = fn;
    s->cb_user = user;
    rc = pool_alloc((void **)&s->cb_buf, sizeof(double) * (s->stride + 8), s->stream);
    cudaError_t e = rc < 0 ? cudaErrorMemoryAllocation : cudaMemsetAsync(s->cb_buf, 0, sizeof(double) * (s->stride + 8), s->stream);
    if (e != cudaSuccess) {
        if (rc >= 0) set_error("create_callback: %s", cudaGetErrorString(e));
        cudaGetLastError();
        lbfgsb200_destroy(s);
        *out = nullptr;
        return LBFGSB200_ERR_NOMEM;
    }
    return 0;
}

int lbfgsb200_create_callback(lbfgsb200_solver_t **out, lbfgsb200_fg_device_fn fn, void *user, size_t n,
                              const lbfgsb200_params_t *params, size_t trace_rows)
{
    return lbfgsb200_create_callback_sharded(out, fn, user, n, params, nullptr, trace_rows);
}

const double *lbfgsb200_device_halo(const lbfgsb200_solver_t *s)
{
    // DevState keeps xL, xR, dL, dR, gL, gR as six consecutive doubles (state.h)
    static_assert(offsetof(DevState, gR) - offsetof(DevState, xL) == 5 * sizeof(double), "halo block layout");
    return s ? &s->d_st->xL : nullptr;
}

size_t lbfgsb200_local_size(const lbfgsb200_solver_t *s) { return s ? s->n_local : 0; }

// ---- checkpoint / resume ------------------------------------------------------------------------
struct CkptHeader {
    char magic[8]; // "LBB200C1"
    uint64_t n_local, stride, nvecs, gram_doubles, trace_rows, devstate_bytes;
    int32_t m, objective, direction, profile, rank, nranks, x_is_alt, pad;
    int64_t k_host;
};

static size_t gram_doubles_of(const lbfgsb200_solver *s) { return s->gram_doubles; }

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
                h.gram_doubles != gram_doubles_of(s) || h.trace_rows != s->trace_rows || st.fused != s->h_snapshot.fused ||
                (s->cb && h.x_is_alt))) { // (user objectives keep x in place)
        set_error("checkpoint_load: the checkpoint was written by a solver of a different shape");
        rc = LBFGSB200_ERR_INVALID;
    }
    if (!rc) { // the whole payload must be there BEFORE the first byte of this solver's state is replaced
        const long pos = ftell(f);
        fseek(f, 0, SEEK_END);
        const long end = ftell(f);
        fseek(f, pos, SEEK_SET);
        const size_t payload = sizeof(double) * ((size_t)h.nvecs * h.stride + h.gram_doubles + (size_t)s->trace_rows * LBFGSB200_TRACE_COLS);
        if (pos < 0 || end < 0 || (size_t)(end - pos) != payload) {
            set_error("checkpoint_load: %s is truncated (%ld payload bytes, expected %zu)", path, end - pos, payload);
            rc = LBFGSB200_ERR_INVALID;
        }
    }
    if (rc) { fclose(f); return rc; }
    s->x0_set = false; // from here on a failure leaves the arena half-replaced: the handle needs set_x0 or a good load
    if (!rc) rc = stream_file_to_dev(f, s->arena, h.nvecs * h.stride, s->stream);
    if (!rc && h.gram_doubles) rc = stream_file_to_dev(f, s->gram, h.gram_doubles, s->stream);
    if (!rc && s->trace_rows) rc = stream_file_to_dev(f, s->trace, s->trace_rows * LBFGSB200_TRACE_COLS, s->stream);
    fclose(f);
    if (rc) return rc;
    // scalars come from the file; every pointer and handle from this solver
    const DevState &cur = s->h_snapshot;
    st.arena0 = s->arena;
    st.x = h.x_is_alt ? s->arena + kArenaRowXb * s->stride : s->arena;
    st.x_alt = s->cb ? st.x : (h.x_is_alt ? s->arena : s->arena + kArenaRowXb * s->stride);
    st.g = cur.g; st.w = cur.w; st.S = cur.S; st.Y = cur.Y;
    st.partials = cur.partials; st.send = cur.send; st.recv = cur.recv; st.trace = cur.trace;
    st.gram = cur.gram; st.gram_rows = cur.gram_rows; st.gram_recv = cur.gram_recv; st.delta = cur.delta;
    st.mail = cur.mail; st.peers = cur.peers; st.p2p = cur.p2p; st.p2p_timeout_ns = cur.p2p_timeout_ns;
    st.cond_outer = cur.cond_outer; st.cond_inner = cur.cond_inner; st.cond_fix = cur.cond_fix; st.use_graph = cur.use_graph;
    st.tl = cur.tl; st.tl_cap = cur.tl_cap; st.tl_n = cur.tl_n; st.tl_sub = cur.tl_sub;
    st.max_iterations = cur.max_iterations; st.tolerance = cur.tolerance; st.lsp = cur.lsp; // the new handle's limits apply
    if (st.status != LBFGSB200_CONVERGED && st.status != LBFGSB200_LS_FAILED && st.k < st.max_iterations) {
        st.status = LBFGSB200_RUNNING; // a run that only ran out of iterations may continue
        st.ctrl.done = 0;
    }
    s->h_snapshot = st;
    *s->h_ctrl = st.ctrl;
    CUDA_TRY(cudaMemcpyAsync(s->d_st, &s->h_snapshot, sizeof(DevState), cudaMemcpyHostToDevice, s->stream));
    CUDA_TRY(cudaStreamSynchronize(s->stream));
    s->k_host = h.k_host;
    s->x0_set = true;
    return 0;
}

// ---- host <-> device copies of whole vectors -------------------------------------------------------
// The reference's callers hand over pageable std::vector storage.  A plain cudaMemcpy from pageable memory is staged
// by the driver through one thread (~10 GB/s); here kCopyThreads host threads stage 4 MB chunks through pinned buffers
// of their own, the memcpy of one chunk overlapping the DMA of the previous ones, which gets a pageable buffer close
// to the pinned PCIe rate.  Pinned and device pointers take the direct route.
constexpr size_t kStageChunk = (size_t)4 << 20;
constexpr int kCopyThreads = 4, kStageBufs = 2 * kCopyThreads;
static char *g_stage[kMaxDevices] = {}; // kStageBufs pinned chunks per device, allocated on first use, kept
static std::mutex g_stage_mutex;        // one staged transfer per process at a time (the buffers are shared)

static bool is_pageable(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

static int staged_copy(lbfgsb200_solver *s, double *dst, const double *src, size_t bytes, bool to_device)
{
    std::lock_guard<std::mutex> lock(g_stage_mutex);
    const int dev = s->device;
    if (!g_stage[dev]) CUDA_TRY(cudaHostAlloc((void **)&g_stage[dev], kStageChunk * kStageBufs, cudaHostAllocPortable));
    const size_t nchunks = (bytes + kStageChunk - 1) / kStageChunk;
    int rcs[kCopyThreads] = {};
    auto worker = [&](int t) {
        cudaSetDevice(dev);
        cudaEvent_t ev[2] = {nullptr, nullptr};
        bool busy[2] = {false, false};
        size_t pending[2] = {0, 0};
        if (cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) != cudaSuccess) { rcs[t] = 1; return; }
        char *buf[2] = {g_stage[dev] + (size_t)(2 * t) * kStageChunk, g_stage[dev] + (size_t)(2 * t + 1) * kStageChunk};
        auto span = [&](size_t c) { return (c + 1) * kStageChunk <= bytes ? kStageChunk : bytes - c * kStageChunk; };
        int slot = 0;
        for (size_t c = t; c < nchunks && !rcs[t]; c += kCopyThreads, slot ^= 1) {
            const size_t off = c * kStageChunk, len = span(c);
            if (busy[slot]) { // the previous transfer through this buffer must have left it
                if (cudaEventSynchronize(ev[slot]) != cudaSuccess) { rcs[t] = 1; break; }
                if (!to_device) memcpy((char *)dst + pending[slot] * kStageChunk, buf[slot], span(pending[slot]));
                busy[slot] = false;
            }
            if (to_device) {
                memcpy(buf[slot], (const char *)src + off, len);
                if (cudaMemcpyAsync((char *)dst + off, buf[slot], len, cudaMemcpyHostToDevice, s->stream) != cudaSuccess) rcs[t] = 1;
            } else {
                if (cudaMemcpyAsync(buf[slot], (const char *)src + off, len, cudaMemcpyDeviceToHost, s->stream) != cudaSuccess) rcs[t] = 1;
            }
            if (cudaEventRecord(ev[slot], s->stream) != cudaSuccess) rcs[t] = 1;
            busy[slot] = true;
            pending[slot] = c;
        }
        for (int k = 0; k < 2; ++k) {
            if (busy[slot]) {
                if (cudaEventSynchronize(ev[slot]) != cudaSuccess) rcs[t] = 1;
                else if (!to_device) memcpy((char *)dst + pending[slot] * kStageChunk, buf[slot], span(pending[slot]));
                busy[slot] = false;
            }
            slot ^= 1;
        }
        cudaEventDestroy(ev[0]);
        cudaEventDestroy(ev[1]);
    };
    std::thread th[kCopyThreads];
    const int nthreads = (int)(nchunks < (size_t)kCopyThreads ? nchunks : (size_t)kCopyThreads);
    for (int t = 1; t < nthreads; ++t) th[t] = std::thread(worker, t);
    worker(0);
    for (int t = 1; t < nthreads; ++t) th[t].join();
    for (int t = 0; t < nthreads; ++t)
        if (rcs[t]) { set_error("staged %s copy failed: %s", to_device ? "host-to-device" : "device-to-host", cudaGetErrorString(cudaGetLastError())); return LBFGSB200_ERR_CUDA; }
    return 0;
}

static int copy_vector(lbfgsb200_solver *s, double *dst, const double *src, size_t count, bool to_device)
{
    const size_t bytes = count * sizeof(double);
    const void *host_side = to_device ? (const void *)src : (const void *)dst;
    if (bytes >= 2 * kStageChunk && is_pageable(host_side)) return staged_copy(s, dst, src, bytes, to_device);
    CUDA_TRY(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, s->stream));
    return 0;
}

int lbfgsb200_set_x0(lbfgsb200_solver_t *s, const double *x0_local)
{
    if (!s || !x0_local) { set_error("set_x0: NULL argument"); return LBFGSB200_ERR_INVALID; }
    NvtxRange nvtx_range("lbfgsb200:set_x0");
    // restore the pristine state (ring empty, pointers un-swapped), then evaluate f(x0), g(x0)
    DevState &st = s->h_snapshot;
    st.x = s->arena;
    st.x_alt = s->cb ? s->arena : s->arena + kArenaRowXb * s->stride; // user objectives: x is updated in place
    st.base = 0;
    st.h = 0;
    st.k = 0;
    st.ctrl.ls_active = 0;
    st.ctrl.done = 0;
    st.ctrl.h = 0;
    st.ctrl.k = 0;
    st.ctrl.need_fix = 0;
    st.steepest = 0;
    st.pend_steepest = 0;
    st.status = LBFGSB200_RUNNING;
    st.use_graph = 0; // re-armed per run by do_iterate
    memset(&st.ls, 0, sizeof st.ls); // FLAVOR_PAR_INLINED carries state from one search to the next
    st.tl_n = 0;
    st.xL = st.xR = st.dL = st.dR = st.gL = st.gR = 0.0;
    CUDA_TRY(cudaMemcpyAsync(s->d_st, &st, sizeof st, cudaMemcpyHostToDevice, s->stream));
    LB_TRY(copy_vector(s, st.x, x0_local, s->n_local, true));
    CUDA_TRY(cudaMemsetAsync(st.w, 0, s->stride * sizeof(double), s->stream)); // d = 0
    // record the graph while the upload is in flight (recording executes nothing; it has its own stream)
    s->profiling = false;
    LB_TRY(ensure_graph(s));
    if (is_multi(s)) {
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
        LB_TRY(callback_accept_segment(s, 1));
    } else if (s->fused) {
        LB_TRY(fused_accept_segment(s, 1));
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
    NvtxRange nvtx_range("lbfgsb200:get_x");
    LB_TRY(copy_vector(s, x_local_out, s->h_snapshot.x, s->n_local, false));
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
    r->flow = s->params.direction == LBFGSB200_DIR_TWO_LOOP ? 0 : (s->fused ? 2 : 1);
    r->graph = s->last_graph ? 1 : 0;
    r->num_gpus = s->comm ? s->comm->nranks : 1;
    r->reserved = 0;
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

static void print_verbose(const lbfgsb200_result_t &r, const double *rows, int64_t got)
{
    // the reference's per-iteration line (seq/lbfgs.cpp:77-78: printed at the top of iteration k
    // with the current f and |grad|), reproduced from the device trace
    const int64_t lines = r.status == LBFGSB200_MAX_ITER ? r.iterations : r.iterations + 1;
    for (int64_t k = 0; k < lines; ++k) {
        if (k == 0) printf("Iteration 0, f = %g, |grad| = %g\n", r.f0, r.gnorm0);
        else if (k - 1 < got)
            printf("Iteration %lld, f = %g, |grad| = %g\n", (long long)k, rows[(k - 1) * LBFGSB200_TRACE_COLS + 1],
                   rows[(k - 1) * LBFGSB200_TRACE_COLS + 2]);
    }
}

static int solve_single(int objective, size_t n, const double *x0_host, double *x_out_host,
                        const lbfgsb200_params_t *params, lbfgsb200_result_t *result, double *trace, size_t trace_rows)
{
    lbfgsb200_solver *s = nullptr;
    const size_t rows_wanted = trace ? trace_rows : (params->verbose ? (size_t)(params->max_iterations < 100000 ? params->max_iterations : 100000) : 0);
    int rc = lbfgsb200_create(&s, objective, n, params, nullptr, rows_wanted);
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
        lbfgsb200_result_t r;
        lbfgsb200_get_result(s, &r);
        std::vector<double> rows((size_t)LBFGSB200_TRACE_COLS * (rows_wanted ? rows_wanted : 1));
        const int64_t got = rows_wanted ? lbfgsb200_get_trace(s, rows.data(), rows_wanted) : 0;
        print_verbose(r, rows.data(), got);
    }
    lbfgsb200_destroy(s);
    return rc;
}

// ---- one process, several GPUs ------------------------------------------------------------------
// The reference is ONE function called from ONE host thread (seq/benchmark.cpp:94, par/L-BFGS-Wolfe.cu:473).  Behind
// that call the library drives P devices itself: one worker thread per GPU, each owning a contiguous shard and
// running the ordinary per-rank solver; the ranks meet in the NVLink mailboxes (peer access enabled in-process, no
// NCCL, no IPC).  The caller's x0 is scattered straight from its buffer (every GPU pulls its shard over its own
// PCIe link, concurrently) and x is gathered the same way.
int lbfgsb200_resolve_num_gpus(int requested, size_t n)
{
    const int visible = lbfgsb200_device_count();
    if (visible < 1) return 0;
    if (const char *e = getenv("LBFGSB200_NUM_GPUS")) {
        if (requested == 0 && atoi(e) > 0) requested = atoi(e);
    }
    if (requested > 0) return requested;
    // automatic: a GPU is worth adding while its shard keeps >= 2^23 elements (64 MB per vector); below that the
    // iteration is bound by the per-step exchange latency, not by bandwidth
    size_t want = n >> 23;
    if (want < 1) want = 1;
    int p = (int)(want < (size_t)visible ? want : (size_t)visible);
    if (p > kMaxRanks) p = kMaxRanks;
    return p;
}

static int solve_group(int P, int objective, size_t n, const double *x0_host, double *x_out_host,
                       const lbfgsb200_params_t *params, lbfgsb200_result_t *result, double *trace, size_t trace_rows)
{
    if (P > lbfgsb200_device_count()) { set_error("num_gpus=%d but only %d CUDA devices are visible", P, lbfgsb200_device_count()); return LBFGSB200_ERR_INVALID; }
    if (n / (size_t)P < 2) { set_error("n=%zu is too small for %d GPUs", n, P); return LBFGSB200_ERR_INVALID; }
    int caller_dev = 0;
    cudaGetDevice(&caller_dev);
    std::vector<int> devices(P);
    for (int r = 0; r < P; ++r) devices[r] = r;
    std::vector<lbfgsb200_comm_t *> comms(P, nullptr);
    int rc = lbfgsb200_comm_create_local(comms.data(), devices.data(), P);
    if (rc < 0) { cudaSetDevice(caller_dev); return rc; }
    lbfgsb200_params_t prm = *params;
    prm.use_graph = 1; // the ranks advance in lock step on the device; the host threads only launch and wait
    const size_t rows_wanted = trace ? trace_rows : (params->verbose ? (size_t)(params->max_iterations < 100000 ? params->max_iterations : 100000) : 0);
    std::vector<lbfgsb200_solver *> solvers(P, nullptr);
    std::vector<int> rcs(P, 0);
    std::vector<std::string> errs(P);
    auto on_all = [&](auto &&fn) {
        std::vector<std::thread> th;
        for (int r = 0; r < P; ++r)
            th.emplace_back([&, r]() {
                cudaSetDevice(devices[r]);
                rcs[r] = fn(r);
                if (rcs[r] < 0) errs[r] = lbfgsb200_last_error();
            });
        for (auto &t : th) t.join();
        for (int r = 0; r < P; ++r)
            if (rcs[r] < 0) { set_error("GPU %d: %s", devices[r], errs[r].c_str()); return rcs[r]; }
        return 0;
    };
    // phase 1 (no exchange yet): every rank must have its solver before any rank waits for a peer
    rc = on_all([&](int r) { return lbfgsb200_create(&solvers[r], objective, n, &prm, comms[r], r == 0 ? rows_wanted : 0); });
    int status = rc;
    if (rc >= 0) {
        rc = on_all([&](int r) {
            lbfgsb200_solver *s = solvers[r];
            int e = lbfgsb200_set_x0(s, x0_host + s->offset);
            if (e < 0) return e;
            e = lbfgsb200_iterate(s, (int64_t)prm.max_iterations + 1);
            if (e < 0) return e;
            const int e2 = lbfgsb200_get_x(s, x_out_host + s->offset);
            return e2 < 0 ? e2 : e;
        });
        status = rc < 0 ? rc : rcs[0];
    }
    if (rc >= 0) {
        cudaSetDevice(devices[0]);
        lbfgsb200_result_t r0;
        lbfgsb200_get_result(solvers[0], &r0);
        for (int r = 1; r < P; ++r) { // launches are per rank; report the sum, bytes likewise
            cudaSetDevice(devices[r]);
            lbfgsb200_result_t rr;
            lbfgsb200_get_result(solvers[r], &rr);
            r0.kernel_launches += rr.kernel_launches;
            r0.bytes_moved += rr.bytes_moved;
            if (rr.device_ms > r0.device_ms) r0.device_ms = rr.device_ms;
        }
        cudaSetDevice(devices[0]);
        if (result) *result = r0;
        if (trace) lbfgsb200_get_trace(solvers[0], trace, trace_rows);
        if (params->verbose) {
            std::vector<double> rows((size_t)LBFGSB200_TRACE_COLS * (rows_wanted ? rows_wanted : 1));
            const int64_t got = rows_wanted ? lbfgsb200_get_trace(solvers[0], rows.data(), rows_wanted) : 0;
            print_verbose(r0, rows.data(), got);
        }
    }
    for (int r = 0; r < P; ++r)
        if (solvers[r]) {
            cudaSetDevice(devices[r]);
            lbfgsb200_destroy(solvers[r]);
        }
    for (int r = 0; r < P; ++r)
        if (comms[r]) {
            cudaSetDevice(devices[r]);
            lbfgsb200_comm_destroy(comms[r]);
        }
    cudaSetDevice(caller_dev);
    return status;
}

int lbfgsb200_solve(int objective, size_t n, const double *x0_host, double *x_out_host,
                    const lbfgsb200_params_t *params, lbfgsb200_result_t *result, double *trace,
                    size_t trace_rows)
{
    if (!x0_host || !x_out_host || !params) { set_error("solve: NULL argument"); return LBFGSB200_ERR_INVALID; }
    LB_TRY(check_params(params));
    if (lbfgsb200_device_count() < 1) {
        set_error("no usable CUDA device: this library has no CPU fallback");
        return LBFGSB200_ERR_CUDA;
    }
    const int P = lbfgsb200_resolve_num_gpus(params->num_gpus, n);
    if (P > 1) return solve_group(P, objective, n, x0_host, x_out_host, params, result, trace, trace_rows);
    return solve_single(objective, n, x0_host, x_out_host, params, result, trace, trace_rows);
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
    // hands the blocks cached in the library's private pool of the CURRENT device back to the driver
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceSynchronize());
    cudaMemPool_t pool = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_pool_mutex);
        if (dev >= 0 && dev < kMaxDevices) pool = g_pools[dev];
    }
    if (pool) CUDA_TRY(cudaMemPoolTrimTo(pool, 0));
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
