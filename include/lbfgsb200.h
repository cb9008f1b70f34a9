/*
 * lbfgsb200.h -- C ABI of the B200-native FP64 L-BFGS hot path.
 *
 * Drop-in boundary for the solver surface of ndzajic1/cuda-lbfgs.  Every entry
 * point cites the reference interface it replaces:
 *   seq/ = sequential-implementation/, par/ = parallel-implementation/.
 *
 * Plain C: pointers, sizes, PODs.  No torch / STL types.  All device work runs
 * in hand-written sm_100a kernels; there is NO CPU fallback: on a machine
 * without a usable CUDA device every compute entry point fails with
 * LBFGSB200_ERR_CUDA.
 *
 * The C++ shim with the reference's exact `LBFGS(f, grad, x0, method, ...)`
 * signature is include/lbfgsb200_compat.hpp.
 */
#ifndef LBFGSB200_H
#define LBFGSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LBFGSB200_VERSION 200
#define LBFGSB200_MAX_M 64        /* history pairs (reference default m=10; sweep goes to 50) */
#define LBFGSB200_TRACE_COLS 8
#define LBFGSB200_UNIQUE_ID_BYTES 128
#define LBFGSB200_PROFILE_CLASSES 6

/* line searches selectable by name in the reference: seq/lbfgs.cpp:40-70 */
typedef enum {
    LBFGSB200_LS_BACKTRACKING = 0,      /* seq/line_search.cpp:19-30,  par/line_search.cpp:25-43  */
    LBFGSB200_LS_INTERPOLATION = 1,     /* seq/line_search.cpp:57-121, par/line_search.cpp:156-228 */
    LBFGSB200_LS_WOLFE = 2,             /* seq/line_search.cpp:125-189, par/line_search.cpp:298-369 */
    LBFGSB200_LS_BACKTRACKING_WOLFE = 3 /* seq/line_search.cpp:33-55,  par/line_search.cpp:45-154 */
} lbfgsb200_ls_t;

/* which tree's line-search code and constants are followed:
 * SEQ = seq/line_search.cpp + seq/config.h (C2=0.9, cubicInterpolate),
 * PAR = par/line_search.cpp + par/constants.h (C2=0.7, safeCubicInterpolate, 0.5 floor) -- what par/L-BFGS.cu calls,
 * PAR_INLINED = the copies of those searches inlined in the other CUDA solvers (par/L-BFGS-Wolfe.cu:260-349,
 *   par/L-BFGS-Interpolation.cu:259-342, par/L-BFGS-Backtracking.cu:292-341, par/L-BFGS-Backtracking_Wolfe.cu:262-397):
 *   as PAR, except that f(x_k) is taken at the last trial point the previous search evaluated, the Wolfe bracket's
 *   f_lo starts at f(x_0) in every iteration, backtracking uses the textbook Armijo test with TOL=1e-10, and
 *   interpolation applies its 0.5 floor to every exit.  Pinned against those programs run on a B200
 *   (tests/golden/cuda_reference_traces.json). */
typedef enum { LBFGSB200_FLAVOR_SEQ = 0, LBFGSB200_FLAVOR_PAR = 1, LBFGSB200_FLAVOR_PAR_INLINED = 2 } lbfgsb200_flavor_t;

/* outer-loop semantics:
 * SEQ  = seq/lbfgs.cpp:72-199 (test ||g||<tol before the step, store a pair only if s.y>0,
 *        non-finite rho / bad gamma / non-descent => d=-g),
 * CUDA = par/L-BFGS*.cu (ring slot always overwritten, pairs with s.y<=1e-10 skipped,
 *        gamma fallback 1.0, test ||g||<=tol after the step, no descent safeguard) */
typedef enum { LBFGSB200_PROFILE_SEQ = 0, LBFGSB200_PROFILE_CUDA = 1 } lbfgsb200_profile_t;

/* search-direction algorithm: explicit two-loop (seq/lbfgs.cpp:93-143) or the
 * compact / Gram form that reads all 2m history vectors once (default; m <= 50).  For the built-in
 * objectives the compact form runs as two fused streaming kernels per iteration: accept step + the
 * inner products of the next direction, and direction + first line-search trial. */
typedef enum {
    LBFGSB200_DIR_TWO_LOOP = 0,
    LBFGSB200_DIR_COMPACT = 1,
    LBFGSB200_DIR_AUTO = 2 /* default: compact for m <= 50, explicit two-loop above */
} lbfgsb200_dir_t;

/* built-in device objectives: par/functions.cpp:6-49, seq/benchmark.cpp:16-56 */
typedef enum {
    LBFGSB200_OBJ_QUADRATIC = 0,
    LBFGSB200_OBJ_ROSENBROCK = 1,
    LBFGSB200_OBJ_TRIDIAG = 2,
    LBFGSB200_OBJ_DEVICE_CALLBACK = 100 /* lbfgsb200_create_callback */
} lbfgsb200_obj_t;

/* User objective on the device: the replacement for the reference's host callbacks
 * `function<double(vector<double>)> f` + `function<vector<double>(vector<double>)> grad`
 * (seq/lbfgs.h:18-19).  Evaluate at the trial point x + (*d_alpha)*d WITHOUT materialising it:
 * write grad f there to g_out[0..n) and { f, grad.d, grad.grad } to d_out3[0..3) (all device
 * memory; d_alpha is a DEVICE scalar, so no host synchronisation is needed).  Enqueue everything
 * on cuda_stream (a cudaStream_t) and return 0, or non-zero to abort the solve.  n is the number of elements this
 * solver owns, global_offset the global index of the first (0 unless the solver is sharded).
 * lbfgsb200_dot() / lbfgsb200_nrm2() may be used for the reductions (host-stepped loop only: they allocate). */
typedef int (*lbfgsb200_fg_device_fn)(const double *x, const double *d, const double *d_alpha,
                                      double *g_out, double *d_out3, size_t n, size_t global_offset,
                                      void *user, void *cuda_stream);

typedef enum {
    LBFGSB200_CONVERGED = 0,       /* "Converged!"                 seq/lbfgs.cpp:82  */
    LBFGSB200_MAX_ITER = 1,        /* "Maximum iterations reached" seq/lbfgs.cpp:201 */
    LBFGSB200_LS_FAILED = 2,       /* "Line search failed"         seq/lbfgs.cpp:166 */
    LBFGSB200_RUNNING = 3,         /* iterate() returned before any exit condition */
    LBFGSB200_ERR_INVALID = -1,    /* std::invalid_argument in the reference (seq/lbfgs.cpp:69) */
    LBFGSB200_ERR_CUDA = -2,       /* checkCudaError + exit(1) in the reference (par/L-BFGS.cu:76-83) */
    LBFGSB200_ERR_NCCL = -3,
    LBFGSB200_ERR_NOMEM = -4
} lbfgsb200_status_t;

/* Replaces the reference's positional arguments (seq/lbfgs.h:17-25: max_iterations,
 * m, tolerance, line_search_method) and its compile-time constants
 * (seq/config.h:5-17, par/constants.h:5-21). */
typedef struct {
    int m;               /* history size, 1..LBFGSB200_MAX_M */
    int max_iterations;
    double tolerance;
    int line_search;     /* lbfgsb200_ls_t */
    int flavor;          /* lbfgsb200_flavor_t */
    int profile;         /* lbfgsb200_profile_t */
    int direction;       /* lbfgsb200_dir_t */
    double c1;           /* C1 */
    double c2;           /* C2 */
    double step0;        /* INITIAL_STEP_SIZE */
    double shrink;       /* BACKTRACKING_ALPHA */
    double backtracking_tol; /* BACKTRACKING_TOL */
    double wolfe_min;    /* WOLFE_INTERP_MIN */
    int ls_max_trials;   /* 20 in the reference (seq/line_search.cpp:73, :143) */
    int use_graph;       /* 1 (default): whole iteration loop runs as one CUDA graph (device-side control flow) */
    int verbose;         /* 1: print the reference's per-iteration line (seq/lbfgs.cpp:77-78) from the trace */
    int grid_ctas;       /* 0 = auto (4 CTAs x SM count); tuning/testing knob */
    int num_gpus;        /* lbfgsb200_solve only: 1 = one GPU; P > 1 = shard over devices 0..P-1 of this process (one
                          * worker thread per GPU, NVLink mailboxes, needs peer access); 0 = automatic: as many visible
                          * GPUs as keep >= 2^23 elements per shard (LBFGSB200_NUM_GPUS in the environment overrides
                          * the automatic choice).  The reference is called as one function from one thread
                          * (seq/benchmark.cpp:94, par/L-BFGS-Wolfe.cu:473); this keeps that call shape on 8 GPUs. */
} lbfgsb200_params_t;

typedef struct {
    int status;          /* lbfgsb200_status_t */
    int64_t iterations;  /* completed steps */
    int64_t trial_evals; /* fused line-search evaluations launched */
    int64_t kernel_launches;
    double f;            /* f at the returned x */
    double gnorm;        /* ||grad f|| at the returned x */
    double device_ms;    /* CUDA-event time of the last iterate()/solve() device region */
    double bytes_moved;  /* algorithmic HBM bytes of that region (bytes model) */
    double f0;           /* f(x0) and ||grad f(x0)||: the reference's "Iteration 0" line (seq/lbfgs.cpp:77-78) */
    double gnorm0;
    int flow;            /* which kernels ran: 0 explicit two-loop, 1 compact (pass A, combine, trial, accept as separate
                          * kernels), 2 compact fused (k_accept_gram + k_combine_trial; m <= 25 with the built-in objectives) */
    int graph;           /* 1: the iterations of the last call ran as one CUDA graph */
    int num_gpus;        /* GPUs the solve ran on */
    int reserved;
} lbfgsb200_result_t;

typedef struct lbfgsb200_solver lbfgsb200_solver_t; /* one per GPU / rank */
typedef struct lbfgsb200_comm lbfgsb200_comm_t;     /* NCCL communicator wrapper */

/* ------------------------------------------------------------------ */
/* library                                                             */
/* ------------------------------------------------------------------ */
int lbfgsb200_version(void);
const char *lbfgsb200_strerror(int status);
const char *lbfgsb200_last_error(void); /* thread-local detail string of the last failure */
int lbfgsb200_device_count(void);       /* 0 when no CUDA device is usable */

/* defaults = the reference's constants for the given line-search tree
 * (seq/config.h or par/constants.h) and lbfgs.h defaults (m=10, max_it=1000, tol=1e-5); compact direction,
 * graph mode, automatic GPU count.  Constants with which a device-side search could not terminate (shrink
 * outside (0,1), non-positive step0 / tolerances) are refused by create / solve with LBFGSB200_ERR_INVALID. */
int lbfgsb200_params_default(lbfgsb200_params_t *p, int flavor);

/* ------------------------------------------------------------------ */
/* solver: replaces LBFGS() seq/lbfgs.cpp:17-203 and LBFGS_CUDA()      */
/* par/L-BFGS.cu:105-382 (+ the four inlined-line-search variants)     */
/* ------------------------------------------------------------------ */

/* One-shot, host buffers in and out (the reference's by-value vector<double> x0 and
 * returned vector<double>).  trace may be NULL; rows of LBFGSB200_TRACE_COLS doubles:
 * k, f, ||g||, alpha, trials, history size, x[0], x[n/2]. */
int lbfgsb200_solve(int objective, size_t n, const double *x0_host, double *x_out_host,
                    const lbfgsb200_params_t *params, lbfgsb200_result_t *result,
                    double *trace, size_t trace_rows);
/* the GPU count lbfgsb200_solve uses for params.num_gpus = requested and a problem of n elements */
int lbfgsb200_resolve_num_gpus(int requested, size_t n);

/* Resumable form.  n_global is the full problem size; with comm != NULL this rank owns the
 * contiguous shard given by lbfgsb200_shard_range(n_global, rank, nranks). */
int lbfgsb200_create(lbfgsb200_solver_t **out, int objective, size_t n_global,
                     const lbfgsb200_params_t *params, lbfgsb200_comm_t *comm,
                     size_t trace_rows);
/* Same, with a user device objective instead of a built-in one (two-loop or compact direction, all line searches).
 * fn is called once per line-search trial and once more at the accepted step.  With params.use_graph (the default)
 * the calls are RECORDED: fn runs once per call site while the solver captures its CUDA graph, and the kernels /
 * async copies it enqueued on the stream it was given are replayed from the device for every evaluation (the pointer
 * arguments are the same in every iteration; alpha is read from *d_alpha on the device).  A callback that cannot be
 * captured (it synchronises, allocates -- lbfgsb200_dot / lbfgsb200_nrm2 allocate their scratch -- or works on
 * another stream without joining it back) makes the solver fall back to its host-stepped loop, where fn is called by
 * the host before every evaluation; use_graph = 0 selects that loop outright. */
int lbfgsb200_create_callback(lbfgsb200_solver_t **out, lbfgsb200_fg_device_fn fn, void *user,
                              size_t n, const lbfgsb200_params_t *params, size_t trace_rows);
/* The same on a sharded solver (comm != NULL: this rank owns lbfgsb200_shard_range(n_global, rank, nranks)).  fn gets
 * this rank's shard (x, d, g_out: n_local elements; global_offset = index of its first element) and writes the
 * PARTIAL sums { f, grad.d, grad.grad } of its shard to d_out3; the solver adds them over the ranks in rank order.
 * An objective that couples neighbouring elements reads the neighbours' boundary values from
 * lbfgsb200_device_halo(). */
int lbfgsb200_create_callback_sharded(lbfgsb200_solver_t **out, lbfgsb200_fg_device_fn fn, void *user,
                                      size_t n_global, const lbfgsb200_params_t *params,
                                      lbfgsb200_comm_t *comm, size_t trace_rows);
/* DEVICE pointer to six doubles { xL, xR, dL, dR, gL, gR }: the LAST element of the left neighbour's shard and the
 * FIRST element of the right neighbour's, of the current iterate x, the search direction d and the gradient g
 * (zeros where there is no neighbour).  Refreshed by the solver once per iteration before fn is called, so the
 * neighbours' trial-point values are xL + alpha*dL and xR + alpha*dR.  Valid for the life of the solver; read it
 * from device code only (a kernel the callback launches). */
const double *lbfgsb200_device_halo(const lbfgsb200_solver_t *s);
/* x0: this rank's shard (n_local doubles), host or device pointer.  Evaluates f(x0), grad f(x0)
 * (seq/lbfgs.cpp:28-30). */
int lbfgsb200_set_x0(lbfgsb200_solver_t *s, const double *x0_local);
/* Runs at most `iterations` further steps (or until converged / failed / max_iterations). */
int lbfgsb200_iterate(lbfgsb200_solver_t *s, int64_t iterations);
/* As iterate(), but forces the host-stepped path and records a CUDA-event pair around every
 * streaming kernel.  class_ms[c] / class_launches[c] receive the summed device time and launch
 * count of: 0 two-loop passes, 1 trial evaluations, 2 accept/update, 3 other streaming kernels,
 * 4 compact pass A (Gram rows), 5 compact pass B (combine). */
int lbfgsb200_iterate_profiled(lbfgsb200_solver_t *s, int64_t iterations,
                               double class_ms[LBFGSB200_PROFILE_CLASSES],
                               int64_t class_launches[LBFGSB200_PROFILE_CLASSES]);
int lbfgsb200_get_x(lbfgsb200_solver_t *s, double *x_local_out); /* host or device pointer */
int lbfgsb200_get_result(lbfgsb200_solver_t *s, lbfgsb200_result_t *r);
int64_t lbfgsb200_get_trace(lbfgsb200_solver_t *s, double *rows, size_t max_rows);
size_t lbfgsb200_local_size(const lbfgsb200_solver_t *s);
/* Checkpoint / resume (new; the reference has none -- SURVEY.md 5): dump this rank's complete solver
 * state (iterate, gradient, direction, (s, y) ring buffer, device scalars, trace) to a file, and load
 * it into a handle created with the same n, m, objective, direction and rank layout.  A resumed run
 * continues bit-for-bit like the uninterrupted one. */
int lbfgsb200_checkpoint_save(lbfgsb200_solver_t *s, const char *path);
int lbfgsb200_checkpoint_load(lbfgsb200_solver_t *s, const char *path);
void lbfgsb200_destroy(lbfgsb200_solver_t *s);

/* Contiguous shard of rank r: every rank owns (n / nranks) rounded down to an even count, the
 * last rank takes the remainder, so each shard starts on a 16-byte boundary of the global
 * vector.  Pure host arithmetic (no CUDA needed). */
void lbfgsb200_shard_range(size_t n_global, int rank, int nranks, size_t *offset, size_t *n_local);

/* ------------------------------------------------------------------ */
/* multi-GPU plumbing (new work; the reference has none)               */
/* ------------------------------------------------------------------ */
int lbfgsb200_comm_unique_id(char id[LBFGSB200_UNIQUE_ID_BYTES]);
int lbfgsb200_comm_create(lbfgsb200_comm_t **out, const char id[LBFGSB200_UNIQUE_ID_BYTES],
                          int rank, int nranks);
/* In-process communicators for nranks devices driven by threads of ONE process (what lbfgsb200_solve builds for
 * num_gpus > 1): out[r] belongs to devices[r].  Peer access instead of IPC, no NCCL.  Each rank's solver must then be
 * created, fed and iterated from its own host thread with devices[r] current. */
int lbfgsb200_comm_create_local(lbfgsb200_comm_t **out, const int *devices, int nranks);
void lbfgsb200_comm_destroy(lbfgsb200_comm_t *c);

/* ------------------------------------------------------------------ */
/* unit-test surface: the vector_utils / kernel layer.  Device pointers */
/* in, results left in DEVICE scalars (no host sync).  stream is a      */
/* cudaStream_t passed as void*.                                        */
/* ------------------------------------------------------------------ */
/* dotProduct  seq/vector_utils.cpp:32-41 ; cublasDdot sites par/L-BFGS.cu:219-267 */
int lbfgsb200_dot(const double *a, const double *b, size_t n, double *d_out, void *stream);
/* vectorNorm  seq/vector_utils.cpp:78-86 ; par/L-BFGS.cu:346-347 */
int lbfgsb200_nrm2(const double *a, size_t n, double *d_out, void *stream);
/* y += alpha*x with alpha read from a device scalar; cublasDaxpy sites par/L-BFGS.cu:233,:272 */
int lbfgsb200_axpy(const double *d_alpha, const double *x, double *y, size_t n, void *stream);
/* scalarProduct seq/vector_utils.cpp:43-51 ; scaleByRho par/L-BFGS.cu:65-73 ; out may alias x */
int lbfgsb200_scal(const double *d_alpha, const double *x, double *out, size_t n, void *stream);
/* Fused trial evaluation at x + alpha*d without materialising it (replaces updateSolution +
 * host f/grad + ddot, par/L-BFGS-Wolfe.cu:276-311).  g_out may be NULL.  d_out3 receives
 * f, grad.d, grad.grad. */
int lbfgsb200_eval_trial(int objective, const double *x, const double *d, const double *d_alpha,
                         size_t n, double *g_out, double *d_out3, void *stream);
/* Two-loop recursion (seq/lbfgs.cpp:93-143) over h pairs stored oldest first as rows of S and Y
 * (row stride `stride` doubles, 16-byte aligned rows).  Writes d; d_out2 = {g.d, fell_back}. */
int lbfgsb200_two_loop(const double *g, const double *S, const double *Y, int h, size_t n,
                       size_t stride, double *d, double *d_out2, void *stream);
/* Accept step (updateSolution + updateVectors, par/L-BFGS.cu:55-63, :19-31, + grad):
 * x += alpha d, g = grad f(x), s = x_new - x, y = g_new - g ; d_out5 = f, g.g, s.y, y.y, s.g */
int lbfgsb200_accept(int objective, double *x, const double *d, double *g, const double *d_alpha,
                     size_t n, double *s, double *y, double *d_out5, void *stream);

/* Host helper: x0 exactly as the reference mains draw it -- std::mt19937(seed) +
 * std::uniform_real_distribution<>(lo,hi), sequentially (seq/main.cpp:34-43,
 * par/L-BFGS-Wolfe.cu:458-465).  Writes elements [offset, offset+count). */
void lbfgsb200_x0_uniform(unsigned seed, double lo, double hi, size_t offset, size_t count,
                          double *out_host);

/* Raw CUDA-runtime helpers for hosts that do not link the CUDA runtime themselves (the
 * ctypes harness): pinned host buffers, device buffers, copies, device selection. */
void *lbfgsb200_host_alloc(size_t bytes);
void lbfgsb200_host_free(void *p);
void *lbfgsb200_device_alloc(size_t bytes);
void lbfgsb200_device_free(void *p);
int lbfgsb200_memcpy(void *dst, const void *src, size_t bytes);
int lbfgsb200_set_device(int ordinal);
/* Solver memory is served from a PRIVATE stream-ordered pool per device (the device's default pool, which other
 * cudaMallocAsync users of the process share, is never touched) and stays cached there after lbfgsb200_destroy() so
 * that the next solver does not pay the allocation again.  The library trims the pool and retries by itself when one
 * of its allocations runs out of memory; this hands the cached blocks of the CURRENT device back on request. */
int lbfgsb200_trim_memory(void);
/* Diagnostic: with LBFGSB200_TIMELINE=<rows> in the environment at create time every scalar kernel records
 * (op, %globaltimer ns at entry, at exit); this copies up to cap_rows rows of 3 u64 out and returns the count.
 * The gaps between rows are the vector kernels plus launch latency (benchmarks/timeline.py). */
long lbfgsb200_debug_timeline(lbfgsb200_solver_t *s, unsigned long long *rows, size_t cap_rows, int reset);
/* free / total device memory as the driver sees it (cached arenas count as used) */
int lbfgsb200_mem_info(size_t *free_bytes, size_t *total_bytes);
int lbfgsb200_device_sync(void);

#ifdef __cplusplus
}
#endif
#endif /* LBFGSB200_H */
