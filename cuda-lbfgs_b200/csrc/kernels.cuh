// kernels.cuh -- the streaming (HBM-bound) kernels of the L-BFGS hot path, sm_100a.
//
// Every kernel here is a pure stream over n FP64 elements:
//   * 128-bit (double2) coalesced loads/stores, kUnroll independent 16-byte requests per
//     stream per thread in flight before the first use (memory-level parallelism instead
//     of occupancy: 4 CTAs x 256 threads per SM, <= 64 registers);
//   * CTA tiles of kTileVec double2 (16 KB per stream) handed out round-robin over a grid
//     that is a multiple of the SM count (148 x 4 on B200);
//   * reductions are deterministic: fixed per-thread order, xor-butterfly warp shuffle,
//     fixed-order block tree, ONE partial per CTA and quantity, no atomics.  The partials
//     are summed in index order by the 1-CTA scalar kernel (scalar_ops.cuh);
//   * every scalar a kernel needs (coefficients, alpha, ring slots) is read from the
//     device-resident DevState: nothing is passed from the host per iteration, so the
//     same launches replay inside a CUDA graph.
// The translation unit is compiled with -fmad=false: element-wise expressions keep the
// reference's operation order and round exactly like its x86-64 (no-FMA) build, so x, g, s,
// y are bit-identical to the oracle given the same scalars.
#pragma once
#include <cuda_runtime.h>

#include "state.h"

namespace lb {

// ------------------------------------------------------------------
// reductions
// ------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Sums NQ per-thread values over the CTA in a fixed order and writes one partial per quantity:
// partials[q * gridDim.x + blockIdx.x].  All threads must call it.
template <int NQ>
__device__ __forceinline__ void block_emit(const double (&v)[NQ], double *__restrict__ partials)
{
    __shared__ double sm[NQ][kThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        double w = warp_sum(v[q]);
        if (lane == 0) sm[q][warp] = w;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            double w = (lane < kThreads / 32) ? sm[q][lane] : 0.0;
#pragma unroll
            for (int o = (kThreads / 64); o > 0; o >>= 1)
                w += __shfl_xor_sync(0xffffffffu, w, o);
            if (lane == 0) partials[(size_t)q * gridDim.x + blockIdx.x] = w;
        }
    }
}

// fixed-order combine of the kUnroll per-slot accumulators
__device__ __forceinline__ double fold(const double (&a)[kUnroll])
{
    static_assert(kUnroll == 4, "fold assumes 4 accumulators");
    return (a[0] + a[1]) + (a[2] + a[3]);
}

__device__ __forceinline__ double2 ld2(const double *p, long long vec_index)
{
    return reinterpret_cast<const double2 *>(p)[vec_index];
}
__device__ __forceinline__ void st2(double *p, long long vec_index, double2 v)
{
    reinterpret_cast<double2 *>(p)[vec_index] = v;
}

// Tile iteration shared by all kernels: full tiles run unguarded and fully unrolled, the
// ragged last tile runs guarded; the odd tail element (n & 1) is handled by the caller.
//   full(j0)         -- this thread owns double2 items j0 + u*kThreads, u < kUnroll, all valid
//   guard(j0, nvec)  -- same items, each to be checked against nvec
template <class Full, class Guard>
__device__ __forceinline__ void tile_loop(long long n, Full full, Guard guard)
{
    const long long nvec = n >> 1;
    const long long ntiles = (nvec + kTileVec - 1) / kTileVec;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const long long j0 = t * kTileVec + threadIdx.x;
        if ((t + 1) * kTileVec <= nvec)
            full(j0);
        else
            guard(j0, nvec);
    }
}

// ------------------------------------------------------------------
// generic vector kernels (unit-test surface + rare paths)
// ------------------------------------------------------------------

// dot(a, b) -> partials[0][cta]   (seq/vector_utils.cpp:32-41, cublasDdot sites)
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
k_dot(const double *__restrict__ a, const double *__restrict__ b, long long n,
      double *__restrict__ partials)
{
    double acc[kUnroll] = {0, 0, 0, 0};
    tile_loop(n,
        [&](long long j0) {
            double2 va[kUnroll], vb[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) va[u] = ld2(a, j0 + u * kThreads);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) vb[u] = ld2(b, j0 + u * kThreads);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) acc[u] += va[u].x * vb[u].x + va[u].y * vb[u].y;
        },
        [&](long long j0, long long nvec) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const long long j = j0 + u * kThreads;
                if (j < nvec) {
                    double2 va = ld2(a, j), vb = ld2(b, j);
                    acc[u] += va.x * vb.x + va.y * vb.y;
                }
            }
        });
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc[0] += a[n - 1] * b[n - 1];
    double v[1] = {fold(acc)};
    block_emit<1>(v, partials);
}

// out = alpha * x (+ y if AXPY), alpha read from a device scalar
// (cublasDaxpy sites par/L-BFGS.cu:233,:272 ; scaleByRho par/L-BFGS.cu:65-73)
template <bool AXPY>
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
k_axpy(const double *__restrict__ d_alpha, const double *__restrict__ x, const double *y,
       double *out, long long n)
{
    const double alpha = *d_alpha;
    tile_loop(n,
        [&](long long j0) {
            double2 vx[kUnroll], vy[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) vx[u] = ld2(x, j0 + u * kThreads);
            if (AXPY) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) vy[u] = ld2(y, j0 + u * kThreads);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                double2 r;
                r.x = alpha * vx[u].x;
                r.y = alpha * vx[u].y;
                if (AXPY) { r.x = vy[u].x + r.x; r.y = vy[u].y + r.y; }
                st2(out, j0 + u * kThreads, r);
            }
        },
        [&](long long j0, long long nvec) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const long long j = j0 + u * kThreads;
                if (j < nvec) {
                    double2 vx = ld2(x, j), r;
                    r.x = alpha * vx.x;
                    r.y = alpha * vx.y;
                    if (AXPY) { double2 vy = ld2(y, j); r.x = vy.x + r.x; r.y = vy.y + r.y; }
                    st2(out, j, r);
                }
            }
        });
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        double r = alpha * x[n - 1];
        if (AXPY) r = y[n - 1] + r;
        out[n - 1] = r;
    }
}

// d = -g   (negateVector par/L-BFGS.cu:44-52 ; seq/lbfgs.cpp:90, :106, :122, :151)
// Runs only when the scalar kernel raised st->steepest; otherwise exits at once.
__global__ void __launch_bounds__(kThreads, kCtasPerSm) k_steepest(const DevState *__restrict__ st)
{
    if (st->ctrl.done || !st->steepest) return;
    const double *__restrict__ g = st->g;
    double *__restrict__ w = st->w;
    const long long n = st->n;
    tile_loop(n,
        [&](long long j0) {
            double2 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) v[u] = ld2(g, j0 + u * kThreads);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                v[u].x = -v[u].x;
                v[u].y = -v[u].y;
                st2(w, j0 + u * kThreads, v[u]);
            }
        },
        [&](long long j0, long long nvec) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const long long j = j0 + u * kThreads;
                if (j < nvec) {
                    double2 v = ld2(g, j);
                    v.x = -v.x;
                    v.y = -v.y;
                    st2(w, j, v);
                }
            }
        });
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) w[n - 1] = -g[n - 1];
}

// (rare) s_newest . g when the previous pair was rejected by the curvature gate
__global__ void __launch_bounds__(kThreads, kCtasPerSm) k_dot_sg(const DevState *__restrict__ st)
{
    if (st->ctrl.done || !st->need_sg) return;
    const double *__restrict__ a = st->S + (size_t)slot_of(*st, st->h - 1) * st->stride;
    const double *__restrict__ b = st->g;
    const long long n = st->n;
    double acc[kUnroll] = {0, 0, 0, 0};
    tile_loop(n,
        [&](long long j0) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                double2 va = ld2(a, j0 + u * kThreads), vb = ld2(b, j0 + u * kThreads);
                acc[u] += va.x * vb.x + va.y * vb.y;
            }
        },
        [&](long long j0, long long nvec) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const long long j = j0 + u * kThreads;
                if (j < nvec) {
                    double2 va = ld2(a, j), vb = ld2(b, j);
                    acc[u] += va.x * vb.x + va.y * vb.y;
                }
            }
        });
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) acc[0] += a[n - 1] * b[n - 1];
    double v[1] = {fold(acc)};
    block_emit<1>(v, st->partials);
}

// ------------------------------------------------------------------
// two-loop recursion passes (seq/lbfgs.cpp:93-143, par/L-BFGS.cu:212-276)
//
// One launch per stored pair and loop.  Each pass streams q (or r), ONE history vector it
// applies and ONE history vector it takes the next inner product with, so the scalar the
// next pass needs is ready when this one ends: 3 reads + 1 write per pass instead of the
// reference's ddot (2 reads) + daxpy (2 reads + 1 write) + host sync.
//   MODE 0  loop 1, position p>0 : q -= a_p y_p ;            dot = s_{p-1} . q
//   MODE 1  loop 1, position 0   : q -= a_0 y_0 ;            dot = y_0 . (gamma q)
//   MODE 2  loop 2, position p<h-1: r += (a_p - b_p) s_p ;   dot = y_{p+1} . r
//   MODE 3  loop 2, position h-1 : r += (...) s_p ; d = -r ; dot = g . d
// `gs` scales the loop-2 input: gamma on the first loop-2 pass (r = gamma q, seq/lbfgs.cpp:
// 126-130), exactly 1.0 afterwards (x*1.0 is exact, so one code path serves both).
// ------------------------------------------------------------------
template <int MODE>
__device__ __forceinline__ double2 pass_elem(double2 in, double2 v, double2 dv, double c, double gs,
                                             double &acc)
{
    double2 o;
    if (MODE == 0) {
        o.x = in.x - c * v.x;
        o.y = in.y - c * v.y;
        acc += dv.x * o.x + dv.y * o.y;
    } else if (MODE == 1) {
        o.x = in.x - c * v.x;
        o.y = in.y - c * v.y;
        acc += v.x * (o.x * gs) + v.y * (o.y * gs);
    } else if (MODE == 2) {
        o.x = in.x * gs + v.x * c;
        o.y = in.y * gs + v.y * c;
        acc += dv.x * o.x + dv.y * o.y;
    } else {
        o.x = -(in.x * gs + v.x * c);
        o.y = -(in.y * gs + v.y * c);
        acc += dv.x * o.x + dv.y * o.y;
    }
    return o;
}

template <int MODE>
__device__ __forceinline__ void pass_run(const double *in, const double *__restrict__ v,
                                         const double *__restrict__ dv, double *out, double c,
                                         double gs, long long n, double *__restrict__ partials)
{
    constexpr bool kHasDv = (MODE != 1);
    double acc[kUnroll] = {0, 0, 0, 0};
    tile_loop(n,
        [&](long long j0) {
            double2 a[kUnroll], b[kUnroll], w[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) a[u] = ld2(in, j0 + u * kThreads);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) b[u] = ld2(v, j0 + u * kThreads);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                w[u] = kHasDv ? ld2(dv, j0 + u * kThreads) : make_double2(0, 0);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                st2(out, j0 + u * kThreads, pass_elem<MODE>(a[u], b[u], w[u], c, gs, acc[u]));
        },
        [&](long long j0, long long nvec) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const long long j = j0 + u * kThreads;
                if (j < nvec) {
                    double2 a = ld2(in, j), b = ld2(v, j);
                    double2 w = kHasDv ? ld2(dv, j) : make_double2(0, 0);
                    st2(out, j, pass_elem<MODE>(a, b, w, c, gs, acc[u]));
                }
            }
        });
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const long long e = n - 1;
        double2 a = make_double2(in[e], 0), b = make_double2(v[e], 0);
        double2 w = make_double2(kHasDv ? dv[e] : 0.0, 0);
        double2 o = pass_elem<MODE>(a, b, w, c, gs, acc[0]);
        out[e] = o.x;
    }
    double r[1] = {fold(acc)};
    block_emit<1>(r, partials);
}

// loop: 1 or 2 ; p: window position (0 = oldest pair).  Positions >= h exit at once, so the
// host (or the graph) always launches m of them.
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
k_two_loop_pass(const DevState *__restrict__ st, int loop, int p)
{
    const int h = st->h;
    if (st->ctrl.done || st->steepest || p >= h) return;
    const long long n = st->n;
    const size_t stride = (size_t)st->stride;
    const double c = st->coef;
    double *w = st->w;
    if (loop == 1) {
        const double *in = (p == h - 1) ? st->g : w; // q starts as the gradient (seq/lbfgs.cpp:95)
        const double *yv = st->Y + (size_t)slot_of(*st, p) * stride;
        if (p > 0) {
            const double *sv = st->S + (size_t)slot_of(*st, p - 1) * stride;
            pass_run<0>(in, yv, sv, w, c, 1.0, n, st->partials);
        } else {
            pass_run<1>(in, yv, nullptr, w, c, st->gamma, n, st->partials);
        }
    } else {
        const double gs = (p == 0) ? st->gamma : 1.0;
        const double *sv = st->S + (size_t)slot_of(*st, p) * stride;
        if (p < h - 1) {
            const double *yv = st->Y + (size_t)slot_of(*st, p + 1) * stride;
            pass_run<2>(w, sv, yv, w, c, gs, n, st->partials);
        } else {
            pass_run<3>(w, sv, st->g, w, c, gs, n, st->partials);
        }
    }
}

// ------------------------------------------------------------------
// objectives: three-point stencils on the trial point xt = x + alpha d
// ------------------------------------------------------------------
// Each functor returns, for element c with neighbours l and r (hl / hr: the neighbour
// exists in the GLOBAL vector), the f-term owned by c and the gradient entry g_c, with the
// reference's exact expression order.

struct ObjQuadratic { // par/functions.cpp:6-24
    static constexpr bool kStencil = false;
    __device__ static __forceinline__ void eval(double, double c, double, bool, bool, double &ft,
                                                double &g)
    {
        const double t = c - 1;
        ft = t * t;
        g = 2.0 * t;
    }
};

struct ObjRosenbrock { // par/functions.cpp:26-49
    static constexpr bool kStencil = true;
    __device__ static __forceinline__ void eval(double l, double c, double r, bool hl, bool hr,
                                                double &ft, double &g)
    {
        // loop index i=c:   term1 = 2(x_c-1) ; term2 = x_{c+1} - x_c^2
        //   f += 100*term2*term2 + (1-x_c)^2 ; g[c] += term1 - 400*x_c*term2 ; g[c+1] += 200*term2
        const double bl = c - l * l; // term2 of loop index c-1
        const double bc = r - c * c; // term2 of loop index c
        const double t2 = 1 - c;
        ft = hr ? (100.0 * bc * bc + t2 * t2) : 0.0;
        const double from_left = hl ? 200.0 * bl : 0.0;           // g[c] += 200*term2 (i = c-1)
        const double own = 2.0 * (c - 1) - 400.0 * c * bc;        // term1 - 400*x*term2 (i = c)
        g = hr ? (from_left + own) : from_left;
    }
};

struct ObjTridiag { // seq/benchmark.cpp:16-56, COEFFICIENT = 1000
    static constexpr bool kStencil = true;
    __device__ static __forceinline__ void eval(double l, double c, double r, bool hl, bool hr,
                                                double &ft, double &g)
    {
        const double diag = 1000.0 * c * c;
        ft = hr ? (diag + (1000.0 / 10.0) * c * r) : diag;
        double gv = 2.0 * 1000.0 * c;                 // gradient[i] = 2*COEFFICIENT*x[i]
        if (hl) gv = gv + (1000.0 / 10.0) * l;        // gradient[i+1] += c/10 * x[i]   (loop i = c-1)
        if (hr) gv = gv + (1000.0 / 10.0) * r;        // gradient[i]   += c/10 * x[i+1] (loop i = c)
        g = gv;
    }
};

struct EvalCtx {
    const double *x, *d;
    double alpha;
    long long n, goff, nglob;
    double xtL, xtR; // trial values of the neighbour shards' boundary elements (halo)
    double *d_store; // PEND trials only: where d = -g is written
};

// PEND (fused compact flow, rare): the descent safeguard (seq/lbfgs.cpp:147-153) fired after the combine pass, so
// the direction in memory is stale.  The trial then reads g instead, uses d = -g and stores it for the later
// trials and the accept step -- no separate d = -g pass, no extra launch in the common case.
template <bool PEND>
__device__ __forceinline__ double2 load_d2(const EvalCtx &c, long long j)
{
    double2 v = ld2(c.d, j);
    if (PEND) {
        v.x = -v.x;
        v.y = -v.y;
        st2(c.d_store, j, v);
    }
    return v;
}
template <bool PEND>
__device__ __forceinline__ double load_d1(const EvalCtx &c, long long e)
{
    const double v = c.d[e];
    return PEND ? -v : v;
}

// trial value of local element e, which may be one outside the shard (halo)
__device__ __forceinline__ double fetch_xt(const EvalCtx &c, long long e)
{
    if (e < 0) return c.xtL;
    if (e >= c.n) return c.xtR;
    return c.x[e] + c.alpha * c.d[e];
}
template <bool PEND>
__device__ __forceinline__ double fetch_xt_p(const EvalCtx &c, long long e)
{
    if (e < 0) return c.xtL;
    if (e >= c.n) return c.xtR;
    return c.x[e] + c.alpha * load_d1<PEND>(c, e);
}

// MODE_TRIAL: sums f, g.d, g.g, optional g store.  (replaces updateSolution + host f/grad +
//             ddot, par/L-BFGS-Wolfe.cu:276-311, without materialising x + alpha d)
// MODE_ACCEPT: x <- x + alpha d ; g <- grad ; s = x_new - x ; y = g_new - g ;
//             sums f, g.g, s.y, y.y, s.g  (updateSolution + updateVectors + ddot(g,g),
//             par/L-BFGS.cu:309-347 ; seq/lbfgs.cpp:159-181)
enum { MODE_TRIAL = 0, MODE_ACCEPT = 1 };

struct EvalOut {
    double *__restrict__ g_out; // TRIAL: optional gradient store (unit-test surface)
    double *__restrict__ xw;    // ACCEPT: new iterate (never aliases x, see note below)
    double *gw;                 // ACCEPT: gradient, read (old) and written (new) in place
    double *__restrict__ s_out;
    double *__restrict__ y_out;
};

// One double2 item (elements 2j, 2j+1).  x2/d2/go were loaded by the caller (batched, so the
// kUnroll requests per stream are all in flight before the first use).  Shuffles are executed
// by all 32 lanes; inactive lanes (ragged last tile) carry zeros and drop out afterwards.
// ex/ed: (unguarded path) the x and d of the one element outside the warp's 64-element span that
// lane 0 (element 2j-1) or lane 31 (element 2j+2) needs, prefetched by the caller in the same
// batch as the main loads so the stencil never waits for a second, dependent memory round trip.
template <class OBJ, int MODE, bool GUARD, bool PEND = false>
__device__ __forceinline__ void eval_item(const EvalCtx &c, long long j, long long nvec, bool active,
                                          double2 x2, double2 d2, double2 go, double ex, double ed,
                                          const EvalOut &o, double (&acc)[5])
{
    const int lane = threadIdx.x & 31;
    const double xt0 = x2.x + c.alpha * d2.x; // add(x, scalarProduct(alpha, d)): mul, then add
    const double xt1 = x2.y + c.alpha * d2.y;
    double l = 0.0, r = 0.0;
    if (OBJ::kStencil) {
        l = __shfl_up_sync(0xffffffffu, xt1, 1);
        r = __shfl_down_sync(0xffffffffu, xt0, 1);
        if (GUARD) {
            if (active) {
                if (lane == 0) l = fetch_xt_p<PEND>(c, 2 * j - 1);
                if (lane == 31 || j + 1 >= nvec) r = fetch_xt_p<PEND>(c, 2 * j + 2);
            }
        } else {
            const double et = ex + c.alpha * ed; // same two roundings as every other trial value
            if (lane == 0) l = (2 * j - 1 < 0) ? c.xtL : et;
            if (lane == 31) r = (2 * j + 2 >= c.n) ? c.xtR : et;
        }
    }
    if (GUARD && !active) return;
    const long long G0 = c.goff + 2 * j;
    const bool hl0 = G0 > 0, hr1 = (G0 + 1) < c.nglob - 1;
    double f0, g0, f1, g1;
    OBJ::eval(l, xt0, xt1, hl0, true, f0, g0);
    OBJ::eval(xt0, xt1, r, true, hr1, f1, g1);
    if (MODE == MODE_TRIAL) {
        acc[0] += f0 + f1;
        acc[1] += g0 * d2.x + g1 * d2.y;
        acc[2] += g0 * g0 + g1 * g1;
        if (o.g_out) st2(o.g_out, j, make_double2(g0, g1));
    } else {
        const double s0 = xt0 - x2.x, s1 = xt1 - x2.y; // s = x_new - x (seq/lbfgs.cpp:177)
        const double y0 = g0 - go.x, y1 = g1 - go.y;   // y = g_new - g (seq/lbfgs.cpp:178)
        st2(o.xw, j, make_double2(xt0, xt1));
        st2(o.gw, j, make_double2(g0, g1));
        st2(o.s_out, j, make_double2(s0, s1));
        st2(o.y_out, j, make_double2(y0, y1));
        acc[0] += f0 + f1;
        acc[1] += g0 * g0 + g1 * g1;
        acc[2] += s0 * y0 + s1 * y1;
        acc[3] += y0 * y0 + y1 * y1;
        acc[4] += s0 * g0 + s1 * g1;
    }
}

// odd tail element n-1 (scalar path, one thread)
template <class OBJ, int MODE, bool PEND = false>
__device__ __forceinline__ void eval_tail(const EvalCtx &c, const EvalOut &o, double (&acc)[5])
{
    const long long e = c.n - 1;
    const double xe = c.x[e], de = load_d1<PEND>(c, e);
    if (PEND) c.d_store[e] = de;
    const double xt = xe + c.alpha * de;
    double l = 0.0, r = 0.0;
    if (OBJ::kStencil) {
        l = fetch_xt_p<PEND>(c, e - 1);
        r = fetch_xt_p<PEND>(c, e + 1);
    }
    const long long G = c.goff + e;
    double ft, gv;
    OBJ::eval(l, xt, r, G > 0, G < c.nglob - 1, ft, gv);
    if (MODE == MODE_TRIAL) {
        acc[0] += ft;
        acc[1] += gv * de;
        acc[2] += gv * gv;
        if (o.g_out) o.g_out[e] = gv;
    } else {
        const double go = o.gw[e];
        const double s = xt - xe, y = gv - go;
        o.xw[e] = xt;
        o.gw[e] = gv;
        o.s_out[e] = s;
        o.y_out[e] = y;
        acc[0] += ft;
        acc[1] += gv * gv;
        acc[2] += s * y;
        acc[3] += y * y;
        acc[4] += s * gv;
    }
}

// NOTE on ACCEPT: warp-edge lanes fetch the OLD x of elements they do not own from memory, so
// the new iterate is never written over x: it goes to the alternate buffer `xw` (DevState::x_alt)
// and the scalar kernel swaps the two pointers afterwards.  g IS updated in place: a thread only
// ever reads the old g of the elements it owns.
template <class OBJ, int MODE, bool PEND = false>
__device__ __forceinline__ void eval_run(const EvalCtx &c, const EvalOut &o,
                                         double *__restrict__ partials)
{
    double acc[5] = {0, 0, 0, 0, 0};
    const long long n = c.n;
    tile_loop(n,
        [&](long long j0) {
            // loads are issued in batches of B items per stream before the first use; the accept
            // kernel carries 7 streams, so it batches 2 (no spills at 64 registers), the trial 4
            constexpr int B = (MODE == MODE_ACCEPT) ? 2 : kUnroll;
#pragma unroll
            for (int ub = 0; ub < kUnroll; ub += B) {
                double2 x2[B], d2[B], go[B];
                double ex[B], ed[B];
                const int lane = threadIdx.x & 31;
                if (OBJ::kStencil) {
                    // branch-free: every lane loads ONE extra element -- lanes 0 / 31 the neighbour
                    // just outside the warp's span, all other lanes their own first element (a
                    // sector the 16-byte load below touches anyway)
#pragma unroll
                    for (int u = 0; u < B; ++u) {
                        const long long j = j0 + (ub + u) * kThreads;
                        long long e = lane == 0 ? 2 * j - 1 : (lane == 31 ? 2 * j + 2 : 2 * j);
                        e = (e < 0 || e >= c.n) ? 2 * j : e;
                        ex[u] = c.x[e];
                        ed[u] = load_d1<PEND>(c, e);
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < B; ++u) ex[u] = ed[u] = 0.0;
                }
#pragma unroll
                for (int u = 0; u < B; ++u) x2[u] = ld2(c.x, j0 + (ub + u) * kThreads);
#pragma unroll
                for (int u = 0; u < B; ++u) d2[u] = load_d2<PEND>(c, j0 + (ub + u) * kThreads);
#pragma unroll
                for (int u = 0; u < B; ++u)
                    go[u] = (MODE == MODE_ACCEPT) ? ld2(o.gw, j0 + (ub + u) * kThreads) : make_double2(0, 0);
#pragma unroll
                for (int u = 0; u < B; ++u)
                    eval_item<OBJ, MODE, false, PEND>(c, j0 + (ub + u) * kThreads, 0, true, x2[u], d2[u], go[u], ex[u], ed[u], o, acc);
            }
        },
        [&](long long j0, long long nvec) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const long long j = j0 + u * kThreads;
                const bool active = j < nvec;
                double2 x2 = make_double2(0, 0), d2 = x2, go = x2;
                if (active) {
                    x2 = ld2(c.x, j);
                    d2 = load_d2<PEND>(c, j);
                    if (MODE == MODE_ACCEPT) go = ld2(o.gw, j);
                }
                eval_item<OBJ, MODE, true, PEND>(c, j, nvec, active, x2, d2, go, 0.0, 0.0, o, acc);
            }
        });
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) eval_tail<OBJ, MODE, PEND>(c, o, acc);
    constexpr int NQ = (MODE == MODE_TRIAL) ? 3 : 5;
    double v[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = acc[q];
    block_emit<NQ>(v, partials);
}

template <int MODE>
__device__ __forceinline__ void eval_dispatch(int objective, const EvalCtx &c, const EvalOut &o,
                                              double *partials)
{
    switch (objective) {
    case LBFGSB200_OBJ_QUADRATIC: eval_run<ObjQuadratic, MODE>(c, o, partials); break;
    case LBFGSB200_OBJ_ROSENBROCK: eval_run<ObjRosenbrock, MODE>(c, o, partials); break;
    default: eval_run<ObjTridiag, MODE>(c, o, partials); break;
    }
}

__device__ __forceinline__ EvalCtx make_ctx(const DevState *st, double alpha)
{
    EvalCtx c;
    c.x = st->x;
    c.d = st->w;
    c.alpha = alpha;
    c.n = st->n;
    c.goff = st->goff;
    c.nglob = st->nglob;
    c.xtL = st->xL + alpha * st->dL;
    c.xtR = st->xR + alpha * st->dR;
    c.d_store = nullptr;
    return c;
}

// One line-search trial at alpha = st->ls.alpha: partials f, g.d, g.g.  No global stores.
// One instantiation per objective (the host knows the objective): no run-time switch, and the
// register budget is that of one stencil, not the union of all of them.
template <class OBJ>
__global__ void __launch_bounds__(kThreads, kCtasPerSm) k_trial(const DevState *__restrict__ st)
{
    if (st->ctrl.done || !st->ctrl.ls_active) return;
    EvalCtx c = make_ctx(st, st->ls.alpha);
    const EvalOut o = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (st->pend_steepest) { // rare: d = -g on the fly, stored for the later trials and the accept step
        c.d = st->g;
        c.d_store = st->w;
        eval_run<OBJ, MODE_TRIAL, true>(c, o, st->partials);
    } else {
        eval_run<OBJ, MODE_TRIAL>(c, o, st->partials);
    }
}

// Accept the step alpha = st->ls.alpha (init != 0: alpha = 0 on a zeroed d, i.e. evaluate f, g at x0).
template <class OBJ>
__global__ void __launch_bounds__(kThreads, kCtasPerSmAccept) k_accept(const DevState *__restrict__ st, int init)
{
    if (st->ctrl.done) return;
    const EvalCtx c = make_ctx(st, init ? 0.0 : st->ls.alpha);
    const size_t sp = (size_t)spare_slot(*st) * (size_t)st->stride;
    const EvalOut o = {nullptr, st->x_alt, st->g, st->S + sp, st->Y + sp};
    eval_run<OBJ, MODE_ACCEPT>(c, o, st->partials);
}

typedef void (*trial_kernel_t)(const DevState *);
typedef void (*accept_kernel_t)(const DevState *, int);
inline trial_kernel_t trial_kernel_for(int objective)
{
    switch (objective) {
    case LBFGSB200_OBJ_QUADRATIC: return k_trial<ObjQuadratic>;
    case LBFGSB200_OBJ_ROSENBROCK: return k_trial<ObjRosenbrock>;
    default: return k_trial<ObjTridiag>;
    }
}
inline accept_kernel_t accept_kernel_for(int objective)
{
    switch (objective) {
    case LBFGSB200_OBJ_QUADRATIC: return k_accept<ObjQuadratic>;
    case LBFGSB200_OBJ_ROSENBROCK: return k_accept<ObjRosenbrock>;
    default: return k_accept<ObjTridiag>;
    }
}

// Accept step for user (callback) objectives: the gradient at x + alpha d was written to g_new
// by the callback and f to *d_f; everything else is the same update as k_accept.
__global__ void __launch_bounds__(kThreads, kCtasPerSm)
k_accept_generic(const DevState *__restrict__ st, const double *__restrict__ g_new, const double *__restrict__ d_f,
                 int init)
{
    if (st->ctrl.done) return;
    const double alpha = init ? 0.0 : st->ls.alpha;
    const double *x = st->x; // (user objectives: x_alt == x, the iterate is updated in place -- no __restrict__)
    const double *__restrict__ d = st->w;
    double *xw = st->x_alt;
    double *gw = st->g;
    const size_t sp = (size_t)spare_slot(*st) * (size_t)st->stride;
    double *__restrict__ s_out = st->S + sp;
    double *__restrict__ y_out = st->Y + sp;
    const long long n = st->n;
    double acc[5] = {0, 0, 0, 0, 0};
    auto item = [&](long long j) {
        const double2 x2 = ld2(x, j), d2 = ld2(d, j), gn = ld2(g_new, j), go = ld2(gw, j);
        const double xt0 = x2.x + alpha * d2.x, xt1 = x2.y + alpha * d2.y;
        const double s0 = xt0 - x2.x, s1 = xt1 - x2.y, y0 = gn.x - go.x, y1 = gn.y - go.y;
        st2(xw, j, make_double2(xt0, xt1));
        st2(gw, j, gn);
        st2(s_out, j, make_double2(s0, s1));
        st2(y_out, j, make_double2(y0, y1));
        acc[1] += gn.x * gn.x + gn.y * gn.y;
        acc[2] += s0 * y0 + s1 * y1;
        acc[3] += y0 * y0 + y1 * y1;
        acc[4] += s0 * gn.x + s1 * gn.y;
    };
    tile_loop(n,
        [&](long long j0) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) item(j0 + u * kThreads);
        },
        [&](long long j0, long long nvec) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                if (j0 + u * kThreads < nvec) item(j0 + u * kThreads);
        });
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        acc[0] += *d_f;
        if (n & 1) {
            const long long e = n - 1;
            const double xt = x[e] + alpha * d[e], gn = g_new[e];
            const double s = xt - x[e], y = gn - gw[e];
            xw[e] = xt; gw[e] = gn; s_out[e] = s; y_out[e] = y;
            acc[1] += gn * gn; acc[2] += s * y; acc[3] += y * y; acc[4] += s * gn;
        }
    }
    block_emit<5>(acc, st->partials);
}

// unit-test surface: explicit pointers, alpha from a device scalar, single shard
__global__ void __launch_bounds__(kThreads, kCtasPerSmAccept)
k_eval_explicit(int mode, int objective, const double *x, const double *d, const double *d_alpha,
                long long n, double *g_out, double *x_new, double *g_io, double *s_out, double *y_out,
                double *partials)
{
    EvalCtx c;
    c.x = x; c.d = d; c.alpha = *d_alpha; c.n = n; c.goff = 0; c.nglob = n; c.xtL = 0; c.xtR = 0; c.d_store = nullptr;
    const EvalOut o = {g_out, x_new, g_io, s_out, y_out};
    if (mode == MODE_TRIAL) eval_dispatch<MODE_TRIAL>(objective, c, o, partials);
    else eval_dispatch<MODE_ACCEPT>(objective, c, o, partials);
}

} // namespace lb
