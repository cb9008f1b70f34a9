// accept_gram.cuh -- the fused compact flow: two streaming kernels per iteration (+ one per extra trial).
//
//   k_accept_gram    accept step at the alpha the line search returned (x_new, g_new, s, y, f) FUSED with pass A
//                    of the NEXT direction (the 3 x (2h'+1) inner products of s_new, y_new, g_new with the new
//                    window).  The unfused sequence writes s, y, g_new in k_accept and reads them straight back
//                    in pass A; here they go from registers into the shared-memory tile the dot products run on:
//                    reads  2(h'-1) + 3  (kept history, x, d, g_old)     writes 4 (x_new, g_new, s, y)
//                    instead of 3 + 4 (accept) + 2h' + 1 (pass A)  =>  3 vector streams fewer per iteration.
//                    Replaces updateSolution + host grad + updateVectors + 2 D2D copies + ddot(g,g)
//                    (par/L-BFGS.cu:309-347) and all 3h+3 cublasDdot calls of the next iteration (:219-267).
//   k_combine_trial  pass B (d = -sum_j delta_j b_j, g.d) FUSED with the first line-search trial: every search
//                    starts at alpha = INITIAL_STEP_SIZE (seq/line_search.cpp:21, :72, :138), d is in registers, so
//                    f, grad f . d at x + step0 d cost one extra read of x instead of a 2-stream trial pass, a
//                    launch and a scalar kernel.  Replaces 2h cublasDaxpy + scaleByRho + negateVector
//                    (par/L-BFGS.cu:233-276) and the first updateSolution + host f/grad + ddot of the search.
//
// Both are HBM-bound streams; tensor cores are not used (FP64, ~1 flop/B).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "compact.cuh"
#include "kernels.cuh"
#include "state.h"

namespace lb {

constexpr int kHaloSlotDoubles = 16; // one 128-byte aligned slot per halo box (TMA destinations are 128-byte aligned)

__device__ __forceinline__ void named_barrier_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}

// Stage layout (doubles), T = tile width:
//   rows 0 .. J-1        the NEW window in column order [S 0..h'-1 | Y 0..h'-1 | g]: kept history rows arrive by
//                        TMA, rows h'-1 (s_new), 2h'-1 (y_new), 2h' (g_new) are written by phase 1
//   rows J, J+1, J+2     x, d, g_old tiles (TMA)
//   then 4 halo slots    x[i0-2..i0-1], x[i0+T..i0+T+1], d[i0-2..i0-1], d[i0+T..i0+T+1]
// hk = pairs kept from the old window (h if the ring is not full, else h-1: the oldest pair is evicted by the
// commit this kernel anticipates; a pair the curvature gate then rejects is handled by the scalar kernel:
// columns of s_new / y_new are ignored, and with a full ring the stand-alone pass A re-computes the g row).
__host__ __device__ inline size_t accept_gram_stage_doubles(int J, int T)
{
    return (size_t)(J + 3) * T + 4 * kHaloSlotDoubles;
}

template <class OBJ, int CW>
__global__ void __launch_bounds__(kWsThreads, 1)
k_accept_gram(const DevState *__restrict__ st, const ArenaMaps *__restrict__ maps, int T, int NG, int init)
{
    if (st->ctrl.done) return;
    extern __shared__ __align__(128) double tile[];
    __shared__ __align__(8) unsigned long long full[kGramStages], empty[kGramStages];
    __shared__ double fsum[kWsConsumerWarps];
    const int h_old = init ? 0 : st->h;
    const int ks = (h_old == st->m) ? 1 : 0; // oldest pair evicted by the anticipated commit
    const int hk = h_old - ks, hp = hk + 1, J = 2 * hp + 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kGramStages; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWsConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n = st->n;
    const long long ntiles = (n + T - 1) / T;
    const size_t stage_doubles = accept_gram_stage_doubles(J, T);
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int NE = kWsConsumerWarps / NG;
    double acc[CW][3];
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.0;
    double facc = 0.0;
    const int cw = warp - 1, cg = cw % NG, eg = cw / NG;

    if (warp == 0) {
        if (lane == 0) {
            const int ns = st->nslots;
            const int a0 = (st->base + ks) % ns;          // physical slot of the first kept pair
            const int ra = min(hk, ns - a0), rb = hk - ra; // the kept window is at most two runs of slots
            const int row_x = (st->x == st->arena0) ? 0 : 1;
            const unsigned bytes = (unsigned)(((size_t)(2 * hk + 3) * T + 8) * sizeof(double));
            for (long long k = 0; k < my_tiles; ++k) {
                const int stage = (int)(k % kGramStages);
                if (k >= kGramStages) mbar_wait(&empty[stage], (unsigned)(((k / kGramStages) - 1) & 1));
                const int col = (int)((blockIdx.x + k * (long long)gridDim.x) * T);
                double *dst = tile + stage * stage_doubles;
                double *halo = dst + (size_t)(J + 3) * T;
                mbar_expect_tx(&full[stage], bytes);
                if (ra) {
                    tma_load_2d(dst, &maps->run[ra], col, kArenaRowS + a0, &full[stage]);
                    tma_load_2d(dst + (size_t)hp * T, &maps->run[ra], col, kArenaRowS + ns + a0, &full[stage]);
                }
                if (rb) {
                    tma_load_2d(dst + (size_t)ra * T, &maps->run[rb], col, kArenaRowS, &full[stage]);
                    tma_load_2d(dst + (size_t)(hp + ra) * T, &maps->run[rb], col, kArenaRowS + ns, &full[stage]);
                }
                tma_load_2d(dst + (size_t)J * T, &maps->run[1], col, row_x, &full[stage]);
                tma_load_2d(dst + (size_t)(J + 1) * T, &maps->run[1], col, kArenaRowW, &full[stage]);
                tma_load_2d(dst + (size_t)(J + 2) * T, &maps->run[1], col, kArenaRowG, &full[stage]);
                tma_load_2d(halo, &maps->halo, col - 2, row_x, &full[stage]);
                tma_load_2d(halo + kHaloSlotDoubles, &maps->halo, col + T, row_x, &full[stage]);
                tma_load_2d(halo + 2 * kHaloSlotDoubles, &maps->halo, col - 2, kArenaRowW, &full[stage]);
                tma_load_2d(halo + 3 * kHaloSlotDoubles, &maps->halo, col + T, kArenaRowW, &full[stage]);
            }
        }
    } else {
        const double alpha = init ? 0.0 : st->ls.alpha;
        const double xtL = st->xL + alpha * st->dL, xtR = st->xR + alpha * st->dR;
        const long long goff = st->goff, nglob = st->nglob;
        double *__restrict__ x_new = st->x_alt;
        double *__restrict__ g_out = st->g;
        const size_t sp = (size_t)spare_slot(*st) * (size_t)st->stride;
        double *__restrict__ s_out = st->S + sp;
        double *__restrict__ y_out = st->Y + sp;
        const int r0 = hp - 1, r1 = 2 * hp - 1, r2 = 2 * hp;
        const int T2 = T >> 1, slice2 = T2 / NE;
        const int t = cw * 32 + lane; // phase-1 element of this thread within the tile (consumer warps 0..T/32-1)
        for (long long k = 0; k < my_tiles; ++k) {
            const int stage = (int)(k % kGramStages);
            mbar_wait(&full[stage], (unsigned)((k / kGramStages) & 1));
            double *cur = tile + stage * stage_doubles;
            // ---------------- phase 1: the accept step on this tile ----------------
            if (t < T) {
                const double *in_x = cur + (size_t)J * T, *in_d = in_x + T, *in_g = in_d + T;
                const double *halo = cur + (size_t)(J + 3) * T;
                const long long e = (blockIdx.x + k * (long long)gridDim.x) * T + t;
                double sv = 0.0, yv = 0.0, gv = 0.0;
                if (e < n) {
                    const double xc = in_x[t], dc = in_d[t], go = in_g[t];
                    const double xt = xc + alpha * dc; // add(x, scalarProduct(alpha, d)): mul, then add
                    double l = 0.0, r = 0.0;
                    if (OBJ::kStencil) {
                        if (e == 0) l = xtL;
                        else if (t > 0) l = in_x[t - 1] + alpha * in_d[t - 1];
                        else l = halo[1] + alpha * halo[2 * kHaloSlotDoubles + 1];
                        if (e + 1 >= n) r = xtR;
                        else if (t + 1 < T) r = in_x[t + 1] + alpha * in_d[t + 1];
                        else r = halo[kHaloSlotDoubles] + alpha * halo[3 * kHaloSlotDoubles];
                    }
                    const long long G = goff + e;
                    double ft;
                    OBJ::eval(l, xt, r, G > 0, G < nglob - 1, ft, gv);
                    sv = xt - xc; // s = x_new - x   (seq/lbfgs.cpp:177)
                    yv = gv - go; // y = g_new - g   (seq/lbfgs.cpp:178)
                    x_new[e] = xt;
                    g_out[e] = gv;
                    s_out[e] = sv;
                    y_out[e] = yv;
                    facc += ft;
                }
                cur[(size_t)r0 * T + t] = sv; // elements beyond n contribute zeros to the inner products
                cur[(size_t)r1 * T + t] = yv;
                cur[(size_t)r2 * T + t] = gv;
            }
            named_barrier_sync(1, 32 * kWsConsumerWarps);
            // ---------------- phase 2: pass A on the tile (as k_gram_tma2d) ----------------
            const double2 *cur2 = reinterpret_cast<const double2 *>(cur);
            const int e_end = (eg + 1) * slice2;
            for (int e = eg * slice2 + lane; e < e_end; e += 32) {
                const double2 a0 = cur2[r0 * T2 + e], a1 = cur2[r1 * T2 + e], a2 = cur2[r2 * T2 + e];
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                    const int j = cg + c * NG;
                    if (j < J) {
                        const double2 v = cur2[j * T2 + e];
                        acc[c][0] = fma(a0.y, v.y, fma(a0.x, v.x, acc[c][0]));
                        acc[c][1] = fma(a1.y, v.y, fma(a1.x, v.x, acc[c][1]));
                        acc[c][2] = fma(a2.y, v.y, fma(a2.x, v.x, acc[c][2]));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }
    }
    __syncthreads();
    double *red = tile; // [NE][J*3]
    if (warp > 0) {
#pragma unroll
        for (int c = 0; c < CW; ++c) {
            const int j = cg + c * NG;
            if (j < J) {
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const double s = warp_sum(acc[c][r]);
                    if (lane == 0) red[eg * (J * 3) + j * 3 + r] = s;
                }
            }
        }
        const double fw = warp_sum(facc);
        if (lane == 0) fsum[cw] = fw;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < J * 3; q += kWsThreads) {
        double s = 0.0;
        for (int g = 0; g < NE; ++g) s += red[g * (J * 3) + q];
        st->partials[(size_t)q * gridDim.x + blockIdx.x] = s;
    }
    if (threadIdx.x == 0) { // partial of f: row 3J
        double s = 0.0;
        for (int w = 0; w < kWsConsumerWarps; ++w) s += fsum[w];
        st->partials[(size_t)(3 * J) * gridDim.x + blockIdx.x] = s;
    }
}

// ---- pass B + first trial ------------------------------------------------------------------------
// CTA tiles of kCtTile double2 items that OVERLAP by 2 items (4 elements) on each side: every CTA forms d for its
// whole tile with the same fma chain (so the overlap is bit-identical in both tiles), stores only the items it
// owns, and has the neighbours' trial values in shared memory for the three-point stencil without a second,
// dependent pass or special edge threads.  Redundant loads: 4 of 512 items (0.8 %, L2 hits).
constexpr int kCtItems = 2 * kThreads;       // 512 double2 items loaded per tile
constexpr int kCtHalo = 2;                   // items of overlap on each side
constexpr int kCtOwn = kCtItems - 2 * kCtHalo; // 508 items owned

template <class OBJ>
__global__ void __launch_bounds__(kThreads, kCombineCtasPerSm) k_combine_trial(const DevState *__restrict__ st)
{
    if (st->ctrl.done) return;
    __shared__ const double *cols[kMaxCols];
    __shared__ double coef[kMaxCols];
    __shared__ __align__(16) double xt_s[2][2 * kCtItems];
    const int h = st->h;
    const bool steep = st->steepest || h == 0;
    const int J = steep ? 1 : 2 * h + 1;
    for (int j = threadIdx.x; j < J; j += kThreads) {
        cols[j] = steep ? st->g : basis_col(st, j, h); // d = -g is the combination with the single coefficient 1
        coef[j] = steep ? 1.0 : st->delta[j];
    }
    __syncthreads();
    const long long n = st->n;
    const long long nvec_pad = (n + 1) >> 1; // rows are zero-padded to a multiple of 32 doubles
    const double alpha = st->lsp.step0;      // every search starts at INITIAL_STEP_SIZE
    const double xtL = st->xL + alpha * st->dL, xtR = st->xR + alpha * st->dR;
    const long long goff = st->goff, nglob = st->nglob;
    const double *__restrict__ x = st->x;
    double *__restrict__ w = st->w;
    double a_gd = 0.0, a_f = 0.0, a_gdt = 0.0;
    const long long ntiles = (nvec_pad + kCtOwn - 1) / kCtOwn;
    int buf = 0;
    for (long long tl = blockIdx.x; tl < ntiles; tl += gridDim.x, buf ^= 1) {
        const long long first = tl * kCtOwn - kCtHalo; // first item loaded by this tile (may be -2)
        const long long i0 = first + threadIdx.x, i1 = i0 + kThreads;
        const bool ok0 = i0 >= 0 && i0 < nvec_pad, ok1 = i1 < nvec_pad;
        double2 s0 = make_double2(0.0, 0.0), s1 = s0, g0 = s0, g1 = s0;
#pragma unroll 8
        for (int j = 0; j < J; ++j) {
            const double c = coef[j];
            const double2 v0 = ok0 ? ld2(cols[j], i0) : make_double2(0.0, 0.0);
            const double2 v1 = ok1 ? ld2(cols[j], i1) : make_double2(0.0, 0.0);
            s0.x = fma(c, v0.x, s0.x); s0.y = fma(c, v0.y, s0.y);
            s1.x = fma(c, v1.x, s1.x); s1.y = fma(c, v1.y, s1.y);
            if (j == J - 1) { g0 = v0; g1 = v1; }
        }
        const double2 x0 = ok0 ? ld2(x, i0) : make_double2(0.0, 0.0);
        const double2 x1 = ok1 ? ld2(x, i1) : make_double2(0.0, 0.0);
        s0.x = -s0.x; s0.y = -s0.y; s1.x = -s1.x; s1.y = -s1.y;
        const bool own0 = ok0 && threadIdx.x >= kCtHalo, own1 = ok1 && threadIdx.x < kThreads - kCtHalo;
        if (own0) {
            st2(w, i0, s0);
            a_gd += g0.x * s0.x + g0.y * s0.y;
        }
        if (own1) {
            st2(w, i1, s1);
            a_gd += g1.x * s1.x + g1.y * s1.y;
        }
        // trial values x + alpha d of the whole tile (overlap included) for the stencil
        double *xs = xt_s[buf];
        reinterpret_cast<double2 *>(xs)[threadIdx.x] = make_double2(x0.x + alpha * s0.x, x0.y + alpha * s0.y);
        reinterpret_cast<double2 *>(xs)[threadIdx.x + kThreads] = make_double2(x1.x + alpha * s1.x, x1.y + alpha * s1.y);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const bool own = u ? own1 : own0;
            if (!own) continue;
            const long long i = u ? i1 : i0;
            const int li = 2 * (threadIdx.x + u * kThreads); // local element index of the item's first element
            const double2 dd = u ? s1 : s0;
            const long long e = 2 * i;
            const double c0 = xs[li], c1 = xs[li + 1];
            double l = 0.0, r = 0.0;
            if (OBJ::kStencil) {
                l = (e == 0) ? xtL : xs[li - 1];
                r = (e + 2 >= n) ? xtR : xs[li + 2];
            }
            const long long G0 = goff + e;
            double f0, q0, f1, q1;
            OBJ::eval(l, c0, c1, G0 > 0, true, f0, q0);
            if (e + 1 < n) {
                OBJ::eval(c0, c1, r, true, (G0 + 1) < nglob - 1, f1, q1);
            } else { // odd n: the item's second element is padding; element e is the last one and its right
                     // neighbour is the next shard's first element (or does not exist)
                OBJ::eval(l, c0, xtR, G0 > 0, G0 < nglob - 1, f0, q0);
                f1 = 0.0;
                q1 = 0.0;
            }
            a_f += f0 + f1;
            a_gdt += q0 * dd.x + q1 * dd.y;
        }
    }
    double v[3] = {a_gd, a_f, a_gdt};
    block_emit<3>(v, st->partials);
}

typedef void (*accept_gram_kernel_t)(const DevState *, const ArenaMaps *, int, int, int);
typedef void (*combine_trial_kernel_t)(const DevState *);

template <int CW>
inline accept_gram_kernel_t accept_gram_kernel_for(int objective)
{
    switch (objective) {
    case LBFGSB200_OBJ_QUADRATIC: return k_accept_gram<ObjQuadratic, CW>;
    case LBFGSB200_OBJ_ROSENBROCK: return k_accept_gram<ObjRosenbrock, CW>;
    default: return k_accept_gram<ObjTridiag, CW>;
    }
}
inline combine_trial_kernel_t combine_trial_kernel_for(int objective)
{
    switch (objective) {
    case LBFGSB200_OBJ_QUADRATIC: return k_combine_trial<ObjQuadratic>;
    case LBFGSB200_OBJ_ROSENBROCK: return k_combine_trial<ObjRosenbrock>;
    default: return k_combine_trial<ObjTridiag>;
    }
}

} // namespace lb
