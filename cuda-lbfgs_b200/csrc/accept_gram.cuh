// accept_gram.cuh -- the two streaming kernels of the fused compact flow (+ k_trial for further trials).
//
//   k_accept_gram    accept step at the alpha the line search returned (x_new, g_new, s, y, f) FUSED with pass A
//                    of the NEXT direction (the 3 x (2h'+1) inner products of s_new, y_new, g_new with the new
//                    window).  The unfused sequence writes s, y, g_new in k_accept and reads them straight back
//                    in pass A; here they go from registers into the shared-memory tile the dot products run on:
//                    reads  2(h'-1) + 3  (kept history, x, d, g_old)     writes 4 (x_new, g_new, s, y)
//                    instead of 3 + 4 (accept) + 2h' + 1 (pass A)  =>  3 vector streams fewer per iteration.
//                    Replaces updateSolution + host grad + updateVectors + 2 D2D copies + ddot(g,g)
//                    (par/L-BFGS.cu:309-347) and all 3h+3 cublasDdot calls of the next iteration (:219-267).
//   k_combine_trial  pass B (d = -sum_j delta_j b_j, g.d) FUSED with the first line-search trial: every search starts at
//                    alpha = INITIAL_STEP_SIZE (seq/line_search.cpp:21, :72, :138) and the neighbours' boundary d is known
//                    before this pass (DevState::bL / bR), so f and grad f . d at x + step0 d cost one extra read of x
//                    instead of a 2-stream trial pass, a launch and a scalar kernel.  Replaces 2h cublasDaxpy +
//                    scaleByRho + negateVector (par/L-BFGS.cu:233-276) and the first updateSolution + host f/grad + ddot
//                    of the search (par/L-BFGS-Wolfe.cu:276-311).
//
// HBM-bound streams; tensor cores are not used (FP64, ~1 flop/B).  Both are warp-specialised, one CTA per SM, with a ring
// of NS <= kMaxStages shared-memory stages filled by tensor-map TMA (cp.async.bulk.tensor.2d, SASS UTMALDG.2D) and
// mbarrier transaction counting.  Two things measured on the B200 shape them:
//   * ONE thread issuing all boxes of a tile back to back is the bottleneck (~200 cycles per tensor-map TMA instruction):
//     the boxes of a tile are issued by different LANES of the producer warp in the same cycle;
//   * a stage of the ring is held for memory latency + the work of every warp role that touches it: the roles are
//     pipelined over tiles (accept of tile k+1 overlaps the inner products of tile k), and k_combine_trial lets two
//     consumer groups work on alternating tiles.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "compact.cuh"
#include "kernels.cuh"
#include "state.h"

namespace lb {

// stage / phase of a ring of NS mbarrier-guarded stages, advanced without integer division
struct PipeState {
    int stage;
    unsigned phase;
    __device__ __forceinline__ void advance(int by, int NS)
    {
        stage += by;
        while (stage >= NS) {
            stage -= NS;
            phase ^= 1u;
        }
    }
};
__device__ __forceinline__ PipeState pipe_at(int first, int NS)
{
    PipeState p = {0, 0u};
    p.advance(first, NS);
    return p;
}

__device__ __forceinline__ void named_barrier_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(threads) : "memory");
}

// ---- accept + pass A ------------------------------------------------------------------------------
//   warp 0        producer: <= 7 tensor-map TMA loads per tile (kept S / Y rows in <= 2 runs each; {x, g_old, d} as ONE
//                 three-row box -- the arena keeps them adjacent in either parity of the iterate; two halo boxes of
//                 2 columns x 4 rows), one per lane, into the next free stage; completion is counted in bytes on
//                 full[stage].  The number of boxes matters: ~75 ns per UTMALDG and SM (11 boxes: 0.83 us per tile
//                 whatever the history size -- the kernel was box-bound below h = 8)
//   warps 1..4    accept (kAgGroups groups of 4 on alternating tiles; one group by default): wait full[stage]; one double2 item per lane: x + alpha d,
//                 the three-point stencil by warp shuffles (warp-edge lanes read the tile, tile-edge lanes the halo
//                 boxes), s, y, g_new, f; the x, d, g_old rows of the stage
//                 are OVERWRITTEN IN PLACE with s, y, g_new, the four vectors are stored to HBM, ready[stage] is raised.
//                 The groups run ahead of the gram warps by up to the pipeline depth.
//   then 12       gram warps: wait ready[stage]; rows (s, y, g_new) x all columns of the new window, exactly the loop of
//                 k_gram_tma2d; release the stage on empty[stage]
// No CTA-wide barrier in the tile loop; the four warps of an accept group meet in one named barrier per tile (between
// the reads of their neighbours' x / d and the in-place writes).
constexpr int kAgGroupWarps = 4;                  // accept warps per group (4 x 32 lanes = the 128 items of a 256-wide tile)
#ifndef LB_AG_GROUPS
#define LB_AG_GROUPS 1
#endif
// accept groups working on alternating tiles (build-time tuning knob).  Measured on B200 at n = 1e8: 1 group 2.88-2.95 ms,
// 2 groups 3.12-3.19 ms at m = 10; no difference at m = 3 -- the kernel is bound by the depth of the stage ring (a stage
// is held for memory latency + accept + inner products), not by the accept rate, and extra warps only cost issue slots.
// (With more than one group the stage count must be a multiple of the group count: a stage shared by two groups can be
// tested one phase early by the group that runs ahead -- see k_combine_trial, which gives every group its own stages.)
constexpr int kAgGroups = LB_AG_GROUPS;
constexpr int kAgAcceptWarps = kAgGroupWarps * kAgGroups;
constexpr int kAgGramWarps = 12;
constexpr int kAgThreads = 32 * (1 + kAgAcceptWarps + kAgGramWarps); // 672
constexpr int kAgMaxCW = 9; // columns per gram warp: at most ceil((2*50+1) / 12)

// Stage layout (doubles), T = tile width, hk = pairs kept from the old window (h if the ring is not full, else h-1:
// the oldest pair is evicted by the commit this kernel anticipates; a pair the curvature gate then rejects is
// handled by the scalar kernel: columns of s_new / y_new are ignored, and with a full ring the stand-alone pass A
// re-computes the g row):
//   rows 0 .. hk-1        kept S rows, window order          (TMA, <= 2 boxes)
//   rows hk .. 2hk-1      kept Y rows                        (TMA, <= 2 boxes)
//   rows 2hk, +1, +2      {x, g_old, d} or {g_old, d, x} (TMA, one box)  ->  s_new, y_new, g_new after the accept warps
//   then 2 halo slots     rows {x_a, g, w, x_b} x columns [i0-2, i0-1] and [i0+T, i0+T+1]   (TMA, 2-column boxes; a
//                         plain global load of these elements would put a DRAM round trip under load -- longer than
//                         a whole tile takes -- on the accept warps' critical path)
constexpr int kHaloSlotDoubles = 16; // one 128-byte aligned slot per halo box (TMA destinations are 128-byte aligned)
__host__ __device__ inline size_t accept_gram_stage_doubles(int J, int T)
{
    return (size_t)J * T + 4 * kHaloSlotDoubles; // J = 2m+1 >= 2hk+3 rows
}

template <class OBJ, int CW>
__global__ void __launch_bounds__(kAgThreads, 1)
k_accept_gram(const DevState *__restrict__ st, const ArenaMaps *__restrict__ maps, int T, int NS, int init_and_boxes)
{
    if (st->ctrl.done) return;
    const int init = init_and_boxes & 1;
    const bool halo_merged = (init_and_boxes & 2) != 0; // two halo boxes of 4 rows instead of four of 1 row
    const bool xgd_merged = (init_and_boxes & 4) != 0;  // {x, g, d} as one 3-row box instead of three boxes
    extern __shared__ __align__(128) double tile[];
    __shared__ __align__(8) unsigned long long full[kMaxStages], ready[kMaxStages], empty[kMaxStages];
    __shared__ double fsum[kAgAcceptWarps];
    const int h_old = init ? 0 : st->h;
    const int ks = (h_old == st->m) ? 1 : 0; // oldest pair evicted by the anticipated commit
    const int hk = h_old - ks, hp = hk + 1, J = 2 * hp + 1; // J columns of the new window
    const int nrow = 2 * hk + 3;                            // rows of a stage
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&ready[s], kAgGroupWarps);
            mbar_init(&empty[s], kAgGramWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n = st->n;
    const long long ntiles = (n + T - 1) / T;
    const int mrow = 2 * st->m + 1;
    const size_t stage_doubles = accept_gram_stage_doubles(mrow, T);
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long tile_stride = (long long)gridDim.x * T; // elements between consecutive tiles of this CTA
    const int T2 = T >> 1;
    const int NE = T2 / 32 > 0 ? T2 / 32 : 1;     // element groups of 32 items (T >= 64)
    const int NG = kAgGramWarps / NE;             // column groups: 3, 4, 6 or 12
    double acc[CW][3];
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.0;
    double facc = 0.0;
    const int gw = warp - 1 - kAgAcceptWarps, cg = gw % NG, eg = gw / NG; // gram-warp coordinates (warp >= 9)
    const bool x_is_a = st->x == st->arena0; // parity of the iterate: rows {x_a, g, w} or {g, w, x_b} hold {x, g, d}

    if (warp == 0) {
        // ---------------- producer ----------------
        //   lane 0: S run A   1: Y run A   2: S run B   3: Y run B
        //   lane 4: {x, g_old, d} -- three consecutive arena rows in either parity of the iterate -- as ONE box, or
        //           lanes 4, 5, 6: x, g_old, d as three boxes
        //   lanes 7, 8: left / right halo as boxes of 2 columns x the 4 rows x_a, g, w, x_b, or
        //           lanes 7 .. 10: x left, x right, d left, d right as four one-row boxes
        //   (which of the two is faster depends on the history size: fewer boxes = less issue time, but the 4-row halo
        //   boxes touch twice as many DRAM pages; the host picks per m, LBFGSB200_AG_BOXES overrides)
        const int ns = st->nslots;
        const int a0 = (st->base + ks) % ns;          // physical slot of the first kept pair
        const int ra = min(hk, ns - a0), rb = hk - ra; // the kept window is at most two runs of slots
        const CUtensorMap *map = &maps->run[1];
        int row = 0;
        size_t off = 0; // destination inside the stage, in doubles
        int dcol = 0;   // column offset of the box relative to the tile
        bool valid = true;
        switch (lane) {
        case 0: map = &maps->run[ra]; row = kArenaRowS + a0; off = 0; valid = ra > 0; break;
        case 1: map = &maps->run[ra]; row = kArenaRowS + ns + a0; off = (size_t)hk * T; valid = ra > 0; break;
        case 2: map = &maps->run[rb]; row = kArenaRowS; off = (size_t)ra * T; valid = rb > 0; break;
        case 3: map = &maps->run[rb]; row = kArenaRowS + ns; off = (size_t)(hk + ra) * T; valid = rb > 0; break;
        // {x, g, d}: rows {x_a, g, w} or {g, w, x_b} of the arena land in that order in stage rows 2hk .. 2hk+2
        case 4:
            if (xgd_merged) { map = &maps->run[3]; row = x_is_a ? kArenaRowXa : kArenaRowG; off = (size_t)(2 * hk) * T; }
            else { row = x_is_a ? kArenaRowXa : kArenaRowXb; off = (size_t)(2 * hk + (x_is_a ? 0 : 2)) * T; }
            break;
        case 5: row = kArenaRowG; off = (size_t)(2 * hk + (x_is_a ? 1 : 0)) * T; valid = !xgd_merged; break;
        case 6: row = kArenaRowW; off = (size_t)(2 * hk + (x_is_a ? 2 : 1)) * T; valid = !xgd_merged; break;
        // halo: merged = [row x_a, g, w, x_b][2 columns] left / right; else one slot each: x left, x right, d left, d right
        case 7: map = halo_merged ? &maps->halo : &maps->halo1; row = halo_merged ? 0 : (x_is_a ? kArenaRowXa : kArenaRowXb);
                off = (size_t)mrow * T; dcol = -2; break;
        case 8: map = halo_merged ? &maps->halo : &maps->halo1; row = halo_merged ? 0 : (x_is_a ? kArenaRowXa : kArenaRowXb);
                off = (size_t)mrow * T + kHaloSlotDoubles; dcol = T; break;
        case 9: map = &maps->halo1; row = kArenaRowW; off = (size_t)mrow * T + 2 * kHaloSlotDoubles; dcol = -2; valid = !halo_merged; break;
        case 10: map = &maps->halo1; row = kArenaRowW; off = (size_t)mrow * T + 3 * kHaloSlotDoubles; dcol = T; valid = !halo_merged; break;
        default: valid = false; break;
        }
        const unsigned bytes = (unsigned)(((size_t)nrow * T + (halo_merged ? 2 * 8 : 4 * 2)) * sizeof(double));
        PipeState ps = {0, 1u}; // the producer waits for the PREVIOUS use of a stage to be released
        long long col = (long long)blockIdx.x * T;
        for (long long k = 0; k < my_tiles; ++k, col += tile_stride, ps.advance(1, NS)) {
            if (lane == 0) {
                if (k >= NS) mbar_wait(&empty[ps.stage], ps.phase);
                mbar_expect_tx(&full[ps.stage], bytes);
            }
            __syncwarp();
            if (valid) tma_load_2d(tile + ps.stage * stage_doubles + off, map, (int)col + dcol, row, &full[ps.stage]);
        }
    } else if (warp <= kAgAcceptWarps) {
        // ---------------- accept warps: group grp handles tiles grp, grp + 2, ... ----------------
        const int grp = (warp - 1) / kAgGroupWarps, aw = (warp - 1) % kAgGroupWarps;
        const double alpha = init ? 0.0 : st->ls.alpha;
        const double xtL = st->xL + alpha * st->dL, xtR = st->xR + alpha * st->dR;
        const long long goff = st->goff, nglob = st->nglob;
        double *__restrict__ x_new = st->x_alt;
        double *__restrict__ g_out = st->g;
        const size_t sp = (size_t)spare_slot(*st) * (size_t)st->stride;
        double *__restrict__ s_out = st->S + sp;
        double *__restrict__ y_out = st->Y + sp;
        const int e = aw * 32 + lane;   // double2 item of this lane within the tile
        const bool act = e < T2;
        // where the elements just outside the tile are, relative to the halo area of the stage
        const int hrx = 2 * (x_is_a ? kArenaRowXa : kArenaRowXb), hrd = 2 * kArenaRowW;
        const int hxl = halo_merged ? hrx + 1 : 1, hdl = halo_merged ? hrd + 1 : 2 * kHaloSlotDoubles + 1;
        const int hxr = halo_merged ? kHaloSlotDoubles + hrx : kHaloSlotDoubles, hdr = halo_merged ? kHaloSlotDoubles + hrd : 3 * kHaloSlotDoubles;
        long long col = ((long long)blockIdx.x + grp * (long long)gridDim.x) * T; // first element of the group's first tile
        PipeState ps = pipe_at(grp, NS);
        for (long long k = grp; k < my_tiles; k += kAgGroups, col += kAgGroups * tile_stride, ps.advance(kAgGroups, NS)) {
            mbar_wait(&full[ps.stage], ps.phase);
            double *cur = tile + ps.stage * stage_doubles;
            double *out_s = cur + (size_t)(2 * hk) * T, *out_y = out_s + T, *out_g = out_y + T; // the three new rows, in place
            const double *in_x = x_is_a ? out_s : out_g, *in_g = x_is_a ? out_y : out_s, *in_d = x_is_a ? out_g : out_y;
            // halo boxes: [row x_a, g, w, x_b][2 columns]; left = elements col-2, col-1, right = col+T, col+T+1
            const double *halo = cur + (size_t)mrow * T;
            const long long ge = col + 2 * (long long)e; // first element of the item
            double2 xc = make_double2(0.0, 0.0), dc = xc, go = xc;
            if (act) {
                xc = reinterpret_cast<const double2 *>(in_x)[e];
                dc = reinterpret_cast<const double2 *>(in_d)[e];
                if (!init) go = reinterpret_cast<const double2 *>(in_g)[e]; // (the arena is not cleared: the x0 evaluation must not read old bits)
            }
            const double xt0 = xc.x + alpha * dc.x; // add(x, scalarProduct(alpha, d)): mul, then add
            const double xt1 = xc.y + alpha * dc.y;
            double l = 0.0, r = 0.0;
            if (OBJ::kStencil) {
                l = __shfl_up_sync(0xffffffffu, xt1, 1);
                r = __shfl_down_sync(0xffffffffu, xt0, 1);
                if (act) {
                    if (lane == 0) {
                        const int tl = 2 * e - 1; // tile coordinate of the left neighbour
                        if (ge == 0) l = xtL;
                        else if (tl >= 0) l = in_x[tl] + alpha * in_d[tl];
                        else l = halo[hxl] + alpha * halo[hdl];
                    }
                    if (ge + 2 >= n) r = xtR;
                    else if (lane == 31) {
                        const int tr = 2 * e + 2;
                        if (tr < T) r = in_x[tr] + alpha * in_d[tr];
                        else r = halo[hxr] + alpha * halo[hdr];
                    }
                }
            }
            // every warp of the group has read what it needs of its neighbours' x / d: the rows may now be overwritten
            named_barrier_sync(1 + grp, 32 * kAgGroupWarps);
            if (act) {
                double2 a0 = make_double2(0.0, 0.0), a1 = a0, a2 = a0; // elements beyond n contribute zeros
                if (ge < n) {
                    const long long G0 = goff + ge;
                    double f0, g0, f1 = 0.0, g1 = 0.0;
                    const bool two = ge + 1 < n;
                    if (two) {
                        OBJ::eval(l, xt0, xt1, G0 > 0, true, f0, g0);
                        OBJ::eval(xt0, xt1, r, true, (G0 + 1) < nglob - 1, f1, g1);
                    } else { // odd n: the item's second element is padding
                        OBJ::eval(l, xt0, xtR, G0 > 0, G0 < nglob - 1, f0, g0);
                    }
                    a0.x = xt0 - xc.x; // s = x_new - x   (seq/lbfgs.cpp:177)
                    a1.x = g0 - go.x;  // y = g_new - g   (seq/lbfgs.cpp:178)
                    a2.x = g0;
                    facc += f0 + f1;
                    const long long j2 = ge >> 1;
                    if (two) {
                        a0.y = xt1 - xc.y;
                        a1.y = g1 - go.y;
                        a2.y = g1;
                        st2(x_new, j2, make_double2(xt0, xt1));
                        st2(g_out, j2, a2);
                        st2(s_out, j2, a0);
                        st2(y_out, j2, a1);
                    } else { // never write the zero padding of the rows
                        x_new[ge] = xt0;
                        g_out[ge] = a2.x;
                        s_out[ge] = a0.x;
                        y_out[ge] = a1.x;
                    }
                }
                reinterpret_cast<double2 *>(out_s)[e] = a0; // s_new
                reinterpret_cast<double2 *>(out_y)[e] = a1; // y_new
                reinterpret_cast<double2 *>(out_g)[e] = a2; // g_new
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&ready[ps.stage]); // release: the three new rows of this stage are in place
        }
    } else {
        // ---------------- gram warps ----------------
        int rowc[CW]; // shared-memory row of window column j = cg + c NG
#pragma unroll
        for (int c = 0; c < CW; ++c) {
            const int j = cg + c * NG;
            rowc[c] = (j < hk) ? j : (j == hk) ? 2 * hk : (j <= 2 * hk) ? j - 1 : j;
        }
        const int e = eg * 32 + lane;
        const bool act = e < T2;
        PipeState ps = {0, 0u};
        for (long long k = 0; k < my_tiles; ++k, ps.advance(1, NS)) {
            mbar_wait(&ready[ps.stage], ps.phase);
            const double2 *cur2 = reinterpret_cast<const double2 *>(tile + ps.stage * stage_doubles);
            if (act) {
                const double2 a0 = cur2[(size_t)(2 * hk) * T2 + e], a1 = cur2[(size_t)(2 * hk + 1) * T2 + e],
                              a2 = cur2[(size_t)(2 * hk + 2) * T2 + e];
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                    if (cg + c * NG < J) { // warp-uniform
                        const double2 v = cur2[(size_t)rowc[c] * T2 + e];
                        acc[c][0] = fma(a0.y, v.y, fma(a0.x, v.x, acc[c][0]));
                        acc[c][1] = fma(a1.y, v.y, fma(a1.x, v.x, acc[c][1]));
                        acc[c][2] = fma(a2.y, v.y, fma(a2.x, v.x, acc[c][2]));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[ps.stage]); // this warp is done reading the stage
        }
    }
    __syncthreads(); // all tiles consumed; the tile storage can be reused for the reduction
    double *red = tile; // [NE][J*3]
    if (warp > kAgAcceptWarps) {
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
    } else if (warp > 0) {
        const double fw = warp_sum(facc);
        if (lane == 0) fsum[warp - 1] = fw;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < J * 3; q += kAgThreads) {
        double s = 0.0;
        for (int g = 0; g < NE; ++g) s += red[g * (J * 3) + q];
        st->partials[(size_t)q * gridDim.x + blockIdx.x] = s;
    }
    if (threadIdx.x == 0) { // partial of f: row 3J
        double s = 0.0;
        for (int w = 0; w < kAgAcceptWarps; ++w) s += fsum[w];
        st->partials[(size_t)(3 * J) * gridDim.x + blockIdx.x] = s;
    }
}

// ---- pass B + first trial ------------------------------------------------------------------------
// Warp 0 issues <= 6 tensor-map TMA loads per tile (S and Y in <= 2 runs each; {x, g} as one box when the iterate is
// x_a, else g and x_b), one per lane; two groups of 4
// consumer warps take alternating tiles, one double2 item per lane: the fma chain over the 2h+1 columns out of shared
// memory (window-column order: the chain k_combine uses, so d has the same bits), the store of d, the trial point, the
// stencil by warp shuffles, warp-edge values going through a small shared-memory exchange and ONE named barrier per
// tile and group.  Tiles OVERLAP by halo2 items on each side (a TMA box of T columns every T - 4 halo2 elements): the halo
// of the stencil comes with the tile, recomputed with the same chain; the overlap is requested by neighbouring CTAs at the
// same time and is served by L2.  halo2 = 4 keeps every box 64-byte aligned (measured best on B200: 2 -> 2.80 ms,
// 4 -> 2.72 ms, 8 -> 2.74 ms at n = 1e8, m = 10).
constexpr int kCtGroupWarps = 4;
#ifndef LB_CT_GROUPS
#define LB_CT_GROUPS 2
#endif
constexpr int kCtGroups = LB_CT_GROUPS; // consumer groups on alternating tiles (2: 2.75 ms, 1: 2.81 ms at n = 1e8, m = 10)
constexpr int kCtWarps = kCtGroupWarps * kCtGroups;
constexpr int kCtThreads = 32 * (kCtWarps + 1);
constexpr int kCtHaloItems = 4; // default double2 items of overlap on each side of a tile

__host__ __device__ inline size_t combine_trial_stage_doubles(int m, int T) { return (size_t)(2 * m + 2) * T; }
// stages of the ring that belong to consumer group g (physical stages g, g + kCtGroups, ... < NS)
__host__ __device__ inline int ct_group_stages(int NS, int g) { return (NS - g + kCtGroups - 1) / kCtGroups; }

template <class OBJ>
__global__ void __launch_bounds__(kCtThreads, 1)
k_combine_trial(const DevState *__restrict__ st, const ArenaMaps *__restrict__ maps, int T, int NS, int halo2)
{
    if (st->ctrl.done) return;
    extern __shared__ __align__(128) double tile[];
    __shared__ __align__(8) unsigned long long full[kMaxStages], empty[kMaxStages];
    __shared__ double coef[kMaxCols];
    __shared__ __align__(16) double2 xs[kCtGroups][2][32 * kCtGroupWarps]; // trial values of a tile, for the warp-edge lanes
    const int h = st->h;
    const bool steep = st->steepest || h == 0;
    const int J = steep ? 1 : 2 * h + 1; // d = -g is the combination with the single coefficient 1 on g
    for (int j = threadIdx.x; j < J; j += kCtThreads) coef[j] = steep ? 1.0 : st->delta[j];
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kCtGroupWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long n = st->n;
    const long long nvec_pad = (n + 1) >> 1;
    const int T2 = T >> 1, own2 = T2 - 2 * halo2; // items per tile / items owned per tile (halo2 items of overlap per side)
    const long long ntiles = (nvec_pad + own2 - 1) / own2;
    const size_t stage_doubles = combine_trial_stage_doubles(st->m, T);
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long item_stride = (long long)gridDim.x * own2; // items between consecutive tiles of this CTA
    double a_gd = 0.0, a_f = 0.0, a_gdt = 0.0;
    const bool x_is_a = st->x == st->arena0; // rows 2hh, 2hh+1 of a stage hold [x, g] (one box) or [g, x] (two boxes)

    if (warp == 0) {
        // producer: lane 0: S run A   1: S run B   2: Y run A   3: Y run B (rows in window-column order), then g and x:
        // adjacent arena rows {x_a, g} when the iterate is x_a -- ONE box, landing as [x, g] -- else lane 4: g, lane 5: x_b
        const int ns = st->nslots, b0 = st->base;
        const int ra = steep ? 0 : min(h, ns - b0), rb = steep ? 0 : h - ra;
        const int hh = steep ? 0 : h;
        const CUtensorMap *map = &maps->run[1];
        int row = 0;
        size_t off = 0;
        bool valid = true;
        switch (lane) {
        case 0: map = &maps->run[ra]; row = kArenaRowS + b0; off = 0; valid = ra > 0; break;
        case 1: map = &maps->run[rb]; row = kArenaRowS; off = (size_t)ra * T; valid = rb > 0; break;
        case 2: map = &maps->run[ra]; row = kArenaRowS + ns + b0; off = (size_t)hh * T; valid = ra > 0; break;
        case 3: map = &maps->run[rb]; row = kArenaRowS + ns; off = (size_t)(hh + ra) * T; valid = rb > 0; break;
        case 4: map = &maps->run[x_is_a ? 2 : 1]; row = x_is_a ? kArenaRowXa : kArenaRowG; off = (size_t)(2 * hh) * T; break;
        case 5: row = kArenaRowXb; off = (size_t)(2 * hh + 1) * T; valid = !x_is_a; break;
        default: valid = false; break;
        }
        const unsigned bytes = (unsigned)((size_t)(2 * hh + 2) * T * sizeof(double));
        // Every consumer group owns its OWN stages (group g: physical stages g, g + G, ...; sub-ring of ct_group_stages
        // entries): a stage is then always filled for, and released by, the same group, in that group's tile order.
        // With one shared ring and an odd stage count the two groups would alternate on a stage, and a group that runs
        // ahead could test a stage's `full` barrier one phase early -- the parity of the phase before last looks
        // "completed" -- and consume a tile that is still in flight (TMA loads complete out of order).
        PipeState ps[kCtGroups];
#pragma unroll
        for (int g = 0; g < kCtGroups; ++g) ps[g] = PipeState{0, 1u}; // (a fresh barrier passes the wait on parity 1)
        long long first = (long long)blockIdx.x * own2 - halo2; // first item of the tile (negative for the very first: zero-filled)
        int g = 0;
        for (long long k = 0; k < my_tiles; ++k, first += item_stride) {
            int stage = 0;
            unsigned phase = 0;
#pragma unroll
            for (int q = 0; q < kCtGroups; ++q) // (static indexing: the states stay in registers)
                if (q == g) {
                    stage = q + kCtGroups * ps[q].stage;
                    phase = ps[q].phase;
                    ps[q].advance(1, ct_group_stages(NS, q));
                }
            if (lane == 0) {
                mbar_wait(&empty[stage], phase);
                mbar_expect_tx(&full[stage], bytes);
            }
            __syncwarp();
            if (valid) tma_load_2d(tile + stage * stage_doubles + off, map, (int)(2 * first), row, &full[stage]);
            if (++g == kCtGroups) g = 0;
        }
    } else {
        const int grp = (warp - 1) / kCtGroupWarps, cwp = (warp - 1) % kCtGroupWarps;
        const int e = cwp * 32 + lane; // item of this lane within the tile
        const bool in_tile = e < T2;
        const bool owner = in_tile && e >= halo2 && e < T2 - halo2;
        const int hh = steep ? 0 : h;
        const double alpha = st->lsp.step0; // every search starts at INITIAL_STEP_SIZE
        const double xtL = st->xL + alpha * st->dL, xtR = st->xR + alpha * st->dR;
        const long long goff = st->goff, nglob = st->nglob;
        double *__restrict__ w = st->w;
        PipeState ps = {0, 0u}; // this group's sub-ring: physical stage = grp + kCtGroups * ps.stage
        const int my_stages = ct_group_stages(NS, grp);
        long long first = ((long long)blockIdx.x + grp * (long long)gridDim.x) * own2 - halo2;
        int buf = 0;
        for (long long k = grp; k < my_tiles; k += kCtGroups, first += kCtGroups * item_stride, ps.advance(1, my_stages), buf ^= 1) {
            const int stage = grp + kCtGroups * ps.stage;
            mbar_wait(&full[stage], ps.phase);
            const double2 *cur2 = reinterpret_cast<const double2 *>(tile + stage * stage_doubles);
            const long long i = first + e; // global item index (may be < 0 or beyond the vector: zeros)
            double2 s = make_double2(0.0, 0.0), gv = s, xv = s;
            if (in_tile) {
#pragma unroll 4
                for (int j = 0; j < J - 1; ++j) { // the S and Y columns, window-column order
                    const double c = coef[j];
                    const double2 v = cur2[(size_t)j * T2 + e];
                    s.x = fma(c, v.x, s.x);
                    s.y = fma(c, v.y, s.y);
                }
                gv = cur2[(size_t)(2 * hh + (x_is_a ? 1 : 0)) * T2 + e]; // last column: g
                xv = cur2[(size_t)(2 * hh + (x_is_a ? 0 : 1)) * T2 + e];
                const double c = coef[J - 1];
                s.x = fma(c, gv.x, s.x);
                s.y = fma(c, gv.y, s.y);
            }
            s.x = -s.x;
            s.y = -s.y;
            const bool own = owner && i < nvec_pad; // (i >= 0 for every owned item)
            if (own) {
                st2(w, i, s);
                a_gd += gv.x * s.x + gv.y * s.y;
            }
            const double2 xt = make_double2(xv.x + alpha * s.x, xv.y + alpha * s.y);
            double l = 0.0, r = 0.0;
            if (OBJ::kStencil) {
                double2 *xb = xs[grp][buf];
                xb[e] = xt;
                l = __shfl_up_sync(0xffffffffu, xt.y, 1);
                r = __shfl_down_sync(0xffffffffu, xt.x, 1);
                named_barrier_sync(1 + grp, 32 * kCtGroupWarps);
                if (lane == 0 && e > 0) l = xb[e - 1].y;
                if (lane == 31 && e + 1 < T2) r = xb[e + 1].x;
            }
            if (own) {
                const long long el = 2 * i;
                if (el == 0) l = xtL;
                if (el + 2 >= n) r = xtR;
                const long long G0 = goff + el;
                double f0, q0, f1 = 0.0, q1 = 0.0;
                if (el + 1 < n) {
                    OBJ::eval(l, xt.x, xt.y, G0 > 0, true, f0, q0);
                    OBJ::eval(xt.x, xt.y, r, true, (G0 + 1) < nglob - 1, f1, q1);
                } else { // odd n: the item's second element is padding, element el is the last of the shard
                    OBJ::eval(l, xt.x, xtR, G0 > 0, G0 < nglob - 1, f0, q0);
                }
                a_f += f0 + f1;
                a_gdt += q0 * s.x + q1 * s.y;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
        }
    }
    __syncthreads();
    // one partial per quantity and CTA (g.d, f, grad f . d at the trial point), fixed order
    __shared__ double red[3][kCtWarps];
    if (warp > 0) {
        const double v0 = warp_sum(a_gd), v1 = warp_sum(a_f), v2 = warp_sum(a_gdt);
        if (lane == 0) { red[0][warp - 1] = v0; red[1][warp - 1] = v1; red[2][warp - 1] = v2; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        double v = 0.0;
        for (int q = 0; q < kCtWarps; ++q) v += red[threadIdx.x][q];
        st->partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = v;
    }
}

typedef void (*accept_gram_kernel_t)(const DevState *, const ArenaMaps *, int, int, int);
typedef void (*combine_trial_kernel_t)(const DevState *, const ArenaMaps *, int, int, int);

// columns per gram warp for history size m and tile width T (the kernel derives the same NG from T)
inline int accept_gram_cw(int m, int T)
{
    const int NE = T / 64 > 0 ? T / 64 : 1, NG = kAgGramWarps / NE;
    return (2 * m + 1 + NG - 1) / NG;
}

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
