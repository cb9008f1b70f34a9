// compact.cuh -- the compact ("vector-free" / Gram) form of the L-BFGS direction.
//
// The explicit two-loop recursion (seq/lbfgs.cpp:93-143) streams the work vector 2h times
// because every pass needs a scalar produced by the previous one: (8h-1) vector streams.  The
// compact form keeps the recursion but runs it on COEFFICIENTS: the direction is a linear
// combination  d = -sum_j delta_j b_j  of the basis  b = [s_0..s_{h-1}, y_0..y_{h-1}, g], and every
// inner product the recursion needs (s_i.q, y_i.r with q, r in span(b)) is a delta-weighted sum
// of entries of the Gram matrix  G = b^T b.  Per iteration only the rows of G that changed are
// computed -- those of the newest pair (s_c, y_c) and of g -- all DIRECTLY from the vectors (no
// linearity tricks: y_c = g_new - g_old would cancel badly near convergence):
//
//   pass A  k_gram     reads the 2h+1 basis vectors ONCE  -> 3 x (2h+1) inner products
//   scalar  OP_COMPACT Gram update + the two loops on delta: O(h^2) flops, one thread
//   pass B  k_combine  reads the 2h+1 basis vectors once, writes d, accumulates g.d
//
// => (4h + 3) vector streams instead of (8h - 1); with trials and accept (4h + 2t + 10) V.
// This is the "batched GEMV over S^T g, Y^T g" variant of the north star: pass A is a skinny
// (3 x n)(n x (2h+1)) product, HBM-bound at ~0.75 flop/B, so it stays on the FP64 pipe.
//
// Pass A stages tiles of ALL basis vectors in shared memory with a two-stage cp.async pipeline,
// then each warp owns every 8th column and each lane every 32nd element, so the number of
// accumulators per thread is independent of h.
#pragma once
#include <cuda.h> // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#include <cuda_runtime.h>

#include "kernels.cuh"
#include "state.h"

namespace lb {

constexpr int kGramWarps = kThreads / 32;      // 8
constexpr int kMaxCompactM = 50;               // history sweep of BASELINE config 5 goes to 50
constexpr int kMaxCols = 2 * kMaxCompactM + 1; // 101
constexpr int kMaxCW = (kMaxCols + kGramWarps - 1) / kGramWarps; // 13 columns per warp
constexpr int kGramCtasPerSm = 2;
constexpr int kCombineCtasPerSm = 3; // 16 x 16-byte loads in flight per thread

// column j of the basis in window order: S positions 0..h-1, Y positions 0..h-1, g
__device__ __forceinline__ const double *basis_col(const DevState *st, int j, int h)
{
    if (j < h) return st->S + (size_t)slot_of(*st, j) * (size_t)st->stride;
    if (j < 2 * h) return st->Y + (size_t)slot_of(*st, j - h) * (size_t)st->stride;
    return st->g;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src, int src_bytes)
{
    // LDGSTS: 16-byte global -> shared copy that bypasses registers; bytes beyond src_bytes are
    // zero-filled (ragged last tile).
    const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gmem_src), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// pass A.  partials[(j*3 + r) * gridDim.x + cta] = sum over this CTA's tiles of row_r . col_j,
// rows r = 0: s_newest, 1: y_newest, 2: g.
//
// Multi-stage cp.async pipeline: while the CTA reduces tile k out of shared memory, the 16-byte
// async copies of tiles k+1 .. k+NS-1 (all 2h+1 vectors x T elements each) are already in
// flight, so whole tiles per CTA -- not a handful of registers per thread -- are outstanding.
// Warp w owns columns w, w+8, ...; lane l owns elements l, l+32, ...: CW*3 accumulators per
// thread, independent of h.  Products are fused (fma): inner products carry no bitwise contract.
// Rows beyond n: arena rows are 256-byte padded with zeros that are never written, and whole
// items beyond the padded length are zero-filled by the copy itself.
template <int CW>
__global__ void __launch_bounds__(kThreads, kGramCtasPerSm)
k_gram(const DevState *__restrict__ st, int T, int NS, int G)
{
    const int h = st->h;
    if (st->ctrl.done || h == 0 || (st->steepest && !st->sg_valid)) return; // rows of a fresh pair are needed even if d = -g
    extern __shared__ __align__(128) double tile[]; // [NS][Jt][T]
    __shared__ const double *cols[kMaxCols];
    const int J = 2 * h + 1;
    // column group of this CTA (blockIdx.y): with G > 1 the basis is split into G groups so that a
    // tile of one group (+ the three row vectors) still fits a large T; the rows are then re-read
    // once per group (+3/(J/G) traffic).  G == 1: the rows are simply three of the columns.
    const int per = (J + G - 1) / G;
    const int c0 = blockIdx.y * per;
    const int Jg = min(J, c0 + per) - c0;
    if (Jg <= 0) return;
    const int Jt = (G == 1) ? J : Jg + 3;
    const int rowcol[3] = {h - 1, 2 * h - 1, 2 * h}; // s_newest, y_newest, g
    const long long n = st->n;
    for (int j = threadIdx.x; j < Jt; j += kThreads)
        cols[j] = basis_col(st, j < Jg ? c0 + j : rowcol[j - Jg], h);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc[CW][3];
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.0;
    const int r0 = (G == 1) ? rowcol[0] : Jg, r1 = (G == 1) ? rowcol[1] : Jg + 1, r2 = (G == 1) ? rowcol[2] : Jg + 2;
    const int T2 = T >> 1; // double2 items per vector per tile (a power of two)
    const int log2T2 = 31 - __clz(T2);
    const long long nvec_pad = (n + 1) >> 1;
    const long long ntiles = (n + T - 1) / T;
    const int total = Jt * T2;
    const size_t stage_doubles = (size_t)Jt * T;

    auto issue = [&](long long t, int stage) {
        const long long base2 = t * T2;
        double2 *dst = reinterpret_cast<double2 *>(tile + stage * stage_doubles);
        for (int idx = threadIdx.x; idx < total; idx += kThreads) {
            const int j = idx >> log2T2, i = idx & (T2 - 1);
            const bool ok = base2 + i < nvec_pad;
            const double2 *src = reinterpret_cast<const double2 *>(cols[j]) + (ok ? base2 + i : 0);
            cp_async16(dst + idx, src, ok ? 16 : 0);
        }
        cp_async_commit();
    };

    // NS-stage pipeline: NS-1 tiles are in flight while one is being reduced.  Exactly one group is
    // committed per prologue slot and per iteration (an empty one when there is nothing left to
    // fetch), so "all but the newest NS-1 groups have landed" always means "tile k has landed".
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    for (int k = 0; k < NS - 1; ++k) {
        if (k < my_tiles) issue(blockIdx.x + (long long)k * gridDim.x, k);
        else cp_async_commit();
    }
    int stage = 0;
    for (long long k = 0; k < my_tiles; ++k, stage = (stage + 1 == NS ? 0 : stage + 1)) {
        const long long kn = k + NS - 1;
        if (kn < my_tiles) issue(blockIdx.x + kn * gridDim.x, (int)(kn % NS));
        else cp_async_commit();
        switch (NS) {
        case 2: cp_async_wait<1>(); break;
        case 3: cp_async_wait<2>(); break;
        default: cp_async_wait<3>(); break;
        }
        __syncthreads();
        const double *cur = tile + stage * stage_doubles;
        for (int e = lane; e < T; e += 32) {
            const double a0 = cur[r0 * T + e], a1 = cur[r1 * T + e], a2 = cur[r2 * T + e];
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                const int j = warp + c * kGramWarps;
                if (j < Jg) {
                    const double v = cur[j * T + e];
                    acc[c][0] = fma(a0, v, acc[c][0]);
                    acc[c][1] = fma(a1, v, acc[c][1]);
                    acc[c][2] = fma(a2, v, acc[c][2]);
                }
            }
        }
        __syncthreads(); // the next iteration's copies overwrite this stage
    }
#pragma unroll
    for (int c = 0; c < CW; ++c) {
        const int j = warp + c * kGramWarps;
        if (j < Jg) { // warp-uniform
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const double s = warp_sum(acc[c][r]);
                if (lane == 0) st->partials[(size_t)((c0 + j) * 3 + r) * gridDim.x + blockIdx.x] = s;
            }
        }
    }
}

// ---- pass A, TMA variants ------------------------------------------------------------------
// Same arithmetic, but the tiles are moved by the copy engine (tensor-map TMA, below): completion is
// counted in bytes on an mbarrier per stage, and NS stages are kept in flight.  No per-thread
// copy instructions (the LDGSTS version spends ~40% of its MIO slots issuing copies), shared memory is
// read back with 128-bit loads, and the consumer warps are split into NG column groups x 16/NG
// element groups so that each row value is re-read NG times.
constexpr int kMaxStages = 8;  // upper bound of the shared-memory ring depth (the kernels take the actual count NS <= kMaxStages)
constexpr int kGramStages = 4; // ring depth of the stand-alone pass A

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    for (int spin = 0; !done; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (spin > (1 << 24)) __trap(); // a lost copy must fail the launch, never hang the GPU
    }
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised pass A: ONE CTA per SM owning ~200 KB of shared memory.
//   warp 0            producer: per tile, <= 5 tiled TMA loads into the next free stage; completion is
//                     counted in bytes on full[stage]
//   warps 1..16       consumers: wait on full[stage], reduce their (column group, element group) share
//                     of the tile with 128-bit shared-memory loads, then release the stage by arriving
//                     on empty[stage] -- no CTA-wide barrier inside the loop
// NS-1 whole tiles (all 2h+1 vectors) per SM are in flight while one is being reduced.
constexpr int kWsConsumerWarps = 16;
constexpr int kWsThreads = 32 * (kWsConsumerWarps + 1);

// ---- pass A, tensor-map TMA variant ---------------------------------------------------------
// The (s, y) ring buffers are 2-D tensors [slot][element] with a fixed row stride, and the window
// of stored pairs is a circular range of slots, i.e. at most two runs of consecutive rows.  So a
// whole tile of the history -- up to m rows x T elements per array -- is fetched with at most FIVE
// tiled TMA loads (cp.async.bulk.tensor.2d, SASS UTMALDG): S run A, S run B, Y run A, Y run B, g,
// instead of 2h+1 one-row bulk copies.  Tensor maps are built on the host per possible run length
// (the box shape is part of the map) and live in global memory.  Out-of-range columns of the last
// tile are zero-filled by the TMA unit, and the transaction count is always the full box.
// Producer / consumer structure, tile layout and arithmetic are those of k_gram_tma.
// Tensor maps over the WHOLE arena, a row-major [4 + 2 nslots][stride] FP64 tensor (rows: x_a, g, w, x_b,
// S slots, Y slots): run[r] = box of T columns x r rows (r consecutive ring slots; r = 3: {x_a, g, w} or {g, w, x_b};
// r = 2: {x_a, g}), halo = box of 2 columns x 4 rows (the elements of x_a, g, w, x_b just outside a tile,
// accept_gram.cuh).  Out-of-range columns (negative or >= stride) are zero-filled by the TMA unit and still count for
// the full box in the transaction bytes.  Built on the host per solver (cuTensorMapEncodeTiled), kept in global memory.
//
// Row order: the iterate ping-pongs between x_a and x_b (the accept step writes x_new to the one that is not x), with
// g and w (= d) between them, so that {x, g, d} are THREE CONSECUTIVE ROWS in either parity -- one TMA box instead
// of three.  The fused kernels are bound by the number of boxes per tile at small histories (~75 ns per UTMALDG
// and SM, measured: k_accept_gram took 0.83 us per tile with 11 boxes whatever the history size).
struct ArenaMaps {
    CUtensorMap run[kMaxSlots + 1];
    CUtensorMap halo;  // 2 columns x 4 rows
    CUtensorMap halo1; // 2 columns x 1 row
};
constexpr int kArenaRowXa = 0, kArenaRowG = 1, kArenaRowW = 2, kArenaRowXb = 3, kArenaRowS = 4;

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, int c0, int c1,
                                            unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}

template <int CW>
__global__ void __launch_bounds__(kWsThreads, 1)
k_gram_tma2d(const DevState *__restrict__ st, const ArenaMaps *__restrict__ maps, int T, int NG, int NS)
{
    const int h = st->h;
    if (st->ctrl.done || h == 0 || (st->steepest && !st->sg_valid)) return;
    extern __shared__ __align__(128) double tile[]; // [NS][J][T]
    __shared__ __align__(8) unsigned long long full[kMaxStages], empty[kMaxStages];
    const int J = 2 * h + 1;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], kWsConsumerWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long ntiles = (st->n + T - 1) / T;
    const size_t stage_doubles = (size_t)J * T;
    const long long my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int NE = kWsConsumerWarps / NG;
    double acc[CW][3];
#pragma unroll
    for (int c = 0; c < CW; ++c) acc[c][0] = acc[c][1] = acc[c][2] = 0.0;
    const int cw = warp - 1, cg = cw % NG, eg = cw / NG;

    if (warp == 0) {
        // window = slots base .. base+h-1 (mod nslots): run A = [base, base+ra), run B = [0, rb).  One box per lane
        // (a tensor-map TMA instruction costs ~200 cycles to issue; five in a row from one thread starve the ring):
        //   lane 0: S run A   1: S run B   2: Y run A   3: Y run B   4: g
        const int ns = st->nslots, base_slot = st->base;
        const int ra = min(h, ns - base_slot), rb = h - ra;
        const CUtensorMap *map = &maps->run[1];
        int row = 0;
        size_t off = 0;
        bool valid = true;
        switch (lane) {
        case 0: map = &maps->run[ra]; row = kArenaRowS + base_slot; off = 0; break;
        case 1: map = &maps->run[rb]; row = kArenaRowS; off = (size_t)ra * T; valid = rb > 0; break;
        case 2: map = &maps->run[ra]; row = kArenaRowS + ns + base_slot; off = (size_t)h * T; break;
        case 3: map = &maps->run[rb]; row = kArenaRowS + ns; off = (size_t)(h + ra) * T; valid = rb > 0; break;
        case 4: row = kArenaRowG; off = (size_t)(2 * h) * T; break;
        default: valid = false; break;
        }
        const unsigned bytes = (unsigned)(J * T * sizeof(double));
        int stage = 0;
        unsigned phase = 1u; // the producer waits for the PREVIOUS use of a stage to be released
        for (long long k = 0; k < my_tiles; ++k) {
            if (lane == 0) {
                if (k >= NS) mbar_wait(&empty[stage], phase);
                mbar_expect_tx(&full[stage], bytes);
            }
            __syncwarp();
            const int col = (int)((blockIdx.x + k * (long long)gridDim.x) * T);
            if (valid) tma_load_2d(tile + stage * stage_doubles + off, map, col, row, &full[stage]);
            if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
    } else {
        const int r0 = h - 1, r1 = 2 * h - 1, r2 = 2 * h;
        const int T2 = T >> 1, slice2 = T2 / NE;
        int stage = 0;
        unsigned phase = 0u;
        for (long long k = 0; k < my_tiles; ++k) {
            mbar_wait(&full[stage], phase);
            const double2 *cur = reinterpret_cast<const double2 *>(tile + stage * stage_doubles);
            const int e_end = eg < NE ? (eg + 1) * slice2 : 0; // columns beyond n were zero-filled by the TMA unit
            for (int e = eg * slice2 + lane; e < e_end; e += 32) { // (16 % NG != 0: the last warps own no slice)
                const double2 a0 = cur[r0 * T2 + e], a1 = cur[r1 * T2 + e], a2 = cur[r2 * T2 + e];
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                    const int j = cg + c * NG;
                    if (j < J) {
                        const double2 v = cur[j * T2 + e];
                        acc[c][0] = fma(a0.y, v.y, fma(a0.x, v.x, acc[c][0]));
                        acc[c][1] = fma(a1.y, v.y, fma(a1.x, v.x, acc[c][1]));
                        acc[c][2] = fma(a2.y, v.y, fma(a2.x, v.x, acc[c][2]));
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stage]);
            if (++stage == NS) { stage = 0; phase ^= 1u; }
        }
    }
    __syncthreads();
    double *red = tile; // [NE][J*3]
    if (warp > 0 && eg < NE) {
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
    }
    __syncthreads();
    for (int q = threadIdx.x; q < J * 3; q += kWsThreads) {
        double s = 0.0;
        for (int g = 0; g < NE; ++g) s += red[g * (J * 3) + q];
        st->partials[(size_t)q * gridDim.x + blockIdx.x] = s;
    }
}

// one CTA per (column, row): fixed-order sum of the per-CTA partials -> rows[j*3 + r]
__global__ void __launch_bounds__(kScalarThreads) k_gram_finalize(DevState *st, int nparts)
{
    const int h = st->h;
    if (st->ctrl.done || h == 0 || (st->steepest && !st->sg_valid)) return;
    const int q = blockIdx.x;
    if (q >= 3 * (2 * h + 1)) return;
    __shared__ double sm[kScalarThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double v = 0.0;
    for (int i = threadIdx.x; i < nparts; i += kScalarThreads) v += st->partials[(size_t)q * nparts + i];
    v = warp_sum(v);
    if (lane == 0) sm[warp] = v;
    __syncthreads();
    if (warp == 0) {
        double w = (lane < kScalarThreads / 32) ? sm[lane] : 0.0;
#pragma unroll
        for (int o = kScalarThreads / 64; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (lane == 0) st->gram_rows[q] = w;
    }
}

// ---- diagnostic timeline (LBFGSB200_TIMELINE): rows of (op, t_in, %globaltimer now); ops >= 100 are sub-marks ----
__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_mark(DevState *st, int op, unsigned long long t_in)
{
    if (st->tl && threadIdx.x == 0 && st->tl_n < st->tl_cap && (op < 100 || st->tl_sub)) {
        unsigned long long *row = st->tl + 3 * (size_t)st->tl_n++;
        row[0] = (unsigned long long)(long long)op;
        row[1] = t_in;
        row[2] = op >= 100 ? (unsigned long long)clock64() : global_ns(); // sub-marks: SM cycles next to the ns in row[1]
    }
}

// Gram bookkeeping + the two loops of seq/lbfgs.cpp:93-143 on coefficients.  Called by every thread of the scalar
// kernel's CTA.
// rows: the finalised 3 x J inner products of pass A (already summed over ranks on multi-GPU).
// Gs: J*J doubles of shared memory; the window Gram matrix is staged there so that the 2h dependent steps of the
//     recursion run at shared-memory / register latency (in HBM they cost ~1.4 us each: 29 us at m = 10).
// fresh: the newest pair was committed by the last accept, so its rows are new.  run == false: only the Gram
// bookkeeping (the direction is d = -g anyway, but later iterations need the rows of this pair).
//
// The recursion keeps, next to the coefficients delta (q = sum_j delta_j b_j), the PROJECTIONS u_i = b_i . q of q on
// every basis vector, one (or up to four) per lane of warp 0.  A step of either loop changes ONE coefficient,
// delta_k += c, i.e. q += c b_k, so every projection follows with one fma, u_i += c G[i][k] -- what the reference's
// vector recursion does to s_i . q when it updates q (seq/lbfgs.cpp:112, :139) -- and the inner product the next
// step needs is simply read from the lane that owns it.  No reductions inside the 2h dependent steps: ~80 cycles
// per step instead of ~450 (a 21-term dot product through the shuffle tree plus an FP64 division).
__device__ void compact_recursion(DevState *st, const double *rows, double *Gs, int fresh, bool run)
{
    const int h = st->h, J = 2 * h + 1, ns = st->nslots, NB = 2 * ns + 1;
    const bool seq = st->profile == LBFGSB200_PROFILE_SEQ;
    double *G = st->gram;
    __shared__ double delta_s[kMaxCols];
    __shared__ int bi[kMaxCols]; // basis index of window column j
    for (int j = threadIdx.x; j < J; j += kScalarThreads)
        bi[j] = j < h ? slot_of(*st, j) : (j < 2 * h ? ns + slot_of(*st, j - h) : 2 * ns);
    __syncthreads();
    // Window Gram matrix: rows / columns of the newest pair and of g come from pass A (and are written through to
    // the persistent matrix G), everything else from G -- ONE round trip to global memory.
    const int js = h - 1, jy = 2 * h - 1, jg = 2 * h; // window columns of s_newest, y_newest, g
    for (int idx = threadIdx.x; idx < J * J; idx += kScalarThreads) {
        const int a = idx / J, b = idx - a * J;
        double v;
        bool from_rows = true;
        if (a == jg) v = rows[b * 3 + 2];
        else if (b == jg) v = rows[a * 3 + 2];
        else if (fresh && a == js) v = rows[b * 3 + 0];
        else if (fresh && b == js) v = rows[a * 3 + 0];
        else if (fresh && a == jy) v = rows[b * 3 + 1];
        else if (fresh && b == jy) v = rows[a * 3 + 1];
        else { v = G[bi[a] * NB + bi[b]]; from_rows = false; }
        Gs[idx] = v;
        if (from_rows) G[bi[a] * NB + bi[b]] = v;
    }
    __syncthreads();
    tl_mark(st, 104, global_ns()); // (window Gram matrix staged)
    if (!run || threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    // Everything the 2h dependent steps touch lives in STATIC shared memory and is addressed directly, the window Gram
    // matrix through its 32-bit shared address computed once and made opaque to the compiler.  (Measured with the
    // diagnostic marks of benchmarks/timeline.py: with generic accesses to the dynamic array -- each re-deriving the
    // shared window from a special register -- and indexed shuffles, a step cost ~1750 cycles; a dependent
    // LDS + DMUL + LDS + DFMA + STS chain is ~150.  A variant that kept u, rho and the lane's Gram row in REGISTERS
    // and broadcast with shuffles -- 58 cycles per step in isolation (benchmarks/micro/latency.cu) -- measured 19 us for
    // the two loops inside this kernel against 7.5 us for the loops below, with or without a convergence point in
    // front of it, and was dropped; why this kernel runs its dependent chains ~3x slower than the micro-benchmark does
    // is not understood.)
    __shared__ double us[kMaxCols], rhos[kMaxCompactM], als[kMaxCompactM];
    // 32-bit shared addresses, computed once and made opaque to the compiler: left to itself it re-derives the shared
    // window from SR_CgaCtaId (an S2R of several hundred cycles) in EVERY iteration of the loops below, for the static
    // arrays as well as for the dynamic one
    auto opaque = [](const void *p) {
        unsigned a = smem_u32(p);
        asm volatile("mov.u32 %0, %0;" : "+r"(a));
        return a;
    };
    const unsigned gs_a = opaque(Gs), us_a = opaque(us), rh_a = opaque(rhos), al_a = opaque(als), ds_a = opaque(delta_s);
    auto lds = [](unsigned a) {
        double v;
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
        return v;
    };
    auto sts = [](unsigned a, double v) { asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory"); };
    auto gs_at = [&](int idx) { return lds(gs_a + 8u * (unsigned)idx); };
    int bad = 0;
    double gamma;
    __syncwarp(); // (converged from here on: thread 0 has just been off on its own in tl_mark)
    for (int i = lane; i < J; i += 32) sts(us_a + 8u * i, gs_at(i * J + jg)); // q = g: u_i = b_i . g
    // rho_p = 1 / (y_p . s_p), once per pair (seq/lbfgs.cpp:102, :135); pairs the CUDA profile skips get rho = 0, which
    // zeroes their alpha and their (alpha - beta) exactly as the skip flag does (par/L-BFGS.cu:222-223)
    for (int p = lane; p < h; p += 32) {
        double rho = 1.0 / gs_at((h + p) * J + p);
        if (seq && !isfinite(rho)) bad = 1;
        if (st->skip[slot_of(*st, p)]) rho = 0.0;
        sts(rh_a + 8u * p, rho);
    }
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
    // Every coefficient is touched exactly once (delta_{y_p} = -alpha_p in loop 1, delta_{s_p} = alpha_p - beta_p in
    // loop 2, delta_g = 1), so the coefficients are assembled after the loops; a step only moves the projections.
    auto bump = [&](int k, double c) { // q += c b_k : u_i += c G[i][k], every lane its own indices
        for (int i = lane; i < J; i += 32) sts(us_a + 8u * i, fma(c, gs_at(i * J + k), lds(us_a + 8u * i)));
        __syncwarp();
    };
    // first loop, newest -> oldest (seq/lbfgs.cpp:100-114)
    for (int p = h - 1; p >= 0; --p) {
        const double a = lds(rh_a + 8u * p) * lds(us_a + 8u * p); // rho_p (s_p . q)
        if (lane == 0) sts(al_a + 8u * p, a);
        bump(h + p, -a);                                          // q -= a y_p
    }
    const double ys = gs_at(js * J + jy), yy = gs_at(jy * J + jy);
    gamma = ys / yy; // :117
    if (seq) {
        if (gamma <= 0 || !isfinite(gamma)) bad = 1;
    } else {
        gamma = (yy > 0 && ys > 1e-10) ? ys / yy : 1.0; // par/L-BFGS.cu:246-255
    }
    for (int i = lane; i < J; i += 32) sts(us_a + 8u * i, lds(us_a + 8u * i) * gamma); // r = gamma q
    __syncwarp();
    // second loop, oldest -> newest (:133-141)
    for (int p = 0; p < h; ++p) {
        const double beta = lds(rh_a + 8u * p) * lds(us_a + 8u * (h + p)); // rho_p (y_p . r)
        const double c = lds(al_a + 8u * p) - beta;                        // (a skipped pair has alpha = beta = 0)
        if (lane == 0) sts(ds_a + 8u * p, c);                              // coefficient of s_p: 0 * gamma + c
        bump(p, c);                                                        // r += (alpha_p - beta) s_p
    }
    for (int p = lane; p < h; p += 32) sts(ds_a + 8u * (h + p), (0.0 - lds(al_a + 8u * p)) * gamma); // coefficient of y_p: (0 - alpha_p) gamma
    if (lane == 0) sts(ds_a + 8u * jg, 1.0 * gamma);                                                 // coefficient of g
    __syncwarp();
    tl_mark(st, 105, global_ns()); // (both loops done)
    for (int i = lane; i < J; i += 32) st->delta[i] = delta_s[i];
    for (int p = lane; p < h; p += 32) st->alpha[p] = als[p];
    __syncwarp();
    if (lane == 0) {
        st->gamma = gamma;
        if (st->fused && st->nranks > 1) {
            // the neighbours' boundary d, with the fma chain of k_combine_trial on THEIR boundary elements
            // (DevState::bL / bR): bit-identical to what they compute, and known before the combine pass runs
            double sl = 0.0, sr = 0.0;
            for (int j = 0; j < J; ++j) {
                sl = fma(delta_s[j], st->bL[bi[j]], sl);
                sr = fma(delta_s[j], st->bR[bi[j]], sr);
            }
            st->dL = -sl;
            st->dR = -sr;
        }
        if (bad) { // non-finite rho / bad gamma => d = -g (:103-108, :119-124)
            st->steepest = 1;
            if (!st->fused) st->vec_streams += 2.0; // the fused flow accounts for its direction pass in OP_F_DIR
        }
    }
}

// In-kernel version of k_gram_finalize for the 1-CTA scalar kernel: warp w sums rows w, w+8, ... of the
// per-CTA partials (lane-strided partial sums, then the fixed shuffle tree: deterministic), four rows
// in flight per warp.  out: 3 x (2h+1) sums followed by zeros up to `count`.
__device__ __forceinline__ void gram_rows_from_partials(const double *__restrict__ partials, int nparts, int nrows,
                                                        int count, double *out /* shared */)
{
    constexpr int W = kScalarThreads / 32, R = 4, C = 8;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q0 = warp; q0 < nrows; q0 += W * R) {
        double v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = 0.0;
        // all R x C loads of a batch are issued before the first use: one round trip to L2 per batch (a plain loop costs one
        // per 32 partials: ~0.7 us each, 5 in a row for the 148 CTAs of a B200)
        for (int i0 = 0; i0 < nparts; i0 += 32 * C) {
            double t[R][C];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int q = q0 + r * W;
#pragma unroll
                for (int c = 0; c < C; ++c) {
                    const int i = i0 + lane + 32 * c;
                    t[r][c] = (q < nrows && i < nparts) ? partials[(size_t)q * nparts + i] : 0.0;
                }
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int c = 0; c < C; ++c) v[r] += t[r][c]; // lane-strided, ascending: the order of the plain loop
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const double t = warp_sum(v[r]);
            if (lane == 0 && q0 + r * W < nrows) out[q0 + r * W] = t;
        }
    }
    for (int q = nrows + threadIdx.x; q < count; q += kScalarThreads) out[q] = 0.0;
    __syncthreads();
}

// pass B (unfused compact flow).  w = d = -(sum_j delta_j b_j), accumulated in window-column order; partial of g.d.
__global__ void __launch_bounds__(kThreads, kCombineCtasPerSm) k_combine(const DevState *__restrict__ st)
{
    const int h = st->h;
    if (st->ctrl.done || st->steepest || h == 0) return;
    __shared__ const double *cols[kMaxCols];
    __shared__ double coef[kMaxCols];
    const int J = 2 * h + 1;
    for (int j = threadIdx.x; j < J; j += kThreads) {
        cols[j] = basis_col(st, j, h);
        coef[j] = st->delta[j];
    }
    __syncthreads();
    const long long n = st->n;
    const long long nvec_pad = (n + 1) >> 1;
    double *__restrict__ w = st->w;
    double acc_gd = 0.0;
    constexpr int U = 2; // double2 items per thread per step
    const long long step = (long long)gridDim.x * kThreads * U;
    for (long long i0 = ((long long)blockIdx.x * kThreads + threadIdx.x); i0 < nvec_pad; i0 += step) {
        // items i0 and i0 + gridDim.x*kThreads
        const long long i1 = i0 + (long long)gridDim.x * kThreads;
        const bool has1 = i1 < nvec_pad;
        double2 s0 = make_double2(0.0, 0.0), s1 = s0, g0 = s0, g1 = s0;
#pragma unroll 8
        for (int j = 0; j < J; ++j) {
            const double c = coef[j];
            const double2 v0 = ld2(cols[j], i0);
            const double2 v1 = has1 ? ld2(cols[j], i1) : make_double2(0.0, 0.0);
            s0.x = fma(c, v0.x, s0.x); s0.y = fma(c, v0.y, s0.y);
            s1.x = fma(c, v1.x, s1.x); s1.y = fma(c, v1.y, s1.y);
            if (j == J - 1) { g0 = v0; g1 = v1; }
        }
        s0.x = -s0.x; s0.y = -s0.y; s1.x = -s1.x; s1.y = -s1.y;
        st2(w, i0, s0);
        acc_gd += g0.x * s0.x + g0.y * s0.y;
        if (has1) {
            st2(w, i1, s1);
            acc_gd += g1.x * s1.x + g1.y * s1.y;
        }
    }
    double v[1] = {acc_gd};
    block_emit<1>(v, st->partials);
}

} // namespace lb
