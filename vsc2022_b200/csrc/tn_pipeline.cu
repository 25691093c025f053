// Temporal-network (TN) alignment, FAST PIPELINE for aligned rows (B200, sm_100a).
//
// Replaces vcsl.vta `tn` (alipay/VCSL @ c39269d5) behind vsc/baseline/localization.py:44-46,58.
// Contract: oracle/tn_networkx.py.  Four launches per batch, all on one stream:
//
//   T1  tn_topk_kernel   HBM-bound.  Persistent warps; each lane pulls ITS OWN row of a 32-row
//                        tile through shared memory in 64-column panels (cp.async.bulk = TMA bulk copy, mbarrier
//                        completion, three stages) and selects the exact top-K of the row
//                        (row_select.cuh).  The similarity matrices are read from HBM once;
//                        the node table (ref index + similarity per node, 6 B) is written out.
//   T1e tn_edges_kernel  one thread per source row: constraints C1-C4 -> predecessor bit-masks.
//   T2  tn_dp_kernel     longest-path sweeps, FOUR pairs per warp (8 lanes each, one lane per
//                        rank of a row layer).  First sweep visits all layers; later sweeps only
//                        the layers downstream of the chain whose edges were zeroed.  End node =
//                        first maximum in Kahn order (networkx); ties are broken by Kahn
//                        generation, unresolved ties hand the pair to the general kernel.
//   T3  tn_maxsim_kernel max similarity inside each kept box (only when requested).
//
// Pairs the pipeline cannot take (lr % 4 != 0, unaligned start, lr < K, tie-heavy rows,
// ambiguous end-node ties) are appended to a work list that tn_fused.cu finishes.
//
// Node v = q*K + rank.  Edge (q_src,a) -> (q_dst,b) is bit
// slot = (step-1-(q_dst-q_src))*K + a of pred[v_dst]; ascending slot == networkx predecessor
// insertion order (see oracle/tn_fast.c).
#include <limits.h>
#include <stdlib.h>

#include "row_select.cuh"
#include "tn_common.cuh"

namespace {

using vsc::kFullMask;
using vsc::tn::Batch;
using vsc::tn::WorkList;
using vsc::tn::kMaxBoxes;
using vsc::tn::kMaxTop;
using vsc::tn::Workspace;
using vsc::tn::NodeRec;
using vsc::tn::kRefMask;
using vsc::tn::kSimOk;

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// Ampere-style 16-byte async copy global -> shared (bypasses L1); one commit group per panel.
template <int OFFSET>
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0+%2], [%1], 16;" ::"r"(dst_smem), "l"(src), "n"(OFFSET) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ------------------------------------------------------------------ T1: row top-K
// A warp owns a TILE of 32 rows (one per lane in the compute phase) and streams it through shared memory
// in 64-column panels.  The panel fill is warp-cooperative: 16 lanes x 16 B cover one 256-byte row
// segment, so every cp.async instruction moves two fully coalesced segments into a padded panel
// (pitch 68 words -> conflict-free LDS.128 for the thread-per-row reads); kStages panels per warp
// are in flight.  (Measured: one cp.async.bulk per lane and panel -- 12.8 M 256-byte TMA bulk
// copies per batch -- was TMA-issue-bound at ~27 cycles/op/SM, 2.4 TB/s; see profiles/.)
// Pass 1 (block maxima + K best maxima) runs on the panels; at the end of a row the K hot blocks
// are re-read straight from global memory (L2 hits) for the exact selection, whose scratch (16-float
// stash + candidate list) is the lane's own row of the panel it has just consumed.
// Shared memory per warp = kStages panels + block maxima for ceil(max_lr/16) blocks, so 11 warps fit one SM
// at 300 columns (8 with the first layout): the kernel is bound by warps in flight, not by instruction issue.
// Work unit = tile: tile t covers rows [32*(t % tpp), +32) of pair t / tpp (tpp = ceil(max_lq/32)); tiles are
// claimed with one atomic each, TWO tiles ahead, so neither the atomic nor the pair-descriptor loads
// sit on a warp's critical path.
constexpr int kT1MaxWarps = 16;
constexpr int kT1MaxThreads = kT1MaxWarps * 32;
constexpr int kTileRows = 32;
constexpr int kPanelCols = 64;
constexpr int kPanelPitch = 68;
#ifndef VSC_T1_STAGES
#define VSC_T1_STAGES 2
#endif
constexpr int kStages = VSC_T1_STAGES;
constexpr int kPanelBytes = kTileRows * kPanelPitch * 4;
constexpr int kScratchStash = 0, kScratchVal = 16, kScratchCol = 32;   // word offsets inside the lane's panel row
static_assert(kScratchCol + vsc::kMaxCand <= kPanelCols, "selection scratch must fit one panel row");

struct Cursor {   // one panel of work
    int pair, row0, col0, lq, lr;
    const float *base;
};

__host__ __device__ inline size_t t1_warp_bytes(int max_lr) {
    const size_t blocks = (size_t)(max_lr + vsc::kBlockCols - 1) / vsc::kBlockCols;
    const size_t b = (size_t)kStages * kPanelBytes + blocks * 32 * 4 + kStages * sizeof(Cursor);
    return (b + 127) / 128 * 128;
}

struct T1Args {
    Batch b;
    Workspace w;
    WorkList out;
    int tiles_per_pair, n_tiles, warp_bytes;
};

template <int K>
__global__ void __launch_bounds__(kT1MaxThreads, 1) tn_topk_kernel(const T1Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *mine = smem_raw + (size_t)warp * a.warp_bytes;
    float *panels = reinterpret_cast<float *>(mine);                       // [kStages][kTileRows * kPanelPitch]
    float *bm = panels + kStages * kTileRows * kPanelPitch;                // [blocks][32]
    Cursor *ring = reinterpret_cast<Cursor *>(mine + a.warp_bytes) - kStages;

    // ---- tile stream: `ic` = panel to request next, `nx` = descriptor of the following tile (loads in flight),
    // `t_req` (lane 0) = id of the tile after that (atomic in flight)
    Cursor ic, nx;
    int t_req = 0;
    bool more = true;
    auto describe = [&](int t, Cursor &c) {   // issues the loads; nothing here waits for them
        c.pair = -1; c.row0 = 0; c.col0 = 0; c.lq = 0; c.lr = 0; c.base = nullptr;
        if (t < a.n_tiles) {
            const int p = t / a.tiles_per_pair;
            c.pair = p; c.row0 = (t - p * a.tiles_per_pair) * kTileRows;
            c.lq = a.b.lq[p]; c.lr = a.b.lr[p]; c.base = a.b.sims + a.b.off[p];
        }
    };
    auto advance = [&]() {   // ic <- next tile this warp can process; pairs it cannot take go to the general kernel
        for (;;) {
            ic = nx;
            const int t = __shfl_sync(kFullMask, t_req, 0);
            if (lane == 0) t_req = atomicAdd(a.w.cursor, 1);
            describe(t, nx);
            if (ic.pair < 0) { more = false; return; }   // tile ids only grow: nothing is left
            const bool ok = (ic.lr & 3) == 0 && (reinterpret_cast<uintptr_t>(ic.base) & 15u) == 0 && ic.lr >= a.b.topk &&
                            ic.lr <= a.b.max_lr && ic.lr <= vsc::kBlockCols * vsc::kMaxRowBlocks && ic.lq <= a.b.max_lq;
            if (!ok) {
                if (lane == 0 && ic.row0 == 0 && atomicExch(&a.w.skip[ic.pair], 1) == 0)
                    a.out.list[atomicAdd(a.out.count, 1)] = ic.pair;
                continue;
            }
            if (ic.row0 >= ic.lq) continue;   // ragged batch: this tile slot is empty (lq <= 0: T2 reports zero boxes)
            return;
        }
    };
    {
        int t0 = 0;
        if (lane == 0) { t0 = atomicAdd(a.w.cursor, 1); t_req = atomicAdd(a.w.cursor, 1); }
        t0 = __shfl_sync(kFullMask, t0, 0);
        describe(t0, nx);
        advance();
    }

    const int half = lane >> 4, ch = (lane & 15) * 4;  // fill role: row parity, column offset
    const uint32_t panel_s = smem_u32(panels) + (uint32_t)(half * kPanelPitch + ch) * 4u;
    auto issue = [&](int stage) {    // request panel `ic` into `stage`, then step `ic`
        if (lane == 0) ring[stage] = ic;
        if (ic.col0 + ch < ic.lr) {
            const float *src = ic.base + (size_t)(ic.row0 + half) * ic.lr + ic.col0 + ch;
            const uint32_t dst = panel_s + (uint32_t)stage * kPanelBytes;
            const int n = (min(ic.lq - ic.row0, kTileRows) - half + 1) >> 1;  // rows of this lane's parity in the tile
            const size_t stride = (size_t)2 * ic.lr;
#define VSC_COPY_ROW(I)                                                        \
            if ((I) < n) cp_async16<(I) * 2 * kPanelPitch * 4>(dst, src);        \
            src += stride;
            VSC_COPY_ROW(0) VSC_COPY_ROW(1) VSC_COPY_ROW(2) VSC_COPY_ROW(3)
            VSC_COPY_ROW(4) VSC_COPY_ROW(5) VSC_COPY_ROW(6) VSC_COPY_ROW(7)
            VSC_COPY_ROW(8) VSC_COPY_ROW(9) VSC_COPY_ROW(10) VSC_COPY_ROW(11)
            VSC_COPY_ROW(12) VSC_COPY_ROW(13) VSC_COPY_ROW(14) VSC_COPY_ROW(15)
#undef VSC_COPY_ROW
        }
        cp_async_commit();
        ic.col0 += kPanelCols;
        if (ic.col0 >= ic.lr) advance();
    };

    unsigned issued = 0, consumed = 0;
    for (; issued < kStages - 1 && more; ++issued) issue(issued);
    vsc::RowTopK<K> sel;
    sel.reset();
    while (consumed < issued) {
        if (more) { issue(issued % kStages); ++issued; }
        else cp_async_commit();      // empty group keeps the wait depth constant at the tail
        const int stage = consumed % kStages;
        cp_async_wait<kStages - 1>();
        __syncwarp();                // every lane's copies for this panel have landed
        const Cursor c = ring[stage];
        const int row = c.row0 + lane;
        const bool last_panel = c.col0 + kPanelCols >= c.lr;
        bool ok = true;
        if (row < c.lq) {
            if (c.col0 == 0) sel.reset();
            float *myrow = panels + stage * (kTileRows * kPanelPitch) + lane * kPanelPitch;
            const float4 *src = reinterpret_cast<const float4 *>(myrow);
            const int chunks = (min(kPanelCols, c.lr - c.col0)) >> 2;
#pragma unroll
            for (int blk = 0; blk < kPanelCols / vsc::kBlockCols; ++blk) {
                const int left = chunks - blk * 4;
                if (left <= 0) break;
                float4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < left) v[k] = src[blk * 4 + k];
                bm[(c.col0 / vsc::kBlockCols + blk) * 32 + lane] = sel.add_block(v, left < 4 ? left : 4);
            }
            if (last_panel) {
                float val[K]; int col[K];
                // the lane's row of this panel is consumed: it becomes the selection scratch
                ok = sel.finish(c.base + (size_t)row * c.lr, c.lr, bm + lane, 32, myrow + kScratchVal,
                                reinterpret_cast<int *>(myrow + kScratchCol), 1, myrow + kScratchStash, val, col);
                const size_t node = (size_t)c.pair * a.b.max_nodes + (size_t)row * K;
                unsigned char *rec = static_cast<unsigned char *>(a.w.rec) + node * a.w.rec_bytes;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    a.w.ref_of[node + i] = (uint16_t)col[i] | (val[i] >= a.b.min_sim ? kSimOk : (uint16_t)0);
                    // whole record in 16-byte stores: masks and distance start at zero
                    if (a.w.rec_bytes == 16) {
                        *reinterpret_cast<float4 *>(rec + i * 16) = make_float4(0.f, 0.f, val[i], 0.f);
                    } else {
                        *reinterpret_cast<float4 *>(rec + i * 32) = make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4 *>(rec + i * 32 + 16) = make_float4(val[i], 0.f, 0.f, 0.f);
                    }
                }
            }
        }
        if (last_panel && __any_sync(kFullMask, !ok)) {   // a row with too many tied candidates: general kernel
            if (lane == 0 && atomicExch(&a.w.skip[c.pair], 1) == 0)
                a.out.list[atomicAdd(a.out.count, 1)] = c.pair;
        }
        __syncwarp();  // every lane is done with `stage` before it is refilled
        ++consumed;
    }
}

// ------------------------------------------------------------------ predecessor-mask layout
// Edge (q_src, a) -> (q_dst, b) is bit  slot = grp * GS + a  of pred[v_dst], grp = step-1-(q_dst-q_src): ascending
// slot order == networkx predecessor insertion order (farthest source row first, then rank; oracle/tn_fast.c).
// GS (group stride) is 8 on the fast variant (32-bit masks, (step-1)*8 <= 32), which turns the slot -> distance-
// window index into one add and one mask; every other parameter set uses 64-bit masks with GS = K.

// ------------------------------------------------------------------ T1e: edges
// One thread per source row.  Per destination row the K*K rank combinations are screened for
// constraint C2 with two instructions each into a hit mask; only the hits (about one per three
// row pairs) go through C3 (no ref already linked from this row inside [r_src, r_dst]), C4
// (destination similarity >= min_sim: a flag T1 stored next to the reference index, so this kernel never
// touches the node records except for the atomic OR) and the atomic OR into the destination's predecessor mask.
template <typename MaskT, int K, int GS>
__global__ void __launch_bounds__(256) tn_edges_kernel(const Batch b, const Workspace w) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = (int)(idx / b.max_lq);
    if (pair >= b.n_pairs) return;
    const int q_src = (int)(idx - (long long)pair * b.max_lq);
    const int lq = b.lq[pair];
    if (q_src >= lq || w.skip[pair]) return;
    const int step = b.step;
    const size_t nb = (size_t)pair * b.max_nodes;
    const uint16_t *ref_of = w.ref_of + nb;
    NodeRec<MaskT> *rec = static_cast<NodeRec<MaskT> *>(w.rec) + nb;

    int r_src[K]; uint32_t window[K];  // window[a]: refs linked from this row, relative to r_src[a]
#pragma unroll
    for (int x = 0; x < K; ++x) { r_src[x] = ref_of[q_src * K + x] & kRefMask; window[x] = 0; }
    const int q_end = min(lq, q_src + step);
    for (int q_dst = q_src + 1; q_dst < q_end; ++q_dst) {
        int r_dst[K];
        uint32_t sim_ok = 0;
#pragma unroll
        for (int bb = 0; bb < K; ++bb) {
            const uint16_t e = ref_of[q_dst * K + bb];
            r_dst[bb] = e & kRefMask;
            sim_ok |= (e & kSimOk) ? 1u << bb : 0u;
        }
        uint32_t hits[K];  // hits[x]: destination ranks b with 0 < r_dst[b] - r_src[x] < step (C2) and C4
#pragma unroll
        for (int x = 0; x < K; ++x) hits[x] = 0;
#pragma unroll
        for (int bb = 0; bb < K; ++bb) {
#pragma unroll
            for (int x = 0; x < K; ++x)
                hits[x] |= ((unsigned)(r_dst[bb] - r_src[x] - 1) < (unsigned)(step - 1) ? 1u : 0u) << bb;
        }
        uint32_t accepted = 0;
        const int slot0 = (step - 1 - (q_dst - q_src)) * GS;
#pragma unroll
        for (int x = 0; x < K; ++x) {
            uint32_t h = hits[x] & sim_ok;                               // C4
            while (h) {
                const int bb = __ffs(h) - 1; h &= h - 1;
                int rd = r_dst[0];
#pragma unroll
                for (int i = 1; i < K; ++i) rd = bb == i ? r_dst[i] : rd;   // register select, no local memory
                const int d = rd - r_src[x];
                if (window[x] & ((2u << d) - 1u)) continue;              // C3
                accepted |= 1u << bb;
                const MaskT bit = (MaskT)1 << (slot0 + x);
                if (sizeof(MaskT) == 8)
                    atomicOr(reinterpret_cast<unsigned long long *>(&rec[q_dst * K + bb].pred), (unsigned long long)bit);
                else
                    atomicOr(reinterpret_cast<unsigned int *>(&rec[q_dst * K + bb].pred), (unsigned int)bit);
            }
        }
        while (accepted) {  // refs linked in this step constrain the later destination rows
            const int bb = __ffs(accepted) - 1; accepted &= accepted - 1;
            int rd = r_dst[0];
#pragma unroll
            for (int i = 1; i < K; ++i) rd = bb == i ? r_dst[i] : rd;
#pragma unroll
            for (int x = 0; x < K; ++x) {
                const int d = rd - r_src[x];
                if ((unsigned)d < (unsigned)step) window[x] |= 1u << d;
            }
        }
    }
}

template <typename MaskT, int GS8>
void launch_edges(const Batch &b, const Workspace &w, int grid, cudaStream_t stream) {
    switch (b.topk) {
        case 1: tn_edges_kernel<MaskT, 1, GS8 ? 8 : 1><<<grid, 256, 0, stream>>>(b, w); break;
        case 2: tn_edges_kernel<MaskT, 2, GS8 ? 8 : 2><<<grid, 256, 0, stream>>>(b, w); break;
        case 3: tn_edges_kernel<MaskT, 3, GS8 ? 8 : 3><<<grid, 256, 0, stream>>>(b, w); break;
        case 4: tn_edges_kernel<MaskT, 4, GS8 ? 8 : 4><<<grid, 256, 0, stream>>>(b, w); break;
        case 5: tn_edges_kernel<MaskT, 5, GS8 ? 8 : 5><<<grid, 256, 0, stream>>>(b, w); break;
        case 6: tn_edges_kernel<MaskT, 6, GS8 ? 8 : 6><<<grid, 256, 0, stream>>>(b, w); break;
        case 7: tn_edges_kernel<MaskT, 7, GS8 ? 8 : 7><<<grid, 256, 0, stream>>>(b, w); break;
        default: tn_edges_kernel<MaskT, 8, 8><<<grid, 256, 0, stream>>>(b, w); break;
    }
}

// ------------------------------------------------------------------ T2: longest-path sweeps
// Four pairs per warp: an octet of lanes owns one pair, lane `sub` owns rank `sub` of the current
// row layer.  The kernel is one long dependent chain per pair (measured: 0.46 ms for a single pair,
// ~9 cycles per instruction with the SM to itself), so everything here is about instructions on that
// chain: (distance, Kahn generation) of the last D >= step layers sit interleaved in a shared-memory
// window read with one 64-bit load per predecessor, the 16-byte node records are prefetched into
// registers four layers ahead (loops unrolled by four so the rotation is free), and the chain of a
// round is walked in shared memory by one lane, then zeroed / scored by all eight lanes of the octet.
// The grid is sized so that every pair of a batch is resident at once.
__device__ unsigned long long g_dp_counters[8];
#ifdef VSC_TN_COUNTERS
#define VSC_CLK_START() long long cyc__[8] = {}; long long t__ = clock64()
#define VSC_CLK(i) do { const long long n__ = clock64(); cyc__[i] += n__ - t__; t__ = n__; } while (0)
#define VSC_CLK_ADD(i, n) cyc__[i] += (n)
#define VSC_CLK_FLUSH() do { if (lane == 0) { cyc__[7] = 1; for (int i__ = 0; i__ < 8; ++i__) atomicAdd(&g_dp_counters[i__], (unsigned long long)cyc__[i__]); } } while (0)
#else
#define VSC_CLK_START() do {} while (0)
#define VSC_CLK(i) do {} while (0)
#define VSC_CLK_ADD(i, n) do {} while (0)
#define VSC_CLK_FLUSH() do {} while (0)
#endif

#ifndef VSC_T2_WARPS
#define VSC_T2_WARPS 1
#endif
// 4 pairs per CTA: small CTAs pack the SMs so a whole batch is one wave (measured at 8000 pairs: 1 warp 0.485 ms,
// 2 warps 0.493 ms, 4 warps 0.74 ms = two waves)
constexpr int kT2Warps = VSC_T2_WARPS;
constexpr int kT2Threads = kT2Warps * 32;
#ifndef VSC_DP_AHEAD
#define VSC_DP_AHEAD 6
#endif
constexpr int kAhead = VSC_DP_AHEAD;    // node-record prefetch distance in row layers (2: 0.554 ms, 4: 0.505, 6: 0.493, 8: 0.493, 12: 0.494)

// Octet reductions by xor-shuffle (distances 1, 2, 4 stay inside an aligned group of 8 lanes).
// redux.sync with a per-octet mask compiles to a loop over the distinct masks and cost 38 % of
// this kernel's stall samples when profiled.
__device__ __forceinline__ uint32_t oct_max(uint32_t v, unsigned m) {
    v = max(v, __shfl_xor_sync(m, v, 1));
    v = max(v, __shfl_xor_sync(m, v, 2));
    return max(v, __shfl_xor_sync(m, v, 4));
}
__device__ __forceinline__ int oct_min(int v, unsigned m) {
    v = min(v, __shfl_xor_sync(m, v, 1));
    v = min(v, __shfl_xor_sync(m, v, 2));
    return min(v, __shfl_xor_sync(m, v, 4));
}
__device__ __forceinline__ int oct_add(int v, unsigned m) {
    v += __shfl_xor_sync(m, v, 1);
    v += __shfl_xor_sync(m, v, 2);
    return v + __shfl_xor_sync(m, v, 4);
}

// The per-layer maxima are taken once per kAhead layers from the window (layer_max) instead of with three shuffles and a
// ballot per layer (measured: see profiles/r02_tn_summary.md); VSC_DP_DEFER_MAX=0 restores the per-layer form.
#ifndef VSC_DP_DEFER_MAX
#define VSC_DP_DEFER_MAX 1
#endif
constexpr bool kDeferMax = VSC_DP_DEFER_MAX != 0;
static_assert(kAhead <= 8, "one lane of the octet per deferred layer");

__host__ __device__ inline int window_depth(int step) {  // power of two >= step (and >= kAhead: deferred layer maxima)
    int d = 1;
    while (d < step || (kDeferMax && d < kAhead)) d <<= 1;  // > step-1, so the slot being written is never one being read
    return d;
}

__host__ __device__ inline size_t t2_pair_bytes(int max_nodes, int max_lq, int step) {
    const size_t d = (size_t)window_depth(step);
    size_t b = d * 8 * 8;                                 // (distance, generation) window of a sweep ...
    if (b < (size_t)max_lq * 2) b = (size_t)max_lq * 2;   // ... shares its bytes with the chain of a round
    b = (b + 15) / 16 * 16;
    b += (size_t)max_lq * 4;                              // layer maxima
    b += (size_t)max_lq;                                  // ranks that attain the layer maximum
    b += (size_t)max_nodes;                               // best-predecessor slots
    return (b + 15) / 16 * 16;
}

// Node records change while the kernel runs (distance, zeroed-edge mask) and the zeroed bits are set
// with L2 atomics, so every record read goes to L2 (ld.global.cg), never through L1.
template <typename Rec>
__device__ __forceinline__ Rec load_rec(const Rec *p) {
    Rec r;
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint4 *d = reinterpret_cast<uint4 *>(&r);
    d[0] = __ldcg(s);
    if (sizeof(Rec) == 32) d[1] = __ldcg(s + 1);
    return r;
}

// Layer maximum and the ranks that attain it, read back from the window by ONE lane (distances are >= +0: their bit
// patterns are ordered as unsigned integers).
template <int K>
__device__ __forceinline__ void layer_max(const float2 *win, int base, uint32_t &best, uint8_t &ranks) {
    uint32_t v[K], m = 0;
#pragma unroll
    for (int r = 0; r < K; ++r) { v[r] = __float_as_uint(win[base + r].x); m = max(m, v[r]); }
    uint32_t at = 0;
#pragma unroll
    for (int r = 0; r < K; ++r) at |= (v[r] == m ? 1u : 0u) << r;
    // ranks >= K of the octet hold distance 0 in the window: they tie with an all-zero layer exactly as in the shuffle form
    if (m == 0u) at = 0xFFu;
    best = m; ranks = (uint8_t)at;
}

// One lane relaxes its own node against the distance window: FIRST maximal predecessor in
// ascending slot order (networkx keeps the first maximum).
template <typename MaskT, int K, int GS, bool FIRST>
__device__ __forceinline__ void relax_node(MaskT pm, MaskT zm, float w, int q, int step, int wmask,
                                           const float2 *win, float &best, int &best_slot, int &gen) {
    best = 0.0f; best_slot = -1; gen = 0;
    const int base8 = (q - step + 1) * 8, wm8 = wmask * 8 + 7;
    while (pm) {
        const int sl = sizeof(MaskT) == 8 ? __ffsll((long long)pm) - 1 : __ffs((int)pm) - 1;
        pm &= pm - 1;
        int wi;
        if (GS == 8) {
            wi = (base8 + sl) & wm8;   // ((q - (step-1-grp)) & wmask) * 8 + rank with slot = grp*8 + rank
        } else {
            const int grp = sl / GS;   // compile-time divisor
            wi = ((q - (step - 1 - grp)) & wmask) * 8 + (sl - grp * GS);
        }
        float cand; int pg = 0;
        if (FIRST) { const float2 e = win[wi]; cand = e.x + w; pg = __float_as_int(e.y); }
        else cand = win[wi].x + (((zm >> sl) & 1) ? 0.0f : w);
        if (best_slot < 0 || cand > best) { best = cand; best_slot = sl; }
        if (FIRST) gen = max(gen, pg + 1);
    }
    if (best_slot >= 0 && !(best >= 0.0f)) { best = 0.0f; best_slot = -1; }  // networkx: negative best -> (0, v)
}

// ---- exact order of two nodes of the SAME Kahn generation in networkx's topological order, without a Kahn pass.
// networkx.topological_generations visits a generation in order and appends a child to the next one when its LAST parent
// is processed, children of one parent in adjacency order (= node id order here).  So for two nodes of generation g:
// the one whose last parent comes first in generation g-1 comes first; same last parent -> lower id first.  Generation 0
// (zero in-degree) is in id order.  The last parent of a generation-1 node is its highest-id predecessor (the highest
// slot: ascending slots are ascending predecessor ids).  For g >= 2 the last parent is known without recursion only if
// exactly one predecessor lies in generation g-1; otherwise `undecided` is set and the pair goes to the exact-order kernel.
template <typename MaskT, int K, int GS>
__device__ __forceinline__ int parent_of_slot(int v, int sl, int step) {
    const int grp = sl / GS;
    return (v / K - (step - 1 - grp)) * K + (sl - grp * GS);
}
template <typename MaskT, int K, int GS>
__device__ int topo_compare(int a, int b, int g, int step, const NodeRec<MaskT> *rec, const uint16_t *gen, bool &undecided) {
    for (;;) {
        if (a == b) return 0;
        if (g <= 0) return a < b ? -1 : 1;
        const MaskT ma = load_rec(&rec[a]).pred, mb = load_rec(&rec[b]).pred;
        int pa = -1, pb = -1;
        if (g == 1) {
            pa = parent_of_slot<MaskT, K, GS>(a, sizeof(MaskT) == 8 ? 63 - __clzll((long long)ma) : 31 - __clz((int)ma), step);
            pb = parent_of_slot<MaskT, K, GS>(b, sizeof(MaskT) == 8 ? 63 - __clzll((long long)mb) : 31 - __clz((int)mb), step);
        } else {
            int na = 0, nb = 0;
            for (MaskT m = ma; m; m &= m - 1) {
                const int u = parent_of_slot<MaskT, K, GS>(a, sizeof(MaskT) == 8 ? __ffsll((long long)m) - 1 : __ffs((int)m) - 1, step);
                if (gen[u] == g - 1) { pa = u; ++na; }
            }
            for (MaskT m = mb; m; m &= m - 1) {
                const int u = parent_of_slot<MaskT, K, GS>(b, sizeof(MaskT) == 8 ? __ffsll((long long)m) - 1 : __ffs((int)m) - 1, step);
                if (gen[u] == g - 1) { pb = u; ++nb; }
            }
            if (na != 1 || nb != 1) { undecided = true; return a < b ? -1 : 1; }
        }
        if (pa == pb) return a < b ? -1 : 1;
        a = pa; b = pb; --g;
    }
}

// The exact order once the LAST PARENT of every ancestor is known (resolve pass in tn_dp_kernel): walk both last-parent
// chains down the generations until they meet or reach generation 0.  Entries of nodes that are no ancestors of the
// candidates may be stale; the clamp keeps their reads inside the pair.
template <typename MaskT, int K, int GS>
__device__ int topo_compare_exact(int a, int b, int g, int step, const uint8_t *last_parent) {
    for (;;) {
        if (a == b) return 0;
        if (g <= 0) return a < b ? -1 : 1;
        const int pa = max(0, parent_of_slot<MaskT, K, GS>(a, __ldcg(&last_parent[a]), step));
        const int pb = max(0, parent_of_slot<MaskT, K, GS>(b, __ldcg(&last_parent[b]), step));
        if (pa == pb) return a < b ? -1 : 1;
        a = pa; b = pb; --g;
    }
}

template <typename MaskT, int K, int GS>
__global__ void __launch_bounds__(kT2Threads, 16 / kT2Warps) tn_dp_kernel(const Batch b, const Workspace w, const WorkList out,
                                                                          const int pairs_per_warp) {
    // WARP-SYNCHRONOUS: the four octets of a warp run every loop together (trip count = the longest of the
    // four, shorter ones are predicated off) and all shuffles use the full mask, so the warp never splits into
    // four serially executed instruction streams.
    extern __shared__ __align__(16) unsigned char t2_smem[];
    const int step = b.step;
    const int lane = threadIdx.x & 31, sub = lane & 7, oct = lane >> 3;
    const unsigned om = 0xFFu << (oct * 8);
    // pairs_per_warp < 4 (small batches): the warp's loops run as long as the LONGEST of its pairs needs, so with few
    // pairs per GPU a pair is faster in a warp of its own; the unused octets are dead (see below)
    const int pair = (blockIdx.x * kT2Warps + (threadIdx.x >> 5)) * pairs_per_warp + oct;
    const bool alive = oct < pairs_per_warp && pair < b.n_pairs && !w.skip[pair < b.n_pairs ? pair : 0];
    if (!__any_sync(kFullMask, alive)) return;
    const int p = alive ? pair : 0;   // dead octets shadow pair 0 read-only and never store

    const int depth = window_depth(step), wmask = depth - 1;
    unsigned char *mine = t2_smem + (size_t)((threadIdx.x >> 5) * 4 + oct) * t2_pair_bytes(b.max_nodes, b.max_lq, step);
    // The distance window is only alive inside a sweep (every sweep refills what it reads) and the chain only
    // between two sweeps, so they share one region: that is what lets a whole 8000-pair batch stay resident.
    const size_t shared_region = (max((size_t)depth * 64, (size_t)b.max_lq * 2) + 15) / 16 * 16;
    float2 *win = reinterpret_cast<float2 *>(mine);                     // {distance, generation bits}
    uint16_t *chain = reinterpret_cast<uint16_t *>(mine);               // nodes of the chain, end node first
    uint32_t *lbest = reinterpret_cast<uint32_t *>(mine + shared_region);  // per-layer maximum distance (float bits)
    uint8_t *lrank = reinterpret_cast<uint8_t *>(lbest + b.max_lq);     // bit r: node (q, r) attains lbest[q]
    int8_t *slot = reinterpret_cast<int8_t *>(lrank + b.max_lq);        // best-predecessor slot per node

    const int lq = alive ? b.lq[p] : 0;
    const int lq_max = __reduce_max_sync(kFullMask, lq);
    const int box_cap = b.max_path + 1;
    int4 *boxes = reinterpret_cast<int4 *>(b.boxes) + (size_t)p * box_cap;
    const size_t nb = (size_t)p * b.max_nodes;
    using Rec = NodeRec<MaskT>;
    Rec *rec = static_cast<Rec *>(w.rec) + nb;
    const uint16_t *ref_of = w.ref_of + nb;
    uint16_t *gen = w.gen + nb;
    const bool ranked = sub < K;

    for (int i = sub; i < depth * 8; i += 8) win[i] = make_float2(0.0f, 0.0f);
    __syncwarp();

    VSC_CLK_START();
    // ---- first sweep: every layer
    // Prefetch loads are UNCONDITIONAL (row and rank clamped into the pair's own records): a predicated load
    // makes the compiler merge old and new value with a move right behind the load, which waits for it --
    // measured as one exposed L2 round trip per layer (740 cycles per layer with the SM to itself).
    const int last_row = lq > 0 ? lq - 1 : 0, my_rank = ranked ? sub : K - 1;
    {
        Rec pre[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) pre[u] = load_rec(&rec[min(u, last_row) * K + my_rank]);
        for (int q0 = 0; q0 < lq_max; q0 += kAhead) {
#pragma unroll
            for (int u = 0; u < kAhead; ++u) {
                const int q = q0 + u;              // layers past lq_max only run predicated-off work
                const bool mine_on = ranked && q < lq;
                const Rec cur = pre[u];
                pre[u] = load_rec(&rec[min(q + kAhead, last_row) * K + my_rank]);
                const int v = q * K + sub;
                float d = 0.0f; int sl = -1, g = 0;
                if (mine_on) {
                    relax_node<MaskT, K, GS, true>(cur.pred, (MaskT)0, cur.sim, q, step, wmask, win, d, sl, g);
                    if (cur.pred) { rec[v].dist = d; gen[v] = (uint16_t)g; }  // nodes without predecessors stay (0, self, gen 0)
                    slot[v] = (int8_t)sl;
                }
                // window slot q & wmask held layer q-depth, which nobody reads any more
                win[(q & wmask) * 8 + sub] = make_float2(d, __int_as_float(g));
                if (!kDeferMax) {
                    const uint32_t db = __float_as_uint(d);
                    const uint32_t lm = oct_max(db, kFullMask);  // dist >= +0: bits are ordered
                    const uint32_t at_max = (__ballot_sync(kFullMask, db == lm) >> (oct * 8)) & 0xFFu;
                    if (sub == 0 && q < lq) { lbest[q] = lm; lrank[q] = (uint8_t)at_max; }
                }
                __syncwarp();
            }
            // (maximum, ranks that attain it) of the kAhead layers just relaxed: lane `sub` takes layer q0 + sub straight
            // from the window -- no shuffles on the layer-to-layer chain
            if (kDeferMax) {
                const int q = q0 + sub;
                if (sub < kAhead && q < lq) layer_max<K>(win, (q & wmask) * 8, lbest[q], lrank[q]);
                __syncwarp();
            }
        }
    }
    VSC_CLK(0);
    int n_boxes = 0;
    bool ambiguous = false;
    bool searching = alive;   // this octet still looks for chains
    for (int round = 0; round <= b.max_path; ++round) {
        if (!__any_sync(kFullMask, searching)) break;
        // ---- end node: maximum distance; ties -> smallest Kahn generation.  One pass over the per-layer maxima in
        // shared memory; the generations (global memory) are only read when more than one node attains the maximum.
        uint32_t mine_max = 0; int q_first = -1, n_layers = 0;
        for (int q = sub; q < lq_max; q += 8) {
            if (q >= lq) break;
            const uint32_t x = lbest[q];
            if (x > mine_max) { mine_max = x; q_first = q; n_layers = 1; }
            else if (x == mine_max && x != 0u) ++n_layers;
        }
        const uint32_t mk = oct_max(mine_max, kFullMask);
        if (mk == 0u) searching = false;  // only zero-length paths left: networkx returns [source]
        const bool holds_max = searching && mine_max == mk;
        const int layers_at_max = oct_add(holds_max ? n_layers : 0, kFullMask);
        // Ties at the maximum: the smallest generation wins.  Several nodes of that generation (rare): topo_compare gives
        // networkx's order among them, or leaves the pair to the exact-order kernel.
        int bg = INT_MAX, bv = -1, cnt = 0;
        bool undecided = false;
        if (holds_max) {
            const uint32_t m0 = lrank[q_first];
            if (layers_at_max == 1 && (m0 & (m0 - 1)) == 0u) {   // a single node: no tie to break
                bv = q_first * K + __ffs((int)m0) - 1; bg = 0; cnt = 1;
            } else {
                for (int q = q_first; q < lq; q += 8) {
                    if (lbest[q] != mk) continue;
                    for (uint32_t m = lrank[q]; m; m &= m - 1) {
                        const int v = q * K + __ffs((int)m) - 1;
                        const int g = gen[v];
                        if (g < bg) { bg = g; bv = v; cnt = 1; }
                        else if (g == bg) ++cnt;
                    }
                }
            }
        }
        const int g_min = oct_min(bg, kFullMask);
        const bool finalist = searching && bg == g_min;
        const int tied = oct_add(finalist ? cnt : 0, kFullMask);
        if (finalist && cnt > 1) {   // this lane's own candidates of the winning generation, in networkx order
            for (int q = q_first; q < lq; q += 8) {
                if (lbest[q] != mk) continue;
                for (uint32_t m = lrank[q]; m; m &= m - 1) {
                    const int v = q * K + __ffs((int)m) - 1;
                    if (v != bv && gen[v] == g_min && topo_compare<MaskT, K, GS>(v, bv, g_min, step, rec, gen, undecided) < 0) bv = v;
                }
            }
        }
        int end = -1;
        if (__any_sync(kFullMask, tied > 1)) {   // the octet's first lane decides between the lanes' winners
            for (int k = 0; k < 8; ++k) {
                const int cand = __shfl_sync(kFullMask, finalist ? bv : -1, oct * 8 + k);
                if (sub == 0 && cand >= 0) {
                    if (end < 0) end = cand;
                    else if (topo_compare<MaskT, K, GS>(cand, end, g_min, step, rec, gen, undecided) < 0) end = cand;
                }
            }
            undecided = (__ballot_sync(kFullMask, undecided) & om) != 0;
            end = __shfl_sync(kFullMask, end, oct * 8);
        } else {
            const unsigned who = __ballot_sync(kFullMask, finalist) & om;
            end = __shfl_sync(kFullMask, bv, who ? __ffs(who) - 1 : lane);
        }
        // ---- ties topo_compare could not order (a node of generation >= 2 with several parents in the generation
        // before): the octet computes the last parent of every possible ancestor of the tied nodes -- rows
        // [first tied row - (step-1) * generation, last tied row], ascending, so that the order of a node's parents
        // only needs entries that are already there -- and orders the tied nodes exactly.  Rare (about one pair in
        // 10^4 on the bench workload); the other octets of the warp idle meanwhile.
        if (__any_sync(kFullMask, undecided && searching)) {
            const bool fix = undecided && searching;
            uint8_t *last_parent = w.last_parent + nb;
            int c_lo = INT_MAX, c_hi = -1;
            if (fix) {
                for (int q = sub; q < lq; q += 8)
                    if (lbest[q] == mk) { c_lo = min(c_lo, q); c_hi = max(c_hi, q); }
            }
            c_lo = oct_min(c_lo, kFullMask);
            c_hi = -oct_min(-c_hi, kFullMask);
            int q = fix ? max(0, c_lo - (step - 1) * g_min) : 0;
            const int q_end = fix ? c_hi : -1;
            while (__any_sync(kFullMask, q <= q_end)) {
                if (q <= q_end && ranked) {
                    const int v = q * K + sub;
                    const MaskT pm = load_rec(&rec[v]).pred;
                    if (pm) {
                        const int g = gen[v];
                        int bu = -1, bsl = 0;
                        for (MaskT m = pm; m; m &= m - 1) {
                            const int sl = sizeof(MaskT) == 8 ? __ffsll((long long)m) - 1 : __ffs((int)m) - 1;
                            const int u = parent_of_slot<MaskT, K, GS>(v, sl, step);
                            if (gen[u] != g - 1) continue;
                            if (bu < 0 || topo_compare_exact<MaskT, K, GS>(u, bu, g - 1, step, last_parent) > 0) { bu = u; bsl = sl; }
                        }
                        __stcg(&last_parent[v], (uint8_t)bsl);
                    }
                }
                if (q <= q_end) ++q;
                __syncwarp();   // the row's entries are visible to the octet before the next row reads them
            }
            int e2 = -1;
            if (fix && sub == 0) {
                for (int qq = c_lo; qq <= c_hi; ++qq) {
                    if (lbest[qq] != mk) continue;
                    for (uint32_t m = lrank[qq]; m; m &= m - 1) {
                        const int v = qq * K + __ffs((int)m) - 1;
                        if (gen[v] != g_min) continue;
                        if (e2 < 0 || topo_compare_exact<MaskT, K, GS>(v, e2, g_min, step, last_parent) < 0) e2 = v;
                    }
                }
            }
            e2 = __shfl_sync(kFullMask, e2, oct * 8);
            if (fix) { end = e2; undecided = false; }
        }
        if (undecided && searching) { ambiguous = true; searching = false; }   // unreachable; kept as a guard

        VSC_CLK(1);
        // ---- walk the chain back through the best-predecessor slots (shared memory only, one lane per octet)
        int len = 0;
        if (sub == 0 && searching) {
            for (int v = end;;) {
                chain[len++] = (uint16_t)v;
                const int sl = slot[v];
                if (sl < 0) break;
                const int grp = sl / GS;
                v = (v / K - (step - 1 - grp)) * K + (sl - grp * GS);
            }
        }
        len = __shfl_sync(kFullMask, len, oct * 8);
        __syncwarp();   // chain[] is visible to the octet
        // ---- all eight lanes: mark the chain's edges spent (fire-and-forget L2 atomics) and fetch the similarities;
        // the float32 score is summed in path order (first node first) by shuffling the values through the octet.
        VSC_CLK(2);
        float score = 0.0f;
        const int len_max = __reduce_max_sync(kFullMask, len);
        for (int i0 = 0; i0 < len_max; i0 += 8) {
            const int i = len - 1 - (i0 + sub);   // path position i0+sub counted from the first node
            float s = 0.0f;
            if (i >= 0) {
                const int v = chain[i];
                const int sl = slot[v];
                if (sl >= 0) {
                    if (sizeof(MaskT) == 8) atomicOr(reinterpret_cast<unsigned long long *>(&rec[v].zero), 1ull << sl);
                    else atomicOr(reinterpret_cast<unsigned int *>(&rec[v].zero), 1u << sl);
                }
                s = __ldcg(&rec[v].sim);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float sk = __shfl_sync(kFullMask, s, oct * 8 + k);
                if (i0 + k < len) score += sk;
            }
        }
        VSC_CLK(3);
        // ---- box of the chain and the filter (one lane per octet)
        int q_first_dst = 0, q_last = -1;
        if (sub == 0 && searching) {
            const int first = chain[len - 1], last = chain[0];
            q_first_dst = (len >= 2 ? (int)chain[len - 2] : last) / K;
            q_last = last / K;
            int q_lo = 0, q_hi = 0, r_lo = 0, r_hi = 0;
            if (score > 0.0f) {  // q and (by C2) r increase strictly along a chain
                q_lo = first / K; q_hi = q_last;
                r_lo = ref_of[first] & kRefMask; r_hi = ref_of[last] & kRefMask;
            }
            const double mean_extent = (double)(r_hi - r_lo + q_hi - q_lo) / 2.0;
            double worst = 0.0;
            for (int k = 0; k < n_boxes; ++k) {
                const int4 g = boxes[k];
                long long ww = (long long)min(q_hi, g.z) - max(q_lo, g.x) + 1;
                long long hh = (long long)min(r_hi, g.w) - max(r_lo, g.y) + 1;
                ww = ww < 0 ? 0 : ww; hh = hh < 0 ? 0 : hh;
                const long long inter = ww * hh;
                double iou = 0.0;   // disjoint boxes (the usual case) skip the float64 division
                if (inter != 0) {
                    const long long a1 = (long long)(q_hi - q_lo + 1) * (r_hi - r_lo + 1);
                    const long long a2 = (long long)(g.z - g.x + 1) * (g.w - g.y + 1);
                    iou = (double)inter / (double)(a1 + a2 - inter);
                }
                if (k == 0 || iou > worst) worst = iou;
            }
            const int shorter = min(r_hi - r_lo, q_hi - q_lo);
            if (mean_extent != 0.0 && __fdiv_rn(score, (float)mean_extent) > b.min_sim &&
                (double)shorter > b.min_length && worst < b.max_iou) {
                boxes[n_boxes] = make_int4(q_lo, r_lo, q_hi, r_hi);
                ++n_boxes;
            }
        }
        q_first_dst = __shfl_sync(kFullMask, q_first_dst, oct * 8);
        q_last = __shfl_sync(kFullMask, q_last, oct * 8);
        n_boxes = __shfl_sync(kFullMask, n_boxes, oct * 8);
        __syncwarp();
        VSC_CLK(4);
        if (round == b.max_path) break;

        // ---- incremental sweep: layers before the first zeroed edge keep their distances, and the wave dies
        // `step-1` layers after the last distance that changed.  Each octet walks its own layer range; the warp
        // loops until the longest range is done.
        const bool sweeping = searching;
        if (sweeping) {
            for (int o = 1; o < step; ++o) {  // refill the window behind the start layer
                const int qq = q_first_dst - o;
                if (qq >= 0) win[(qq & wmask) * 8 + sub].x = ranked ? __ldcg(&rec[qq * K + sub].dist) : 0.0f;
            }
        }
        int last_changed = INT_MIN / 2;
        int q = q_first_dst;
        Rec pre[kAhead];
#pragma unroll
        for (int u = 0; u < kAhead; ++u) pre[u] = load_rec(&rec[min(q + u, last_row) * K + my_rank]);
        __syncwarp();
        for (;;) {
            // the exit test runs once per kAhead steps (an early exit inside the unrolled body brings back the
            // register moves behind the prefetch loads); the few extra steps are predicated off
            if (!__any_sync(kFullMask, sweeping && q < lq && (q <= q_last || q <= last_changed + step - 1))) break;
            const int q_group = q;
#pragma unroll
            for (int u = 0; u < kAhead; ++u) {
                // once `on` turns false for an octet it stays false (nothing can update last_changed), so the
                // register rotation only has to be right while the octet is running
                const bool on = sweeping && q < lq && (q <= q_last || q <= last_changed + step - 1);
                const Rec cur = pre[u];
                pre[u] = load_rec(&rec[min(q + kAhead, last_row) * K + my_rank]);
                const int v = q * K + sub;
                float d = 0.0f; int sl = -1, g = 0;
                bool changed = false;
                if (ranked && on) {
                    relax_node<MaskT, K, GS, false>(cur.pred, cur.zero, cur.sim, q, step, wmask, win, d, sl, g);
                    changed = __float_as_uint(d) != __float_as_uint(cur.dist);
                    if (changed) rec[v].dist = d;
                    slot[v] = (int8_t)sl;
                }
                if (on) win[(q & wmask) * 8 + sub].x = d;
                if (__ballot_sync(kFullMask, changed) & om) last_changed = q;
                if (!kDeferMax) {
                    const uint32_t db = __float_as_uint(d);
                    const uint32_t lm = oct_max(db, kFullMask);
                    const uint32_t at_max = (__ballot_sync(kFullMask, db == lm) >> (oct * 8)) & 0xFFu;
                    if (sub == 0 && on) { lbest[q] = lm; lrank[q] = (uint8_t)at_max; }
                }
                if (on) ++q;
                VSC_CLK_ADD(6, 1);
                __syncwarp();
            }
            if (kDeferMax) {   // the layers [q_group, q) this octet has just relaxed
                const int ql = q_group + sub;
                if (sub < kAhead && ql < q) layer_max<K>(win, (ql & wmask) * 8, lbest[ql], lrank[ql]);
                __syncwarp();
            }
        }
        VSC_CLK(5);
    }
    VSC_CLK_FLUSH();
    if (sub == 0 && alive) {
        if (ambiguous) {
            w.skip[pair] = 1;
            if (w.exact_list) w.exact_list[atomicAdd(w.exact_count, 1)] = pair;
            else out.list[atomicAdd(out.count, 1)] = pair;
        } else {
            b.n_boxes[pair] = n_boxes;
            if (b.status) b.status[pair] = 0;
        }
    }
}

template <typename MaskT, int K, int GS>
cudaError_t launch_dp_k(const Batch &b, const Workspace &w, const WorkList &out, int grid, size_t smem,
                        cudaStream_t stream, int pairs_per_warp) {
    cudaError_t e = cudaFuncSetAttribute(tn_dp_kernel<MaskT, K, GS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(tn_dp_kernel<MaskT, K, GS>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    tn_dp_kernel<MaskT, K, GS><<<grid, kT2Threads, smem, stream>>>(b, w, out, pairs_per_warp);
    return cudaGetLastError();
}
template <typename MaskT, int GS8>
cudaError_t launch_dp(const Batch &b, const Workspace &w, const WorkList &out, int grid, size_t smem,
                      cudaStream_t stream, int ppw) {
    switch (b.topk) {
        case 1: return launch_dp_k<MaskT, 1, GS8 ? 8 : 1>(b, w, out, grid, smem, stream, ppw);
        case 2: return launch_dp_k<MaskT, 2, GS8 ? 8 : 2>(b, w, out, grid, smem, stream, ppw);
        case 3: return launch_dp_k<MaskT, 3, GS8 ? 8 : 3>(b, w, out, grid, smem, stream, ppw);
        case 4: return launch_dp_k<MaskT, 4, GS8 ? 8 : 4>(b, w, out, grid, smem, stream, ppw);
        case 5: return launch_dp_k<MaskT, 5, GS8 ? 8 : 5>(b, w, out, grid, smem, stream, ppw);
        case 6: return launch_dp_k<MaskT, 6, GS8 ? 8 : 6>(b, w, out, grid, smem, stream, ppw);
        case 7: return launch_dp_k<MaskT, 7, GS8 ? 8 : 7>(b, w, out, grid, smem, stream, ppw);
        default: return launch_dp_k<MaskT, 8, 8>(b, w, out, grid, smem, stream, ppw);
    }
}

// ------------------------------------------------------------------ T3: MaxSim per box
// max(sims[q0:q1, r0:r1]) per kept box (exclusive upper bounds, localization.py:91), mostly WITHOUT reading the matrix again:
// the K nodes of a row are its K largest similarities, so (a) if one of them lies in [r0, r1) the row's maximum inside the box
// is the best such node, and (b) if none does, every element of the row inside the box is <= the row's K-th best.  Pass 1 takes
// M = the best in-range node over the box's rows from the node records (a few hundred bytes per row instead of the 4*(r1-r0)
// bytes of the row segment, which were evicted from L2 long ago); pass 2 scans the matrix only for rows of kind (b) whose K-th
// best exceeds M -- on planted copies none, on noise boxes a few.  Bit-exact: a maximum of float32 values has no rounding.
__global__ void __launch_bounds__(128) tn_maxsim_kernel(const Batch b, const Workspace w) {
    const int pair = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (pair >= b.n_pairs || w.skip[pair]) return;
    const int box_cap = b.max_path + 1, lr = b.lr[pair], K = b.topk;
    const float *sims = b.sims + b.off[pair];
    const size_t nb = (size_t)pair * b.max_nodes;
    const uint16_t *ref_of = w.ref_of + nb;
    const unsigned char *rec = static_cast<const unsigned char *>(w.rec) + nb * w.rec_bytes + w.sim_off;
    const int nbx = b.n_boxes[pair];
    for (int k = 0; k < nbx; ++k) {
        const int32_t *g = b.boxes + ((size_t)pair * box_cap + k) * 4;
        const int q0 = g[0], r0 = g[1], q1 = g[2], r1 = g[3];
        float best = -INFINITY;
        if (q1 > q0 && r1 > r0) {
            for (int base = q0; base < q1; base += 32) {
                const int row = base + lane;
                if (row < q1) {
                    for (int i = 0; i < K; ++i) {
                        const int r = ref_of[row * K + i] & kRefMask;
                        if (r >= r0 && r < r1)
                            best = fmaxf(best, *reinterpret_cast<const float *>(rec + (size_t)(row * K + i) * w.rec_bytes));
                    }
                }
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) best = fmaxf(best, __shfl_xor_sync(kFullMask, best, d));
            const float m_nodes = best;
            for (int base = q0; base < q1; base += 32) {
                const int row = base + lane;
                bool scan = false;
                if (row < q1) {
                    bool inside = false;
                    float low = INFINITY;
                    for (int i = 0; i < K; ++i) {
                        const int r = ref_of[row * K + i] & kRefMask;
                        inside = inside || (r >= r0 && r < r1);
                        low = fminf(low, *reinterpret_cast<const float *>(rec + (size_t)(row * K + i) * w.rec_bytes));
                    }
                    scan = !inside && low > m_nodes;
                }
                unsigned todo = __ballot_sync(kFullMask, scan);
                while (todo) {   // the warp scans the row segment together
                    const int rr = base + __ffs((int)todo) - 1;
                    todo &= todo - 1;
                    const float *seg = sims + (size_t)rr * lr;
                    float m = -INFINITY;
                    for (int c = r0 + lane; c < r1; c += 32) m = fmaxf(m, seg[c]);
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) m = fmaxf(m, __shfl_xor_sync(kFullMask, m, d));
                    best = fmaxf(best, m);
                }
            }
        }
        if (lane == 0) b.box_maxsim[(size_t)pair * box_cap + k] = best;
    }
}

__global__ void iota_kernel(int32_t *count, int32_t *list, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = i;
    if (i == 0) *count = n;
}

// Warps per T1 CTA: as many as shared memory holds (one CTA per SM), at most kT1MaxWarps.
int t1_warps(int max_lr) {
    const size_t fit = (size_t)(227 * 1024) / t1_warp_bytes(max_lr);
    return (int)(fit < (size_t)kT1MaxWarps ? fit : (size_t)kT1MaxWarps);
}

template <int K>
int launch_topk(const T1Args &a, int grid, int warps, cudaStream_t stream) {
    const size_t smem = (size_t)warps * a.warp_bytes;
    VSC_CUDA_CHECK(cudaFuncSetAttribute(tn_topk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VSC_CUDA_CHECK(cudaFuncSetAttribute(tn_topk_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    tn_topk_kernel<K><<<grid, warps * 32, smem, stream>>>(a);
    VSC_CUDA_CHECK(cudaGetLastError());
    return VSC_OK;
}

}  // namespace

namespace vsc {
namespace tn {

// Optional per-stage device timing of the last pipeline call (vsc_tn_set_profiling).
static bool g_profile = false;
static cudaEvent_t g_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
static bool g_ev_valid = false;
static void mark(int i, cudaStream_t stream) {
    if (!g_profile) return;
    if (!g_ev[i]) cudaEventCreate(&g_ev[i]);
    cudaEventRecord(g_ev[i], stream);
    if (i == 4) g_ev_valid = true;
}

// what the graph stage (edges + sweeps) needs, whatever produced the node records
bool graph_supported(const Batch &b) {
    if (b.topk < 1 || b.topk > kMaxTop) return false;
    if ((b.step - 1) * 8 > 32 && (b.step - 1) * b.topk > 64) return false;   // predecessor mask width
    if ((reinterpret_cast<uintptr_t>(b.boxes) & 15u) != 0) return false;
    if (b.max_nodes > 65535 || b.max_lr > kRefMask) return false;
    if (t2_pair_bytes(b.max_nodes, b.max_lq, b.step) * kT2Warps * 4 > 200 * 1024) return false;
    return true;
}

bool pipeline_supported(const Batch &b) {
    if (!graph_supported(b)) return false;
    if (b.max_lr > vsc::kBlockCols * vsc::kMaxRowBlocks || b.max_lr < b.topk) return false;
    if ((reinterpret_cast<uintptr_t>(b.sims) & 15u) != 0) return false;
    if ((long long)b.n_pairs * ((b.max_lq + kTileRows - 1) / kTileRows) >= (1ll << 31) - 65536) return false;  // tile ids
    return t1_warps(b.max_lr) >= 1;
}

// One stream-ordered allocation holding the node records, carved by alignment; flags and counters zeroed.
int workspace_alloc(const Batch &b, Workspace *w, void **base_out, cudaStream_t stream) {
    const bool wide = (b.step - 1) * 8 > 32;   // fast variant: 32-bit masks with a group stride of 8 slots
    const size_t P = (size_t)b.n_pairs, N = (size_t)b.max_nodes;
    size_t sz = 0;
    auto take = [&](size_t bytes) { size_t at = sz; sz += (bytes + 255) / 256 * 256; return at; };
    const size_t rec_bytes = wide ? sizeof(NodeRec<uint64_t>) : sizeof(NodeRec<uint32_t>);
    const size_t o_rec = take(P * N * rec_bytes);
    const size_t o_ref = take(P * N * 2), o_gen = take(P * N * 2), o_last = take(P * N);
    const size_t o_skip = take(P * 4), o_cursor = take(4);
    unsigned char *base = nullptr;
    VSC_CUDA_CHECK(cudaMallocAsync(&base, sz, stream));
    *base_out = base;
    w->rec = base + o_rec; w->rec_bytes = (int)rec_bytes; w->sim_off = wide ? 16 : 8;
    w->ref_of = reinterpret_cast<uint16_t *>(base + o_ref); w->gen = reinterpret_cast<uint16_t *>(base + o_gen);
    w->last_parent = base + o_last;   // only written / read by the rare exact tie resolution: never initialised
    w->skip = reinterpret_cast<int32_t *>(base + o_skip); w->cursor = reinterpret_cast<int32_t *>(base + o_cursor);
    w->exact_count = nullptr; w->exact_list = nullptr;
    VSC_CUDA_CHECK(cudaMemsetAsync(w->gen, 0, P * N * 2, stream));
    VSC_CUDA_CHECK(cudaMemsetAsync(w->skip, 0, (o_cursor - o_skip) + 4, stream));
    return VSC_OK;
}

// T1: row top-K of similarity matrices resident in memory
static int launch_row_topk(const Batch &b, const Workspace &w, const WorkList &out, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    T1Args a; a.b = b; a.w = w; a.out = out;
    a.tiles_per_pair = (b.max_lq + kTileRows - 1) / kTileRows;
    a.n_tiles = b.n_pairs * a.tiles_per_pair;
    a.warp_bytes = (int)t1_warp_bytes(b.max_lr);
    const int warps = t1_warps(b.max_lr);
    const int grid = sms;  // persistent: one CTA per SM
    int rc;
    switch (b.topk) {
        case 1: rc = launch_topk<1>(a, grid, warps, stream); break;
        case 2: rc = launch_topk<2>(a, grid, warps, stream); break;
        case 3: rc = launch_topk<3>(a, grid, warps, stream); break;
        case 4: rc = launch_topk<4>(a, grid, warps, stream); break;
        case 5: rc = launch_topk<5>(a, grid, warps, stream); break;
        case 6: rc = launch_topk<6>(a, grid, warps, stream); break;
        case 7: rc = launch_topk<7>(a, grid, warps, stream); break;
        default: rc = launch_topk<8>(a, grid, warps, stream); break;
    }
    vsc::count_launch();
    return rc;
}

// edges + longest-path sweeps (+ MaxSim when the similarity matrices are in memory) on the node records of `w`
// Graph-stage variant: 0 = layer-by-layer kernels (default: fewer warp instructions at batch sizes that fill the GPU),
// 1 = compact graph by Kahn generation (tn_graph.cu; profiles/r02_tn_graph_variants.md).  VSC_TN_GRAPH=compact or
// vsc_tn_set_graph_variant(1) selects the latter.
// pairs per SM up to which the DP runs one / two pairs per warp (see launch_graph)
// (measured, 300x300 pairs: one pair per warp wins up to ~4 pairs per SM, two up to ~8, four beyond)
static int g_dp_ppw1_per_sm = 4, g_dp_ppw2_per_sm = 8;
static int g_dp_ppw_forced = [] { const char *e = getenv("VSC_DP_PPW"); return e ? atoi(e) : 0; }();
static int g_graph_variant = [] { const char *e = getenv("VSC_TN_GRAPH"); return e && e[0] == 'c' ? 1 : 0; }();
static bool use_graph_v2() { return g_graph_variant == 1; }

static int launch_graph(const Batch &b, const Workspace &w, const WorkList &out, cudaStream_t stream) {
    const bool wide = (b.step - 1) * 8 > 32;
    int rc = VSC_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == VSC_OK) { vsc::set_error("%s: %s", what, cudaGetErrorString(e)); rc = VSC_ERR_CUDA; }
    };
    if (use_graph_v2() && graph_v2_supported(b)) {
        unsigned char *graphs = nullptr;
        VSC_CUDA_CHECK(cudaMallocAsync(&graphs, graph_v2_scratch_bytes(b), stream));
        rc = launch_graph_v2(b, w, out, graphs, stream, [](cudaStream_t s) { mark(2, s); });
        cudaFreeAsync(graphs, stream);
    } else {
    {
        const long long threads = (long long)b.n_pairs * b.max_lq;
        const int grid = (int)((threads + 255) / 256);
        if (wide) launch_edges<uint64_t, 0>(b, w, grid, stream);
        else launch_edges<uint32_t, 1>(b, w, grid, stream);
        fail(cudaGetLastError(), "tn_edges_kernel");
        vsc::count_launch();
    }
    mark(2, stream);
    if (rc == VSC_OK) {
        // pairs per warp: four when the batch fills the GPU (fewest warp instructions), fewer for small batches (a warp
        // runs as long as its slowest pair; measured in profiles/r02_tn_summary.md).  VSC_DP_PPW / vsc_tn_set_dp_pairs_per_warp override.
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        int ppw = b.n_pairs <= sms * g_dp_ppw1_per_sm ? 1 : (b.n_pairs <= sms * g_dp_ppw2_per_sm ? 2 : 4);
        if (g_dp_ppw_forced == 1 || g_dp_ppw_forced == 2 || g_dp_ppw_forced == 4) ppw = g_dp_ppw_forced;
        const int pairs_per_cta = kT2Warps * ppw;
        const int grid = (b.n_pairs + pairs_per_cta - 1) / pairs_per_cta;
        const size_t smem = t2_pair_bytes(b.max_nodes, b.max_lq, b.step) * kT2Warps * 4;
        fail(wide ? launch_dp<uint64_t, 0>(b, w, out, grid, smem, stream, ppw)
                  : launch_dp<uint32_t, 1>(b, w, out, grid, smem, stream, ppw), "tn_dp_kernel");
        fail(cudaGetLastError(), "tn_dp_kernel");
        vsc::count_launch();
    }
    }
    mark(3, stream);
    if (rc == VSC_OK && b.box_maxsim && b.sims) {
        tn_maxsim_kernel<<<(b.n_pairs + 3) / 4, 128, 0, stream>>>(b, w);
        fail(cudaGetLastError(), "tn_maxsim_kernel");
        vsc::count_launch();
    }
    mark(4, stream);
    return rc;
}

int launch_pipeline(const Batch &b, const Workspace &w, const WorkList &out, cudaStream_t stream) {
    mark(0, stream);
    int rc = launch_row_topk(b, w, out, stream);
    mark(1, stream);
    if (rc == VSC_OK) rc = launch_graph(b, w, out, stream);
    return rc;
}

int launch_pipeline_from_features(const PairOperands &op, const Batch &b, const Workspace &w, const WorkList &out,
                                  cudaStream_t stream) {
    mark(0, stream);
    int rc = launch_pair_topk(op, b, w, out, const_cast<float *>(b.sims), b.off, 0, stream);
    mark(1, stream);
    if (rc == VSC_OK) rc = launch_graph(b, w, out, stream);
    return rc;
}

}  // namespace tn
}  // namespace vsc

extern "C" int vsc_tn_set_graph_variant(int variant) {
    if (variant != 0 && variant != 1) { vsc::set_error("vsc_tn_set_graph_variant: 0 (layers) or 1 (compact)"); return VSC_ERR_INVALID; }
    vsc::tn::g_graph_variant = variant;
    return VSC_OK;
}

extern "C" int vsc_tn_set_dp_pairs_per_warp(int pairs) {
    if (pairs != 0 && pairs != 1 && pairs != 2 && pairs != 4) { vsc::set_error("vsc_tn_set_dp_pairs_per_warp: 0 (auto), 1, 2 or 4"); return VSC_ERR_INVALID; }
    vsc::tn::g_dp_ppw_forced = pairs;
    return VSC_OK;
}

extern "C" int vsc_tn_set_profiling(int on) {
    vsc::tn::g_profile = on != 0;
    vsc::tn::g_ev_valid = false;
    return VSC_OK;
}
// Device time (ms) of the stages of the most recent fast-pipeline call: row top-K, edges,
// longest-path sweeps, MaxSim.  The caller must have synchronised the stream.
extern "C" int vsc_tn_last_stage_ms(float *out4) {
    using namespace vsc::tn;
    if (!g_profile || !g_ev_valid) { vsc::set_error("vsc_tn_last_stage_ms: no profiled call"); return VSC_ERR_INVALID; }
    for (int i = 0; i < 4; ++i) VSC_CUDA_CHECK(cudaEventElapsedTime(&out4[i], g_ev[i], g_ev[i + 1]));
    return VSC_OK;
}

// Development aid: DP phase clocks (all zero unless built with -DVSC_TN_COUNTERS), summed over warps:
// cycles in {first sweep, end-node search, chain walk, zero + score, box filter, incremental sweeps},
// incremental layer steps, warps.
extern "C" int vsc_tn_debug_counters(unsigned long long *out8) {
    VSC_CUDA_CHECK(cudaMemcpyFromSymbol(out8, g_dp_counters, sizeof(unsigned long long) * 8));
    unsigned long long zero[8] = {};
    VSC_CUDA_CHECK(cudaMemcpyToSymbol(g_dp_counters, zero, sizeof zero));
    return VSC_OK;
}

namespace {

__global__ void strided_offsets_kernel(int64_t *off, int n, int64_t stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) off[i] = (int64_t)i * stride;
}

int check_params(const vsc_tn_params *p, int32_t n_pairs, int32_t max_lq, int32_t max_lr, const char *who) {
    if (!p || n_pairs < 0) { vsc::set_error("%s: bad arguments", who); return VSC_ERR_INVALID; }
    if (p->tn_top_k < 1 || p->tn_top_k > kMaxTop || p->tn_max_step < 1 || p->tn_max_step > 31 ||
        (p->tn_max_step - 1) * p->tn_top_k > 64 || p->max_path < 0 || p->max_path + 1 > kMaxBoxes) {
        vsc::set_error("%s: unsupported parameters (need tn_top_k<=%d, "
                       "(tn_max_step-1)*tn_top_k<=64, max_path<%d)", who, kMaxTop, kMaxBoxes);
        return VSC_ERR_INVALID;
    }
    if (max_lr > 65535 || max_lq < 0 || max_lr < 0) {
        vsc::set_error("%s: max_lr %d out of range (<= 65535)", who, max_lr);
        return VSC_ERR_INVALID;
    }
    return VSC_OK;
}

void fill_batch(vsc::tn::Batch &b, const vsc_tn_params *p, int32_t max_lq, int32_t max_lr) {
    b.step = p->tn_max_step; b.topk = p->tn_top_k; b.max_path = p->max_path;
    b.min_sim = p->min_sim; b.min_length = p->min_length; b.max_iou = p->max_iou;
    b.max_lq = max_lq > 0 ? max_lq : 1;
    b.max_lr = max_lr;
    b.max_nodes = b.max_lq * (p->tn_top_k < max_lr ? p->tn_top_k : (max_lr > 0 ? max_lr : 1));
}

// The whole alignment of a batch.  `op` != null: the similarities come from descriptor panels (b.sims is then either
// null or the buffer the matrices are to be written to); `features_direct`: the row top-K may run straight out of
// tensor memory (every pair has at least topk columns).
int run_tn(vsc::tn::Batch b, const vsc::tn::PairOperands *op, bool features_direct, int force_exact_order,
           cudaStream_t stream) {
    using namespace vsc::tn;
    if (b.max_nodes > 65535) {
        vsc::set_error("vcsl_tn_batch: %d graph nodes exceed the 16-bit node index", b.max_nodes);
        return VSC_ERR_CAPACITY;
    }
    // two device-side work lists: [count, ids...]
    int32_t *lists = nullptr;
    const size_t list_len = (size_t)b.n_pairs + 1;
    VSC_CUDA_CHECK(cudaMallocAsync(&lists, sizeof(int32_t) * 2 * list_len, stream));
    WorkList A{lists, lists + 1}, B{lists + list_len, lists + list_len + 1};
    int rc = VSC_OK;
    void *ws_base = nullptr;
    cudaError_t e = cudaMemsetAsync(lists, 0, sizeof(int32_t), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(lists + list_len, 0, sizeof(int32_t), stream);
    if (e != cudaSuccess) { vsc::set_error("memset: %s", cudaGetErrorString(e)); rc = VSC_ERR_CUDA; }
    if (rc == VSC_OK) {
        Workspace w;
        if (force_exact_order) {
            iota_kernel<<<(b.n_pairs + 255) / 256, 256, 0, stream>>>(B.count, B.list, b.n_pairs);
            vsc::count_launch();
        } else if (op && features_direct && graph_supported(b) && pair_topk_supported(b)) {
            rc = workspace_alloc(b, &w, &ws_base, stream);
            w.exact_count = B.count; w.exact_list = B.list;
            if (rc == VSC_OK) rc = launch_pipeline_from_features(*op, b, w, A, stream);
            if (rc == VSC_OK) {   // pairs handed back read their node records instead of a similarity matrix
                Batch bn = b;
                bn.node_ref = w.ref_of; bn.node_rec = w.rec; bn.node_rec_bytes = w.rec_bytes; bn.node_sim_off = w.sim_off;
                rc = launch_fused(bn, false, &A, &B, 2, stream);
                if (rc == VSC_OK) rc = launch_fused(bn, true, &B, nullptr, 1, stream);
            }
            if (ws_base) cudaFreeAsync(ws_base, stream);
            cudaFreeAsync(lists, stream);
            return rc;
        } else if (pipeline_supported(b)) {
            rc = workspace_alloc(b, &w, &ws_base, stream);
            w.exact_count = B.count; w.exact_list = B.list;
            if (rc == VSC_OK) rc = launch_pipeline(b, w, A, stream);
            if (rc == VSC_OK) rc = launch_fused(b, false, &A, &B, 2, stream);
        } else {
            rc = launch_fused(b, false, nullptr, &B, 2, stream);
        }
    }
    if (rc == VSC_OK) rc = launch_fused(b, true, &B, nullptr, 1, stream);
    if (ws_base) cudaFreeAsync(ws_base, stream);
    cudaFreeAsync(lists, stream);
    return rc;
}

}  // namespace

extern "C" int vcsl_tn_batch(const float *d_sims, const int64_t *d_off, const int32_t *d_lq,
                             const int32_t *d_lr, int32_t n_pairs, int32_t max_lq, int32_t max_lr,
                             const vsc_tn_params *p, int32_t *d_boxes, int32_t *d_n_boxes,
                             float *d_box_maxsim, int32_t *d_status, int32_t force_exact_order,
                             vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = check_params(p, n_pairs, max_lq, max_lr, "vcsl_tn_batch");
    if (rc != VSC_OK) return rc;
    if (n_pairs == 0) return VSC_OK;
    vsc::keep_pool_cached();
    if (!d_sims || !d_off || !d_lq || !d_lr || !d_boxes || !d_n_boxes) {
        vsc::set_error("vcsl_tn_batch: null device pointer"); return VSC_ERR_INVALID;
    }
    vsc::tn::Batch b = {};
    b.sims = d_sims; b.off = d_off; b.lq = d_lq; b.lr = d_lr; b.n_pairs = n_pairs;
    b.boxes = d_boxes; b.n_boxes = d_n_boxes; b.box_maxsim = d_box_maxsim; b.status = d_status;
    fill_batch(b, p, max_lq, max_lr);
    return run_tn(b, nullptr, false, force_exact_order, stream);
}

extern "C" int vcsl_tn_batch_from_features(const void *d_q_panel, int64_t q_rows, const void *d_r_panel, int64_t r_rows,
                                           int32_t k, const int32_t *d_q_start, const int32_t *d_lq,
                                           const int32_t *d_r_start, const int32_t *d_lr, int32_t n_pairs,
                                           int32_t max_lq, int32_t max_lr, int32_t min_lr, float similarity_bias,
                                           const vsc_tn_params *p, float *d_sims_out, const int64_t *d_off,
                                           int32_t *d_boxes, int32_t *d_n_boxes, float *d_box_maxsim, int32_t *d_status,
                                           int32_t force_exact_order, const vsc_gemm_format *fmt, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    int rc = check_params(p, n_pairs, max_lq, max_lr, "vcsl_tn_batch_from_features");
    if (rc != VSC_OK) return rc;
    if (n_pairs == 0) return VSC_OK;
    vsc::keep_pool_cached();
    if (!d_q_panel || !d_r_panel || !d_q_start || !d_lq || !d_r_start || !d_lr || !d_boxes || !d_n_boxes ||
        (d_sims_out && !d_off)) {
        vsc::set_error("vcsl_tn_batch_from_features: null device pointer"); return VSC_ERR_INVALID;
    }
    vsc::tn::Batch b = {};
    b.lq = d_lq; b.lr = d_lr; b.n_pairs = n_pairs;
    b.boxes = d_boxes; b.n_boxes = d_n_boxes; b.box_maxsim = d_box_maxsim; b.status = d_status;
    fill_batch(b, p, max_lq, max_lr);
    vsc::tn::PairOperands op = {d_q_panel, d_r_panel, q_rows, r_rows, k, d_q_start, d_r_start, similarity_bias, 0, 0, 0, nullptr};
    if (fmt) { op.ab_f16 = fmt->ab_f16; op.ldq = fmt->lda; op.ldr = fmt->ldb; op.out_scale = fmt->d_out_scale; }

    // The row top-K runs out of tensor memory when every pair has at least tn_top_k columns and at most 512; the
    // similarity matrices are only written when the caller wants them back or a MaxSim score has to read them.
    const bool direct = !force_exact_order && min_lr >= p->tn_top_k && vsc::tn::graph_supported(b) &&
                        vsc::tn::pair_topk_supported(b);
    const bool need_sims = d_sims_out || d_box_maxsim || !direct;
    float *sims = d_sims_out;
    int64_t *off_tmp = nullptr;
    float *sims_tmp = nullptr;
    if (need_sims && !sims) {   // scratch matrices: pair p at p * stride
        const int64_t stride = (((int64_t)b.max_lq * (max_lr > 0 ? max_lr : 1)) + 3) & ~(int64_t)3;
        VSC_CUDA_CHECK(cudaMallocAsync(&sims_tmp, sizeof(float) * (size_t)stride * n_pairs + 16, stream));
        VSC_CUDA_CHECK(cudaMallocAsync(&off_tmp, sizeof(int64_t) * n_pairs, stream));
        strided_offsets_kernel<<<(n_pairs + 255) / 256, 256, 0, stream>>>(off_tmp, n_pairs, stride);
        vsc::count_launch();
        sims = sims_tmp; d_off = off_tmp;
    }
    b.sims = sims; b.off = d_off;
    if (direct) {
        rc = run_tn(b, &op, true, 0, stream);
    } else {
        rc = vsc_pair_similarity(d_q_panel, q_rows, d_r_panel, r_rows, k, d_q_start, d_lq, d_r_start, d_lr, n_pairs,
                                 max_lq, max_lr, similarity_bias, sims, d_off, fmt, stream_);
        if (rc == VSC_OK) rc = run_tn(b, nullptr, false, force_exact_order, stream);
    }
    if (sims_tmp) cudaFreeAsync(sims_tmp, stream);
    if (off_tmp) cudaFreeAsync(off_tmp, stream);
    return rc;
}
