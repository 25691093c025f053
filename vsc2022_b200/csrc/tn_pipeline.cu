// Temporal-network (TN) alignment, FAST PIPELINE for aligned rows (B200, sm_100a).
//
// Replaces vcsl.vta `tn` (alipay/VCSL @ c39269d5) behind vsc/baseline/localization.py:44-46,58.
// Contract: oracle/tn_networkx.py.  Four launches per batch, all on one stream:
//
//   T1  tn_topk_kernel   HBM-bound.  Persistent warps; each lane pulls ITS OWN row of a 32-row
//                        tile into shared memory with one cp.async.bulk (TMA bulk copy, mbarrier
//                        completion, two stages) and selects the exact top-K of the row
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

#include "row_select.cuh"
#include "tn_common.cuh"

namespace {

using vsc::kFullMask;
using vsc::tn::Batch;
using vsc::tn::WorkList;
using vsc::tn::kMaxBoxes;
using vsc::tn::kMaxTop;

// ------------------------------------------------------------------ workspace
struct Workspace {
    uint16_t *ref_of;   // [P][N]
    float *sim_of;      // [P][N]
    void *pred;         // [P][N] uint32 or uint64
    void *zero;         // [P][N]
    float *dist;        // [P][N]
    int8_t *slot;       // [P][N]
    uint16_t *gen;      // [P][N]
    uint16_t *chain;    // [P][max_lq]
    uint32_t *lbest;    // [P][max_lq]  max dist bits per row layer
    uint8_t *skip;      // [P] 1 = handed to the general kernel
    int32_t *cursor;    // T1 pair counter
};

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared; completion is signalled on `bar` (complete_tx::bytes).
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ T1: row top-K
constexpr int kT1Warps = 2;
constexpr int kT1Threads = kT1Warps * 32;
constexpr int kTileRows = 32;
constexpr int kStages = 2;

__host__ __device__ inline int tile_pitch(int max_lr) {  // words; (pitch/4) odd -> LDS.128 conflict-free
    int p = (max_lr + 3) & ~3;
    if (((p >> 2) & 1) == 0) p += 4;
    return p;
}
__host__ __device__ inline size_t t1_warp_bytes(int pitch) {
    return (size_t)kStages * kTileRows * pitch * 4            // tiles
           + (size_t)vsc::kMaxRowBlocks * 32 * 4              // block maxima
           + (size_t)vsc::kMaxCand * 32 * 8                   // candidates (value, column)
           + 64;                                              // mbarriers (+pad)
}

struct Cursor {
    int pair, row0, lq, lr;
    const float *base;
};

struct T1Args {
    Batch b;
    Workspace w;
    WorkList out;
    int pitch;
};

// Claim the next pair this warp can process; pairs it cannot take go to the general kernel.
__device__ inline bool claim_pair(const T1Args &a, int lane, Cursor &c) {
    for (;;) {
        int p = 0;
        if (lane == 0) p = atomicAdd(a.w.cursor, 1);
        p = __shfl_sync(kFullMask, p, 0);
        if (p >= a.b.n_pairs) return false;
        const int lq = a.b.lq[p], lr = a.b.lr[p];
        const int64_t off = a.b.off[p];
        const bool ok = (lr & 3) == 0 && (off & 3) == 0 && lr >= a.b.topk && lr <= a.b.max_lr &&
                        lr <= vsc::kBlockCols * vsc::kMaxRowBlocks && lq <= a.b.max_lq;
        if (!ok) {
            if (lane == 0) {
                a.w.skip[p] = 1;
                a.out.list[atomicAdd(a.out.count, 1)] = p;
            }
            continue;
        }
        if (lq <= 0) continue;  // nothing to read; T2 reports zero boxes
        c.pair = p; c.row0 = 0; c.lq = lq; c.lr = lr; c.base = a.b.sims + off;
        return true;
    }
}

template <int K>
__global__ void __launch_bounds__(kT1Threads, 1) tn_topk_kernel(const T1Args a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pitch = a.pitch;
    unsigned char *mine = smem_raw + (size_t)warp * ((t1_warp_bytes(pitch) + 127) / 128 * 128);
    float *tile = reinterpret_cast<float *>(mine);
    float *bm = tile + (size_t)kStages * kTileRows * pitch;
    float *cand_val = bm + vsc::kMaxRowBlocks * 32;
    int *cand_col = reinterpret_cast<int *>(cand_val + vsc::kMaxCand * 32);
    uint64_t *bar = reinterpret_cast<uint64_t *>(cand_col + vsc::kMaxCand * 32);

    if (lane == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    auto issue = [&](const Cursor &c, int stage) {
        const int rows = min(kTileRows, c.lq - c.row0);
        const uint32_t row_bytes = (uint32_t)c.lr * 4u;
        if (lane == 0) mbar_expect_tx(&bar[stage], row_bytes * rows);
        __syncwarp();
        if (lane < rows)
            bulk_load(tile + ((size_t)stage * kTileRows + lane) * pitch,
                      c.base + (size_t)(c.row0 + lane) * c.lr, row_bytes, &bar[stage]);
    };

    Cursor cur;
    bool have = claim_pair(a, lane, cur);
    if (have) issue(cur, 0);
    int stage = 0;
    uint32_t parity = 0;  // bit s = phase parity of stage s
    bool pair_overflow = false;
    while (have) {
        Cursor nxt = cur;
        bool have_next = true;
        if (cur.row0 + kTileRows < cur.lq) nxt.row0 = cur.row0 + kTileRows;
        else have_next = claim_pair(a, lane, nxt);
        if (have_next) issue(nxt, stage ^ 1);

        mbar_wait(&bar[stage], (parity >> stage) & 1u);
        parity ^= 1u << stage;

        const int row = cur.row0 + lane;
        bool ok = true;
        if (row < cur.lq) {
            float val[K]; int col[K];
            ok = vsc::select_row<K>(tile + ((size_t)stage * kTileRows + lane) * pitch, cur.lr, bm + lane,
                                    cand_val + lane, cand_col + lane, 32, val, col);
            const size_t node = (size_t)cur.pair * a.b.max_nodes + (size_t)row * K;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                a.w.ref_of[node + i] = (uint16_t)col[i];
                a.w.sim_of[node + i] = val[i];
            }
        }
        pair_overflow |= __any_sync(kFullMask, !ok);
        if (!have_next || nxt.pair != cur.pair) {
            if (pair_overflow && lane == 0) {
                a.w.skip[cur.pair] = 1;
                a.out.list[atomicAdd(a.out.count, 1)] = cur.pair;
            }
            pair_overflow = false;
        }
        __syncwarp();  // every lane is done with `stage` before it is refilled
        cur = nxt; have = have_next; stage ^= 1;
    }
}

// ------------------------------------------------------------------ T1e: edges
template <typename MaskT>
__global__ void __launch_bounds__(256) tn_edges_kernel(const Batch b, const Workspace w) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int pair = (int)(idx / b.max_lq);
    if (pair >= b.n_pairs) return;
    const int q_src = (int)(idx - (long long)pair * b.max_lq);
    const int lq = b.lq[pair];
    if (q_src >= lq || w.skip[pair]) return;
    const int K = b.topk, step = b.step;
    const size_t nb = (size_t)pair * b.max_nodes;
    const uint16_t *ref_of = w.ref_of + nb;
    const float *sim_of = w.sim_of + nb;
    MaskT *pred = reinterpret_cast<MaskT *>(w.pred) + nb;

    int r_src[kMaxTop]; uint32_t window[kMaxTop];  // window[a]: refs linked from this row, relative to r_src[a]
#pragma unroll
    for (int x = 0; x < kMaxTop; ++x) {
        r_src[x] = x < K ? (int)ref_of[q_src * K + x] : INT_MIN / 2;
        window[x] = 0;
    }
    const int q_end = min(lq, q_src + step);
    for (int q_dst = q_src + 1; q_dst < q_end; ++q_dst) {
        uint32_t accepted = 0;
        for (int bb = 0; bb < K; ++bb) {
            const int vd = q_dst * K + bb;
            if (!(sim_of[vd] >= b.min_sim)) continue;  // C4
            const int rd = ref_of[vd];
            MaskT bits = 0;
#pragma unroll
            for (int x = 0; x < kMaxTop; ++x) {
                const int d = rd - r_src[x];
                if (d > 0 && d < step && !(window[x] & ((2u << d) - 1u)))  // C2, C3
                    bits |= (MaskT)1 << ((step - 1 - (q_dst - q_src)) * K + x);
            }
            if (bits) {
                accepted |= 1u << bb;
                if (sizeof(MaskT) == 8)
                    atomicOr(reinterpret_cast<unsigned long long *>(&pred[vd]), (unsigned long long)bits);
                else
                    atomicOr(reinterpret_cast<unsigned int *>(&pred[vd]), (unsigned int)bits);
            }
        }
        for (int bb = 0; bb < K; ++bb) {
            if (!((accepted >> bb) & 1)) continue;
            const int rd = ref_of[q_dst * K + bb];
#pragma unroll
            for (int x = 0; x < kMaxTop; ++x) {
                const int d = rd - r_src[x];
                if (d >= 0 && d < step) window[x] |= 1u << d;
            }
        }
    }
}

// ------------------------------------------------------------------ T2: longest-path sweeps
constexpr int kT2Threads = 128;  // 4 warps = 16 pairs per CTA

__device__ __forceinline__ uint32_t oct_max(uint32_t v, unsigned mask) {
    v = max(v, __shfl_xor_sync(mask, v, 1));
    v = max(v, __shfl_xor_sync(mask, v, 2));
    return max(v, __shfl_xor_sync(mask, v, 4));
}
__device__ __forceinline__ int oct_min(int v, unsigned mask) {
    v = min(v, __shfl_xor_sync(mask, v, 1));
    v = min(v, __shfl_xor_sync(mask, v, 2));
    return min(v, __shfl_xor_sync(mask, v, 4));
}
__device__ __forceinline__ int oct_add(int v, unsigned mask) {
    v += __shfl_xor_sync(mask, v, 1);
    v += __shfl_xor_sync(mask, v, 2);
    return v + __shfl_xor_sync(mask, v, 4);
}

template <typename MaskT>
struct PairState {
    const MaskT *pred; MaskT *zero;
    const float *sim_of; const uint16_t *ref_of;
    float *dist; int8_t *slot; uint16_t *gen; uint16_t *chain; uint32_t *lbest;
};

// One lane relaxes its own node: FIRST maximal predecessor in ascending slot order.
template <typename MaskT, bool FIRST>
__device__ __forceinline__ void relax_node(const PairState<MaskT> &s, const int16_t *slot_off, int v,
                                           int layer_base, float &best, int &best_slot, int &gen) {
    MaskT pm = s.pred[v];
    best = 0.0f; best_slot = -1; gen = 0;
    if (!pm) return;
    const MaskT zm = FIRST ? (MaskT)0 : s.zero[v];
    const float w = s.sim_of[v];
    while (pm) {
        const int sl = sizeof(MaskT) == 8 ? __ffsll((long long)pm) - 1 : __ffs((int)pm) - 1;
        pm &= pm - 1;
        const int src = layer_base + slot_off[sl];
        const float cand = s.dist[src] + (((zm >> sl) & 1) ? 0.0f : w);
        if (best_slot < 0 || cand > best) { best = cand; best_slot = sl; }
        if (FIRST) gen = max(gen, (int)s.gen[src] + 1);
    }
    if (!(best >= 0.0f)) { best = 0.0f; best_slot = -1; }  // networkx: negative best -> (0, v)
}

template <typename MaskT>
__global__ void __launch_bounds__(kT2Threads) tn_dp_kernel(const Batch b, const Workspace w, const WorkList out) {
    __shared__ int16_t slot_off[64];
    const int K = b.topk, step = b.step;
    if (threadIdx.x < 64) {
        const int sl = threadIdx.x;
        slot_off[sl] = (int16_t)(sl % K - (step - 1 - sl / K) * K);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, sub = lane & 7, oct = lane >> 3;
    const unsigned om = 0xFFu << (oct * 8);
    const int pair = (blockIdx.x * (kT2Threads / 32) + (threadIdx.x >> 5)) * 4 + oct;
    if (pair >= b.n_pairs || w.skip[pair]) return;  // whole octet leaves together

    const int lq = b.lq[pair];
    const int box_cap = b.max_path + 1;
    int32_t *boxes = b.boxes + (size_t)pair * box_cap * 4;
    const size_t nb = (size_t)pair * b.max_nodes;
    PairState<MaskT> s;
    s.pred = reinterpret_cast<const MaskT *>(w.pred) + nb;
    s.zero = reinterpret_cast<MaskT *>(w.zero) + nb;
    s.sim_of = w.sim_of + nb; s.ref_of = w.ref_of + nb;
    s.dist = w.dist + nb; s.slot = w.slot + nb; s.gen = w.gen + nb;
    s.chain = w.chain + (size_t)pair * b.max_lq;
    s.lbest = w.lbest + (size_t)pair * b.max_lq;
    const bool ranked = sub < K;

    // first sweep: every layer
    for (int q = 0; q < lq; ++q) {
        const int base = q * K, v = base + sub;
        float d = 0.0f;
        if (ranked) {
            int sl, g;
            relax_node<MaskT, true>(s, slot_off, v, base, d, sl, g);
            s.dist[v] = d; s.slot[v] = (int8_t)sl; s.gen[v] = (uint16_t)g; s.zero[v] = 0;
        }
        const uint32_t lm = oct_max(ranked ? __float_as_uint(d) : 0u, om);  // dist >= +0: bits are ordered
        if (sub == 0) s.lbest[q] = lm;
        __syncwarp(om);
    }

    int n_boxes = 0;
    bool ambiguous = false;
    for (int round = 0; round <= b.max_path; ++round) {
        // end node: maximum distance; ties -> smallest Kahn generation
        uint32_t mk = 0;
        for (int q = sub; q < lq; q += 8) mk = max(mk, s.lbest[q]);
        mk = oct_max(mk, om);
        if (mk == 0u) break;  // only zero-length paths left: networkx returns [source]
        int bg = INT_MAX, bv = -1, cnt = 0;
        for (int q = sub; q < lq; q += 8) {
            if (s.lbest[q] != mk) continue;
            for (int r = 0; r < K; ++r) {
                const int v = q * K + r;
                if (__float_as_uint(s.dist[v]) != mk) continue;
                const int g = s.gen[v];
                if (g < bg) { bg = g; bv = v; cnt = 1; }
                else if (g == bg) ++cnt;
            }
        }
        const int g_min = oct_min(bg, om);
        const bool mine = bg == g_min;
        if (oct_add(mine ? cnt : 0, om) > 1) { ambiguous = true; break; }
        const unsigned who = __ballot_sync(om, mine) & om;
        const int end = __shfl_sync(om, bv, __ffs(who) - 1);

        int q_first_dst = 0, q_last = 0;
        if (sub == 0) {
            int len = 0;
            for (int v = end;;) {
                s.chain[len++] = (uint16_t)v;
                const int sl = s.slot[v];
                if (sl < 0) break;
                s.zero[v] |= (MaskT)1 << sl;  // spent edge
                v = (v / K) * K + slot_off[sl];
            }
            float score = 0.0f;
            for (int i = len - 1; i >= 0; --i) score += s.sim_of[s.chain[i]];
            const int first = s.chain[len - 1], last = s.chain[0];
            q_first_dst = (len >= 2 ? (int)s.chain[len - 2] : last) / K;
            q_last = last / K;
            int q_lo = 0, q_hi = 0, r_lo = 0, r_hi = 0;
            if (score > 0.0f) {  // q and (by C2) r increase strictly along a chain
                q_lo = first / K; q_hi = q_last;
                r_lo = s.ref_of[first]; r_hi = s.ref_of[last];
            }
            const double mean_extent = (double)(r_hi - r_lo + q_hi - q_lo) / 2.0;
            double worst = 0.0;
            for (int k = 0; k < n_boxes; ++k) {
                const int32_t *g = boxes + 4 * k;
                long long ww = (long long)min(q_hi, g[2]) - max(q_lo, g[0]) + 1;
                long long hh = (long long)min(r_hi, g[3]) - max(r_lo, g[1]) + 1;
                ww = ww < 0 ? 0 : ww; hh = hh < 0 ? 0 : hh;
                const long long inter = ww * hh;
                const long long a1 = (long long)(q_hi - q_lo + 1) * (r_hi - r_lo + 1);
                const long long a2 = (long long)(g[2] - g[0] + 1) * (g[3] - g[1] + 1);
                const double iou = (double)inter / (double)(a1 + a2 - inter);
                if (k == 0 || iou > worst) worst = iou;
            }
            const int shorter = min(r_hi - r_lo, q_hi - q_lo);
            if (mean_extent != 0.0 && __fdiv_rn(score, (float)mean_extent) > b.min_sim &&
                (double)shorter > b.min_length && worst < b.max_iou) {
                int32_t *o = boxes + 4 * n_boxes;
                o[0] = q_lo; o[1] = r_lo; o[2] = q_hi; o[3] = r_hi;
                ++n_boxes;
            }
        }
        q_first_dst = __shfl_sync(om, q_first_dst, oct * 8);
        q_last = __shfl_sync(om, q_last, oct * 8);
        n_boxes = __shfl_sync(om, n_boxes, oct * 8);
        __syncwarp(om);
        if (round == b.max_path) break;

        // incremental sweep: layers before the first zeroed edge keep their distances, and the
        // wave dies `step-1` layers after the last distance that changed
        int last_changed = INT_MIN / 2;
        for (int q = q_first_dst; q < lq && (q <= q_last || q <= last_changed + step - 1); ++q) {
            const int base = q * K, v = base + sub;
            float d = 0.0f;
            bool changed = false;
            if (ranked) {
                int sl, g;
                const uint32_t before = __float_as_uint(s.dist[v]);
                relax_node<MaskT, false>(s, slot_off, v, base, d, sl, g);
                changed = __float_as_uint(d) != before;
                s.dist[v] = d; s.slot[v] = (int8_t)sl;
            }
            if (__ballot_sync(om, changed) & om) last_changed = q;
            const uint32_t lm = oct_max(ranked ? __float_as_uint(d) : 0u, om);
            if (sub == 0) s.lbest[q] = lm;
            __syncwarp(om);
        }
    }
    if (sub == 0) {
        if (ambiguous) {
            w.skip[pair] = 1;
            out.list[atomicAdd(out.count, 1)] = pair;
        } else {
            b.n_boxes[pair] = n_boxes;
            if (b.status) b.status[pair] = 0;
        }
    }
}

// ------------------------------------------------------------------ T3: MaxSim per box
__global__ void __launch_bounds__(128) tn_maxsim_kernel(const Batch b, const Workspace w) {
    const int pair = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (pair >= b.n_pairs || w.skip[pair]) return;
    const int box_cap = b.max_path + 1, lr = b.lr[pair];
    const float *sims = b.sims + b.off[pair];
    const int nbx = b.n_boxes[pair];
    for (int k = 0; k < nbx; ++k) {
        const int32_t *g = b.boxes + ((size_t)pair * box_cap + k) * 4;
        const int h = g[2] - g[0], wd = g[3] - g[1];  // exclusive upper bounds (localization.py:91)
        float best = -INFINITY;
        for (int e = lane; e < h * wd; e += 32) {
            const int r = e / wd, c = e - r * wd;
            best = fmaxf(best, sims[(size_t)(g[0] + r) * lr + g[1] + c]);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) best = fmaxf(best, __shfl_xor_sync(kFullMask, best, d));
        if (lane == 0) b.box_maxsim[(size_t)pair * box_cap + k] = best;
    }
}

__global__ void iota_kernel(int32_t *count, int32_t *list, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = i;
    if (i == 0) *count = n;
}

template <int K>
int launch_topk(const T1Args &a, int grid, size_t smem, cudaStream_t stream) {
    VSC_CUDA_CHECK(cudaFuncSetAttribute(tn_topk_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VSC_CUDA_CHECK(cudaFuncSetAttribute(tn_topk_kernel<K>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    tn_topk_kernel<K><<<grid, kT1Threads, smem, stream>>>(a);
    VSC_CUDA_CHECK(cudaGetLastError());
    return VSC_OK;
}

size_t t1_smem_bytes(int max_lr) {
    return kT1Warps * ((t1_warp_bytes(tile_pitch(max_lr)) + 127) / 128 * 128);
}

}  // namespace

namespace vsc {
namespace tn {

bool pipeline_supported(const Batch &b) {
    if (b.max_lr > vsc::kBlockCols * vsc::kMaxRowBlocks || b.max_lr < b.topk) return false;
    if (b.topk < 1 || b.topk > kMaxTop) return false;
    if ((reinterpret_cast<uintptr_t>(b.sims) & 15u) != 0) return false;
    if (b.max_nodes > 65535) return false;
    return t1_smem_bytes(b.max_lr) <= 227 * 1024;
}

int launch_pipeline(const Batch &b, const WorkList &out, cudaStream_t stream) {
    const bool wide = (b.step - 1) * b.topk > 32;
    const size_t mask_bytes = wide ? 8 : 4;
    const size_t P = (size_t)b.n_pairs, N = (size_t)b.max_nodes, L = (size_t)b.max_lq;
    // one stream-ordered allocation, carved by alignment
    size_t sz = 0;
    auto take = [&](size_t bytes) { size_t at = sz; sz += (bytes + 255) / 256 * 256; return at; };
    const size_t o_pred = take(P * N * mask_bytes), o_zero = take(P * N * mask_bytes);
    const size_t o_sim = take(P * N * 4), o_dist = take(P * N * 4), o_lbest = take(P * L * 4);
    const size_t o_ref = take(P * N * 2), o_gen = take(P * N * 2), o_chain = take(P * L * 2);
    const size_t o_slot = take(P * N), o_skip = take(P), o_cursor = take(4);
    unsigned char *base = nullptr;
    VSC_CUDA_CHECK(cudaMallocAsync(&base, sz, stream));
    Workspace w;
    w.pred = base + o_pred; w.zero = base + o_zero;
    w.sim_of = reinterpret_cast<float *>(base + o_sim); w.dist = reinterpret_cast<float *>(base + o_dist);
    w.lbest = reinterpret_cast<uint32_t *>(base + o_lbest);
    w.ref_of = reinterpret_cast<uint16_t *>(base + o_ref); w.gen = reinterpret_cast<uint16_t *>(base + o_gen);
    w.chain = reinterpret_cast<uint16_t *>(base + o_chain);
    w.slot = reinterpret_cast<int8_t *>(base + o_slot);
    w.skip = base + o_skip; w.cursor = reinterpret_cast<int32_t *>(base + o_cursor);
    int rc = VSC_OK;
    auto fail = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == VSC_OK) { vsc::set_error("%s: %s", what, cudaGetErrorString(e)); rc = VSC_ERR_CUDA; }
    };
    fail(cudaMemsetAsync(w.pred, 0, P * N * mask_bytes, stream), "memset pred");
    fail(cudaMemsetAsync(w.skip, 0, (o_cursor - o_skip) + 4, stream), "memset flags");

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (rc == VSC_OK) {
        T1Args a; a.b = b; a.w = w; a.out = out; a.pitch = tile_pitch(b.max_lr);
        const size_t smem = t1_smem_bytes(b.max_lr);
        const int per_sm = (int)((227 * 1024) / (smem + 1024)) > 0 ? (int)((227 * 1024) / (smem + 1024)) : 1;
        const int grid = sms * per_sm;
        switch (b.topk) {
            case 1: rc = launch_topk<1>(a, grid, smem, stream); break;
            case 2: rc = launch_topk<2>(a, grid, smem, stream); break;
            case 3: rc = launch_topk<3>(a, grid, smem, stream); break;
            case 4: rc = launch_topk<4>(a, grid, smem, stream); break;
            case 5: rc = launch_topk<5>(a, grid, smem, stream); break;
            case 6: rc = launch_topk<6>(a, grid, smem, stream); break;
            case 7: rc = launch_topk<7>(a, grid, smem, stream); break;
            default: rc = launch_topk<8>(a, grid, smem, stream); break;
        }
        vsc::count_launch();
    }
    if (rc == VSC_OK) {
        const long long threads = (long long)b.n_pairs * b.max_lq;
        const int grid = (int)((threads + 255) / 256);
        if (wide) tn_edges_kernel<uint64_t><<<grid, 256, 0, stream>>>(b, w);
        else tn_edges_kernel<uint32_t><<<grid, 256, 0, stream>>>(b, w);
        fail(cudaGetLastError(), "tn_edges_kernel");
        vsc::count_launch();
    }
    if (rc == VSC_OK) {
        const int pairs_per_cta = (kT2Threads / 32) * 4;
        const int grid = (b.n_pairs + pairs_per_cta - 1) / pairs_per_cta;
        if (wide) tn_dp_kernel<uint64_t><<<grid, kT2Threads, 0, stream>>>(b, w, out);
        else tn_dp_kernel<uint32_t><<<grid, kT2Threads, 0, stream>>>(b, w, out);
        fail(cudaGetLastError(), "tn_dp_kernel");
        vsc::count_launch();
    }
    if (rc == VSC_OK && b.box_maxsim) {
        tn_maxsim_kernel<<<(b.n_pairs + 3) / 4, 128, 0, stream>>>(b, w);
        fail(cudaGetLastError(), "tn_maxsim_kernel");
        vsc::count_launch();
    }
    cudaFreeAsync(base, stream);
    return rc;
}

}  // namespace tn
}  // namespace vsc

extern "C" int vcsl_tn_batch(const float *d_sims, const int64_t *d_off, const int32_t *d_lq,
                             const int32_t *d_lr, int32_t n_pairs, int32_t max_lq, int32_t max_lr,
                             const vsc_tn_params *p, int32_t *d_boxes, int32_t *d_n_boxes,
                             float *d_box_maxsim, int32_t *d_status, int32_t force_exact_order,
                             vsc_stream_t stream_) {
    using namespace vsc::tn;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!p || n_pairs < 0) { vsc::set_error("vcsl_tn_batch: bad arguments"); return VSC_ERR_INVALID; }
    if (n_pairs == 0) return VSC_OK;
    if (!d_sims || !d_off || !d_lq || !d_lr || !d_boxes || !d_n_boxes) {
        vsc::set_error("vcsl_tn_batch: null device pointer"); return VSC_ERR_INVALID;
    }
    if (p->tn_top_k < 1 || p->tn_top_k > kMaxTop || p->tn_max_step < 1 || p->tn_max_step > 31 ||
        (p->tn_max_step - 1) * p->tn_top_k > 64 || p->max_path < 0 || p->max_path + 1 > kMaxBoxes) {
        vsc::set_error("vcsl_tn_batch: unsupported parameters (need tn_top_k<=%d, "
                       "(tn_max_step-1)*tn_top_k<=64, max_path<%d)", kMaxTop, kMaxBoxes);
        return VSC_ERR_INVALID;
    }
    if (max_lr > 65535 || max_lq < 0 || max_lr < 0) {
        vsc::set_error("vcsl_tn_batch: max_lr %d out of range (<= 65535)", max_lr);
        return VSC_ERR_INVALID;
    }
    Batch b;
    b.sims = d_sims; b.off = d_off; b.lq = d_lq; b.lr = d_lr; b.n_pairs = n_pairs;
    b.step = p->tn_max_step; b.topk = p->tn_top_k; b.max_path = p->max_path;
    b.min_sim = p->min_sim; b.min_length = p->min_length; b.max_iou = p->max_iou;
    b.boxes = d_boxes; b.n_boxes = d_n_boxes; b.box_maxsim = d_box_maxsim; b.status = d_status;
    b.max_lq = max_lq > 0 ? max_lq : 1;
    b.max_lr = max_lr;
    b.max_nodes = b.max_lq * (p->tn_top_k < max_lr ? p->tn_top_k : (max_lr > 0 ? max_lr : 1));
    if (b.max_nodes > 65535) {
        vsc::set_error("vcsl_tn_batch: %d graph nodes exceed the 16-bit node index", b.max_nodes);
        return VSC_ERR_CAPACITY;
    }
    // two device-side work lists: [count, ids...]
    int32_t *lists = nullptr;
    const size_t list_len = (size_t)n_pairs + 1;
    VSC_CUDA_CHECK(cudaMallocAsync(&lists, sizeof(int32_t) * 2 * list_len, stream));
    WorkList A{lists, lists + 1}, B{lists + list_len, lists + list_len + 1};
    int rc = VSC_OK;
    cudaError_t e = cudaMemsetAsync(lists, 0, sizeof(int32_t), stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(lists + list_len, 0, sizeof(int32_t), stream);
    if (e != cudaSuccess) { vsc::set_error("memset: %s", cudaGetErrorString(e)); rc = VSC_ERR_CUDA; }
    if (rc == VSC_OK) {
        if (force_exact_order) {
            iota_kernel<<<(n_pairs + 255) / 256, 256, 0, stream>>>(B.count, B.list, n_pairs);
            vsc::count_launch();
        } else if (pipeline_supported(b)) {
            rc = launch_pipeline(b, A, stream);
            if (rc == VSC_OK) rc = launch_fused(b, false, &A, &B, 2, stream);
        } else {
            rc = launch_fused(b, false, nullptr, &B, 2, stream);
        }
    }
    if (rc == VSC_OK) rc = launch_fused(b, true, &B, nullptr, 1, stream);
    cudaFreeAsync(lists, stream);
    return rc;
}
