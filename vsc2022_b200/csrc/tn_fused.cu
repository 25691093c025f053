// Temporal-network (TN) alignment, GENERAL kernel: one CTA per candidate pair, any shape or
// alignment, optional exact Kahn order.  The fast path for aligned rows is tn_pipeline.cu; this
// kernel finishes whatever the pipeline hands back (ambiguous ties, odd shapes, tie-heavy rows).
//
// Replaces vcsl.vta `tn` (alipay/VCSL @ c39269d5, vcsl/vta.py) as called from
// vsc/baseline/localization.py:44-46,58.  The algorithm contract is
// oracle/tn_networkx.py; the array formulation mirrored here is validated on the
// CPU in oracle/tn_fast.c.
//
// Per pair (Lq x Lr float32 similarities, read from HBM exactly once):
//   phase 1  all warps: stream rows, exact per-row top-k (value desc, index asc)
//            -> node table (ref_of, sim_of) in shared memory        [HBM-bound]
//   phase 2  all threads: edge bit-masks (constraints C1-C4)         [on-chip]
//   phase 3  warp 0: longest-path sweeps.  The first sweep visits every row
//            layer; each later sweep only re-relaxes the layers downstream of
//            the chain whose edges were just zeroed.  End-node ties follow
//            networkx: first maximum in Kahn order.  The fast kernel breaks ties
//            by Kahn generation and hands a pair to the exact-order kernel
//            (same code, full Kahn positions) if a tie is still ambiguous.
//   phase 4  all threads: max similarity inside every kept box (MaxSim score).
//
// Node v = q*top + rank.  Edge (q_src,a) -> (q_dst,b) is bit
// slot = (step-1-(q_dst-q_src))*top + a of pred_mask[v_dst]; ascending slot ==
// networkx predecessor insertion order.
#include <limits.h>

#include <vector>

#include "tn_common.cuh"

namespace {

using vsc::tn::Batch;
using vsc::tn::WorkList;
using vsc::tn::kMaxBoxes;
using vsc::tn::kMaxTop;

using vsc::float_to_key;
using vsc::kFullMask;
using vsc::key_to_float;

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kColsPerLane = 10; // phase-1 register tile: 320 columns per pass

struct TnArgs {
    Batch b;
    const int32_t *in_count;   // null: every pair of the batch
    const int32_t *in_list;
    int32_t *out_count;        // unresolved pairs (exact_order=false only)
    int32_t *out_list;
    int status_code;
};

struct Scalars {
    int32_t boxes[kMaxBoxes * 4];
    int32_t n_boxes;
    int32_t ambiguous;
    float red[kWarps];
    int16_t slot_off[64];  // source node = layer_base + slot_off[slot]
};

template <typename MaskT>
struct Smem {
    MaskT *pred, *zero;
    float *sim_of, *dist;
    uint16_t *ref_of, *order, *act, *chain, *queue;
    int8_t *slot;
    uint8_t *indeg;
    Scalars *sc;
};

template <typename MaskT, bool EXACT>
__host__ __device__ inline size_t smem_bytes(int max_nodes, int max_lq) {
    size_t n = (size_t)max_nodes;
    size_t b = n * (2 * sizeof(MaskT) + 2 * sizeof(float) + 3 * sizeof(uint16_t) + 1);
    b += (size_t)max_lq * sizeof(uint16_t);
    if (EXACT) b += n * (sizeof(uint16_t) + 1);
    return b + sizeof(Scalars) + 64;  // + alignment slack
}

template <typename MaskT, bool EXACT>
__device__ inline Smem<MaskT> carve(unsigned char *base, int max_nodes, int max_lq) {
    Smem<MaskT> s;
    size_t n = (size_t)max_nodes;
    unsigned char *p = base;
    s.sc = reinterpret_cast<Scalars *>(p); p += (sizeof(Scalars) + 15) / 16 * 16;
    s.pred = reinterpret_cast<MaskT *>(p); p += n * sizeof(MaskT);
    s.zero = reinterpret_cast<MaskT *>(p); p += n * sizeof(MaskT);
    s.sim_of = reinterpret_cast<float *>(p); p += n * sizeof(float);
    s.dist = reinterpret_cast<float *>(p); p += n * sizeof(float);
    s.ref_of = reinterpret_cast<uint16_t *>(p); p += n * sizeof(uint16_t);
    s.order = reinterpret_cast<uint16_t *>(p); p += n * sizeof(uint16_t);
    s.act = reinterpret_cast<uint16_t *>(p); p += n * sizeof(uint16_t);
    s.chain = reinterpret_cast<uint16_t *>(p); p += (size_t)max_lq * sizeof(uint16_t);
    s.queue = nullptr; s.indeg = nullptr;
    if (EXACT) { s.queue = reinterpret_cast<uint16_t *>(p); p += n * sizeof(uint16_t); }
    s.slot = reinterpret_cast<int8_t *>(p); p += n;
    if (EXACT) { s.indeg = reinterpret_cast<uint8_t *>(p); p += n; }
    return s;
}

// ---------------------------------------------------------------- phase 1
// Exact top-k of one row by a full warp.  On return lane r (< top) holds the
// r-th best (key, column).  Equal values: lower column first.
__device__ inline void warp_row_topk(const float *__restrict__ row, int lr, int top, int lane,
                                     uint32_t &best_key, int &best_col) {
    best_key = 0; best_col = INT_MAX;
    for (int c0 = 0; c0 < lr; c0 += 32 * kColsPerLane) {
        uint32_t key[kColsPerLane];
#pragma unroll
        for (int j = 0; j < kColsPerLane; ++j) {
            int col = c0 + j * 32 + lane;
            key[j] = col < lr ? float_to_key(__ldcs(row + col)) : 0u;
        }
        uint32_t carry_key = best_key; int carry_col = best_col;  // winners of earlier passes
        uint32_t new_key = 0; int new_col = INT_MAX;
        for (int r = 0; r < top; ++r) {
            uint32_t local = carry_key;
#pragma unroll
            for (int j = 0; j < kColsPerLane; ++j) local = max(local, key[j]);
            uint32_t m = __reduce_max_sync(kFullMask, local);
            if (m == 0u) break;
            int mine = (carry_key == m) ? carry_col : INT_MAX;
#pragma unroll
            for (int j = kColsPerLane - 1; j >= 0; --j)
                if (key[j] == m) mine = min(mine, c0 + j * 32 + lane);
            int col = __reduce_min_sync(kFullMask, mine);
            if (carry_key == m && carry_col == col) carry_key = 0u;
#pragma unroll
            for (int j = 0; j < kColsPerLane; ++j)
                if (c0 + j * 32 + lane == col) key[j] = 0u;
            if (lane == r) { new_key = m; new_col = col; }
        }
        best_key = new_key; best_col = new_col;
    }
}

// ---------------------------------------------------------------- phase 3
template <typename MaskT>
struct Relaxed { float dist; int slot; int gen; };

// Full-warp relaxation of node v (pred mask pm != 0).  Lanes enumerate slots.
template <typename MaskT, bool WANT_GEN>
__device__ inline Relaxed<MaskT> relax(const Smem<MaskT> &s, int v, int layer_base, MaskT pm,
                                       int n_slots, int lane) {
    const MaskT zm = s.zero[v];
    const float w = s.sim_of[v];
    uint32_t best_key = 0; int best_slot = -1; int gen = 0;
#pragma unroll
    for (int base = 0; base < (int)sizeof(MaskT) * 8; base += 32) {
        if (base >= n_slots) break;
        const int sl = base + lane;
        const bool has = sl < n_slots && ((pm >> sl) & 1);
        const int src = layer_base + s.sc->slot_off[sl & 63];
        float cand = 0.0f; int g = 0;
        if (has) {
            cand = s.dist[src] + (((zm >> sl) & 1) ? 0.0f : w);
            if (WANT_GEN) g = s.order[src] + 1;
        }
        const uint32_t key = has ? float_to_key(cand) : 0u;
        const uint32_t m = __reduce_max_sync(kFullMask, key);
        if (m > best_key) {  // strict: lower slots win ties
            best_key = m;
            best_slot = base + __ffs(__ballot_sync(kFullMask, key == m)) - 1;
        }
        if (WANT_GEN) gen = max(gen, (int)__reduce_max_sync(kFullMask, (unsigned)g));
    }
    Relaxed<MaskT> r;
    r.dist = key_to_float(best_key);
    r.slot = best_slot;
    r.gen = gen;
    if (!(r.dist >= 0.0f)) { r.dist = 0.0f; r.slot = -1; }  // networkx: negative best -> (0, v)
    return r;
}

// Literal Kahn order (one FIFO queue == networkx generations concatenated); warp 0.
template <typename MaskT>
__device__ inline void kahn_order(const Smem<MaskT> &s, int n, int lq, int top, int step,
                                  int n_slots, int lane) {
    int tail = 0;
    for (int v0 = 0; v0 < n; v0 += 32) {
        int v = v0 + lane;
        int deg = v < n ? __popcll((unsigned long long)s.pred[v]) : -1;
        if (v < n) s.indeg[v] = (uint8_t)deg;
        unsigned z = __ballot_sync(kFullMask, deg == 0);
        if (deg == 0) s.queue[tail + __popc(z & ((1u << lane) - 1u))] = (uint16_t)v;
        tail += __popc(z);
    }
    __syncwarp();
    for (int head = 0; head < tail; ++head) {
        const int u = s.queue[head];
        if (lane == 0) s.order[u] = (uint16_t)head;
        const int q = u / top, a = u - q * top;
        for (int base = 0; base < n_slots; base += 32) {
            // child enumeration order: q_dst ascending, then dst rank ascending
            const int ci = base + lane;
            const int o = ci / top + 1, b = ci - (o - 1) * top;
            const int c = (q + o) * top + b;
            bool hit = false;
            if (ci < n_slots && q + o < lq) {
                const int bit = (step - 1 - o) * top + a;
                hit = (s.pred[c] >> bit) & 1;
            }
            bool ready = false;
            if (hit) { uint8_t d = s.indeg[c] - 1; s.indeg[c] = d; ready = d == 0; }
            unsigned z = __ballot_sync(kFullMask, ready);
            if (ready) s.queue[tail + __popc(z & ((1u << lane) - 1u))] = (uint16_t)c;
            tail += __popc(z);
        }
        __syncwarp();
    }
}

template <typename MaskT, bool EXACT>
__global__ void __launch_bounds__(kThreads) tn_kernel(const TnArgs args) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const Batch &a = args.b;
    const Smem<MaskT> s = carve<MaskT, EXACT>(smem_raw, a.max_nodes, a.max_lq);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n_work = args.in_count ? *args.in_count : a.n_pairs;

    for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
        const int pair = args.in_count ? args.in_list[work] : work;
        const int lq = a.lq[pair], lr = a.lr[pair];
        const int top = min(a.topk, lr);
        const int n = lq * top;
        const int step = a.step;
        const int n_slots = (step - 1) * top;
        const float *__restrict__ sims = a.sims ? a.sims + a.off[pair] : nullptr;
        const int box_cap = a.max_path + 1;

        __syncthreads();  // previous pair fully retired before smem is reused
        for (int i = tid; i < n; i += kThreads) {
            s.pred[i] = 0; s.zero[i] = 0; s.dist[i] = 0.0f; s.slot[i] = -1; s.order[i] = 0;
        }
        if (tid < 64) {
            int sl = tid, o = 0, r = 0;
            if (top > 0) { o = step - 1 - sl / top; r = sl % top; }
            s.sc->slot_off[sl] = (int16_t)(r - o * top);
        }
        if (tid == 0) { s.sc->n_boxes = 0; s.sc->ambiguous = 0; }
        __syncthreads();

        // ---- phase 1: row top-k (each warp streams whole rows, coalesced); or the node records the fast pipeline
        // already holds for this pair (from-features path: there may be no similarity matrix in memory at all)
        if (a.node_ref) {
            for (int i = tid; i < n; i += kThreads) {
                const int q = i / top, r = i - q * top;
                const size_t gi = (size_t)pair * a.max_nodes + (size_t)q * a.topk + r;
                s.ref_of[i] = a.node_ref[gi] & vsc::tn::kRefMask;
                s.sim_of[i] = *reinterpret_cast<const float *>(static_cast<const unsigned char *>(a.node_rec) +
                                                              gi * a.node_rec_bytes + a.node_sim_off);
            }
        } else
        for (int q = warp; q < lq; q += kWarps) {
            uint32_t key; int col;
            warp_row_topk(sims + (size_t)q * lr, lr, top, lane, key, col);
            if (lane < top) {
                s.ref_of[q * top + lane] = (uint16_t)col;
                s.sim_of[q * top + lane] = key_to_float(key);
            }
        }
        __syncthreads();

        // ---- phase 2: edges.  One thread per source row; `window[a]` holds the refs
        // already linked from this row, relative to r_src[a] (constraint C3).
        for (int q_src = tid; q_src < lq; q_src += kThreads) {
            int r_src[kMaxTop]; uint32_t window[kMaxTop];
#pragma unroll
            for (int x = 0; x < kMaxTop; ++x) {
                r_src[x] = x < top ? (int)s.ref_of[q_src * top + x] : INT_MIN / 2;
                window[x] = 0;
            }
            const int q_end = min(lq, q_src + step);
            for (int q_dst = q_src + 1; q_dst < q_end; ++q_dst) {
                uint32_t accepted = 0;
                for (int b = 0; b < top; ++b) {
                    const int vd = q_dst * top + b;
                    if (!(s.sim_of[vd] >= a.min_sim)) continue;  // C4
                    const int rd = s.ref_of[vd];
                    MaskT bits = 0;
#pragma unroll
                    for (int x = 0; x < kMaxTop; ++x) {
                        const int d = rd - r_src[x];
                        if (d > 0 && d < step && !(window[x] & ((2u << d) - 1u)))  // C2, C3
                            bits |= (MaskT)1 << ((step - 1 - (q_dst - q_src)) * top + x);
                    }
                    if (bits) {
                        accepted |= 1u << b;
                        if (sizeof(MaskT) == 8)
                            atomicOr(reinterpret_cast<unsigned long long *>(&s.pred[vd]),
                                     (unsigned long long)bits);
                        else
                            atomicOr(reinterpret_cast<unsigned int *>(&s.pred[vd]), (unsigned int)bits);
                    }
                }
                for (int b = 0; b < top; ++b) {
                    if (!((accepted >> b) & 1)) continue;
                    const int rd = s.ref_of[q_dst * top + b];
#pragma unroll
                    for (int x = 0; x < kMaxTop; ++x) {
                        const int d = rd - r_src[x];
                        if (d >= 0 && d < step) window[x] |= 1u << d;
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 3: longest-path sweeps (warp 0)
        if (warp == 0 && n > 0) {
            if (EXACT) kahn_order<MaskT>(s, n, lq, top, step, n_slots, lane);
            int n_act = 0;
            // first sweep: every layer
            for (int q = 0; q < lq; ++q) {
                const int base = q * top;
                const bool live = lane < top && s.pred[base + lane] != 0;
                unsigned active = __ballot_sync(kFullMask, live);
                while (active) {
                    const int k = __ffs(active) - 1; active &= active - 1;
                    const int v = base + k;
                    Relaxed<MaskT> r = relax<MaskT, !EXACT>(s, v, base, s.pred[v], n_slots, lane);
                    if (lane == 0) {
                        s.dist[v] = r.dist; s.slot[v] = (int8_t)r.slot;
                        if (!EXACT) s.order[v] = (uint16_t)r.gen;
                        s.act[n_act] = (uint16_t)v;
                    }
                    ++n_act;
                }
                __syncwarp();
            }

            for (int round = 0; round <= a.max_path; ++round) {
                // end node: max dist; ties -> smallest order (generation or Kahn position)
                uint32_t bk = 0; int bo = INT_MAX, bv = -1, cnt = 0;
                for (int i = lane; i < n_act; i += 32) {
                    const int v = s.act[i];
                    const uint32_t k = __float_as_uint(s.dist[v]);  // dist >= +0
                    const int o = s.order[v];
                    if (k > bk || (k == bk && o < bo)) { bk = k; bo = o; bv = v; cnt = 1; }
                    else if (k == bk && o == bo) ++cnt;
                }
                const uint32_t m = __reduce_max_sync(kFullMask, bk);
                if (m == 0u) break;  // only zero-length paths left: networkx returns [source]
                const int o_min = __reduce_min_sync(kFullMask, bk == m ? bo : INT_MAX);
                const bool mine = bk == m && bo == o_min;
                const int total = __reduce_add_sync(kFullMask, mine ? cnt : 0);
                if (total > 1) {  // cannot happen with exact Kahn positions
                    if (lane == 0) s.sc->ambiguous = 1;
                    break;
                }
                const int end = __shfl_sync(kFullMask, bv, __ffs(__ballot_sync(kFullMask, mine)) - 1);

                int q_first_dst = 0, q_last = 0;
                if (lane == 0) {
                    int len = 0;
                    for (int v = end;;) {
                        s.chain[len++] = (uint16_t)v;
                        const int sl = s.slot[v];
                        if (sl < 0) break;
                        s.zero[v] |= (MaskT)1 << sl;  // spent edge
                        v = (v / top) * top + s.sc->slot_off[sl];
                    }
                    float score = 0.0f;
                    for (int i = len - 1; i >= 0; --i) score += s.sim_of[s.chain[i]];
                    const int first = s.chain[len - 1], last = s.chain[0];
                    q_first_dst = (len >= 2 ? (int)s.chain[len - 2] : last) / top;
                    q_last = last / top;
                    int q_lo = 0, q_hi = 0, r_lo = 0, r_hi = 0;
                    if (score > 0.0f) {  // q and (by C2) r increase strictly along a chain
                        q_lo = first / top; q_hi = q_last;
                        r_lo = s.ref_of[first]; r_hi = s.ref_of[last];
                    }
                    const double mean_extent = (double)(r_hi - r_lo + q_hi - q_lo) / 2.0;
                    const int nb = s.sc->n_boxes;
                    double worst = 0.0;
                    for (int k = 0; k < nb; ++k) {
                        const int32_t *g = s.sc->boxes + 4 * k;
                        long long w = (long long)min(q_hi, g[2]) - max(q_lo, g[0]) + 1;
                        long long h = (long long)min(r_hi, g[3]) - max(r_lo, g[1]) + 1;
                        w = w < 0 ? 0 : w; h = h < 0 ? 0 : h;
                        const long long inter = w * h;
                        const long long a1 = (long long)(q_hi - q_lo + 1) * (r_hi - r_lo + 1);
                        const long long a2 = (long long)(g[2] - g[0] + 1) * (g[3] - g[1] + 1);
                        const double iou = (double)inter / (double)(a1 + a2 - inter);
                        if (k == 0 || iou > worst) worst = iou;
                    }
                    const int shorter = min(r_hi - r_lo, q_hi - q_lo);
                    if (mean_extent != 0.0 && __fdiv_rn(score, (float)mean_extent) > a.min_sim &&
                        (double)shorter > a.min_length && worst < a.max_iou) {
                        int32_t *o = s.sc->boxes + 4 * nb;
                        o[0] = q_lo; o[1] = r_lo; o[2] = q_hi; o[3] = r_hi;
                        s.sc->n_boxes = nb + 1;
                    }
                }
                q_first_dst = __shfl_sync(kFullMask, q_first_dst, 0);
                q_last = __shfl_sync(kFullMask, q_last, 0);
                __syncwarp();
                if (round == a.max_path) break;

                // incremental sweep: only layers at/after the first zeroed edge can change,
                // and the wave dies `step-1` layers after the last changed distance.
                int last_changed = INT_MIN / 2;
                for (int q = q_first_dst; q < lq && (q <= q_last || q <= last_changed + step - 1); ++q) {
                    const int base = q * top;
                    const bool live = lane < top && s.pred[base + lane] != 0;
                    unsigned active = __ballot_sync(kFullMask, live);
                    while (active) {
                        const int k = __ffs(active) - 1; active &= active - 1;
                        const int v = base + k;
                        const uint32_t before = __float_as_uint(s.dist[v]);
                        Relaxed<MaskT> r = relax<MaskT, false>(s, v, base, s.pred[v], n_slots, lane);
                        if (__float_as_uint(r.dist) != before) last_changed = q;
                        __syncwarp();  // every lane has read the old value before lane 0 overwrites it
                        if (lane == 0) { s.dist[v] = r.dist; s.slot[v] = (int8_t)r.slot; }
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();

        // ---- phase 4: outputs (+ MaxSim score: max over sims[q_lo:q_hi, r_lo:r_hi], exclusive)
        const bool redo = !EXACT && s.sc->ambiguous;
        if (redo) {
            if (tid == 0) args.out_list[atomicAdd(args.out_count, 1)] = pair;
            continue;
        }
        const int nb = s.sc->n_boxes;
        if (tid == 0) {
            a.n_boxes[pair] = nb;
            if (a.status) a.status[pair] = args.status_code;
        }
        for (int i = tid; i < nb * 4; i += kThreads)
            a.boxes[(size_t)pair * box_cap * 4 + i] = s.sc->boxes[i];
        if (a.box_maxsim && sims) {
            for (int k = 0; k < nb; ++k) {
                const int32_t *g = s.sc->boxes + 4 * k;
                const int h = g[2] - g[0], w = g[3] - g[1];
                float best = -INFINITY;
                for (int e = tid; e < h * w; e += kThreads) {
                    const int r = e / w, c = e - r * w;
                    best = fmaxf(best, sims[(size_t)(g[0] + r) * lr + g[1] + c]);
                }
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) best = fmaxf(best, __shfl_xor_sync(kFullMask, best, d));
                if (lane == 0) s.sc->red[warp] = best;
                __syncthreads();
                if (tid == 0) {
                    for (int x = 1; x < kWarps; ++x) best = fmaxf(best, s.sc->red[x]);
                    a.box_maxsim[(size_t)pair * box_cap + k] = best;
                }
                __syncthreads();
            }
        }
    }
}

template <typename MaskT, bool EXACT>
int launch(const TnArgs &a, int grid, cudaStream_t stream) {
    const size_t bytes = smem_bytes<MaskT, EXACT>(a.b.max_nodes, a.b.max_lq);
    int dev = 0, max_optin = 0;
    VSC_CUDA_CHECK(cudaGetDevice(&dev));
    VSC_CUDA_CHECK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    if (bytes > (size_t)max_optin) {
        vsc::set_error("vcsl_tn_batch: pair with %d graph nodes needs %zu B of shared memory (> %d)",
                       a.b.max_nodes, bytes, max_optin);
        return VSC_ERR_CAPACITY;
    }
    VSC_CUDA_CHECK(cudaFuncSetAttribute(tn_kernel<MaskT, EXACT>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    VSC_CUDA_CHECK(cudaFuncSetAttribute(tn_kernel<MaskT, EXACT>,
                                        cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    tn_kernel<MaskT, EXACT><<<grid, kThreads, bytes, stream>>>(a);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

}  // namespace

namespace vsc {
namespace tn {

int launch_fused(const Batch &b, bool exact_order, const WorkList *in, const WorkList *out,
                 int status_code, cudaStream_t stream) {
    TnArgs a;
    a.b = b;
    a.in_count = in ? in->count : nullptr;
    a.in_list = in ? in->list : nullptr;
    a.out_count = out ? out->count : nullptr;
    a.out_list = out ? out->list : nullptr;
    a.status_code = status_code;
    // a work list is consumed by a persistent grid (its length is only known on the device)
    const int grid = in ? (b.n_pairs < 592 ? b.n_pairs : 592) : b.n_pairs;
    const bool wide = (b.step - 1) * b.topk > 32;
    if (exact_order)
        return wide ? launch<uint64_t, true>(a, grid, stream) : launch<uint32_t, true>(a, grid, stream);
    return wide ? launch<uint64_t, false>(a, grid, stream) : launch<uint32_t, false>(a, grid, stream);
}

}  // namespace tn
}  // namespace vsc
