// Temporal-network (TN) alignment, GRAPH STAGE on a compact graph (B200, sm_100a): edges -> compaction -> longest
// paths by Kahn generation.  Replaces the networkx part of vcsl.vta `tn` (alipay/VCSL @ c39269d5: DiGraph
// construction + up to max_path+1 dag_longest_path sweeps + box filter) behind vsc/baseline/localization.py:44-46,58.
// Contract: oracle/tn_networkx.py; the formulation itself is modelled line by line on the CPU in
// oracle/tn_graph_model.c, which the CPU test suite pins against oracle/tn_fast.c.
//
// Why: a 300x300 pair has 1500 top-5 nodes but only ~350 of them have a predecessor, ~380 edges in total and only
// ~100 edges between two such "active" nodes.  The layer-by-layer kernels (tn_pipeline.cu) walk 300 row layers per
// sweep -- one long dependent chain per pair (0.32 ms for a single pair).  Here:
//   tn_build_kernel  (one CTA per pair)   C2 screening through per-row reference bitmaps in shared memory, C3 / C4 and
//                    the predecessor-slot order exactly as oracle/tn_fast.c; then the graph is compacted to
//                    {active nodes in node order, their incoming edges in slot order, the source nodes that feed them}.
//   tn_paths_kernel  (one warp per pair)  generation 1 (all predecessors are sources) in one parallel step, the inner
//                    nodes generation by generation (G ~ 5 without a copy, ~40 with one); per extracted chain only the
//                    generations downstream of it are relaxed again.  Everything lives in ~15 KB of shared memory.
// Pairs whose graph exceeds the fixed tables, or whose end-node tie generations cannot break, go to the general /
// exact-order kernel (tn_fused.cu) through the work list, like before.
#include <limits.h>

#include "tn_common.cuh"

namespace {

using vsc::kFullMask;
using vsc::tn::Batch;
using vsc::tn::kRefMask;
using vsc::tn::kSimOk;
using vsc::tn::WorkList;
using vsc::tn::Workspace;

constexpr int kMaxActive = 448, kMaxSources = 448, kMaxEdges = 576, kMaxGen = 255, kMaxChain = 256;
constexpr int kEntries = kMaxActive + kMaxSources;
constexpr int kBuildThreads = 128;

// ---- global image of one pair's compact graph (what tn_build_kernel writes and tn_paths_kernel loads)
struct GraphHeader { int32_t A, S, E, ok; };
constexpr size_t kOffNode = sizeof(GraphHeader);                                   // u16[kEntries]
constexpr size_t kOffRef = kOffNode + 2 * kEntries;                                // u16[kEntries]
constexpr size_t kOffSim = kOffRef + 2 * kEntries;                                 // f32[kEntries]
constexpr size_t kOffEoff = kOffSim + 4 * kEntries;                                // u16[kMaxActive + 2]
constexpr size_t kOffEsrc = kOffEoff + 2 * (kMaxActive + 2);                       // u16[kMaxEdges]
constexpr size_t kOffInner = kOffEsrc + 2 * kMaxEdges;                             // u8[kMaxActive]
constexpr size_t kGraphBytes = (kOffInner + kMaxActive + 127) / 128 * 128;
static_assert(kOffSim % 4 == 0 && kOffEoff % 2 == 0, "alignment of the graph image");

__device__ __forceinline__ float node_sim(const Workspace &w, size_t node) {
    return *reinterpret_cast<const float *>(static_cast<const unsigned char *>(w.rec) + node * w.rec_bytes + w.sim_off);
}

// ------------------------------------------------------------------ build
__host__ __device__ inline int bitmap_words(int max_lr, int step) { return (max_lr + step + 31) / 32 + 1; }
__host__ __device__ inline size_t build_smem_bytes(int max_nodes, int max_lq, int max_lr, int step) {
    size_t b = (size_t)max_nodes * (2 + 4 + 2);                            // refs, pred, idx
    b = (b + 3) / 4 * 4;
    b += (size_t)max_lq * bitmap_words(max_lr, step) * 4;                  // row bitmaps
    return b + 64;
}

// block-wide exclusive scan of one int per thread (kBuildThreads threads); returns the exclusive prefix, *total = sum
__device__ __forceinline__ int block_scan(int v, int *warp_sums, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(kFullMask, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int base = 0, sum = 0;
#pragma unroll
    for (int k = 0; k < kBuildThreads / 32; ++k) {
        const int s = warp_sums[k];
        if (k < warp) base += s;
        sum += s;
    }
    __syncthreads();
    *total = sum;
    return base + incl - v;
}

template <int K>
__global__ void __launch_bounds__(kBuildThreads) tn_build_kernel(const Batch b, const Workspace w, const WorkList out,
                                                                 unsigned char *graphs) {
    extern __shared__ __align__(16) unsigned char sm[];
    __shared__ int warp_sums[kBuildThreads / 32];
    const int pair = blockIdx.x, tid = threadIdx.x;
    if (w.skip[pair]) return;
    const int lq = b.lq[pair], step = b.step;
    const int n = lq * K;
    GraphHeader *hdr = reinterpret_cast<GraphHeader *>(graphs + (size_t)pair * kGraphBytes);
    if (lq <= 0) { if (tid == 0) *hdr = GraphHeader{0, 0, 0, 1}; return; }
    const int words = bitmap_words(b.max_lr, step);
    uint16_t *refs = reinterpret_cast<uint16_t *>(sm);                     // [n] reference | kSimOk
    uint16_t *idx = refs + b.max_nodes;                                    // [n] entry index (0xFFFF: not in the graph)
    uint32_t *pred = reinterpret_cast<uint32_t *>(sm + (((size_t)b.max_nodes * 4 + 3) / 4 * 4));   // [n]
    uint32_t *rowbits = pred + b.max_nodes;                                // [lq][words]
    const size_t nb = (size_t)pair * b.max_nodes;

    for (int v = tid; v < n; v += kBuildThreads) { refs[v] = w.ref_of[nb + v]; pred[v] = 0; idx[v] = 0xFFFF; }
    for (int i = tid; i < lq * words; i += kBuildThreads) rowbits[i] = 0;
    __syncthreads();
    for (int v = tid; v < n; v += kBuildThreads) {
        const int r = refs[v] & kRefMask;
        atomicOr(&rowbits[(v / K) * words + (r >> 5)], 1u << (r & 31));
    }
    __syncthreads();

    // ---- edges: one thread per source row (oracle/tn_graph_model.c gm_build)
    for (int q_src = tid; q_src < lq; q_src += kBuildThreads) {
        int r_src[K]; uint32_t window[K];
#pragma unroll
        for (int a = 0; a < K; ++a) { r_src[a] = refs[q_src * K + a] & kRefMask; window[a] = 0; }
        for (int o = 1; o < step && q_src + o < lq; ++o) {
            const int q_dst = q_src + o;
            const uint32_t *bits = rowbits + q_dst * words;
            const uint16_t *rd_row = refs + q_dst * K;
            uint32_t accepted = 0;
#pragma unroll
            for (int a = 0; a < K; ++a) {
                const int lo = r_src[a] + 1;
                const unsigned long long two = (unsigned long long)bits[lo >> 5] | ((unsigned long long)bits[(lo >> 5) + 1] << 32);
                uint32_t hit = (uint32_t)(two >> (lo & 31)) & ((1u << (step - 1)) - 1u);
                while (hit) {                                              // C2 candidates (rare)
                    const int d = __ffs(hit); hit &= hit - 1;
                    const int rd = r_src[a] + d;
                    int bb = 0;
                    while ((rd_row[bb] & kRefMask) != rd) ++bb;
                    if (!(rd_row[bb] & kSimOk)) continue;                  // C4
                    if (window[a] & ((2u << d) - 1u)) continue;            // C3
                    atomicOr(&pred[q_dst * K + bb], 1u << ((step - 1 - o) * K + a));
                    accepted |= 1u << bb;
                }
            }
            while (accepted) {
                const int bb = __ffs(accepted) - 1; accepted &= accepted - 1;
                const int rd = rd_row[bb] & kRefMask;
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    const int d = rd - r_src[a];
                    if ((unsigned)d < (unsigned)step) window[a] |= 1u << d;
                }
            }
        }
    }
    __syncthreads();

    // ---- compaction.  Every thread owns a contiguous chunk of nodes, so entry indices follow node order.
    const int chunk = (n + kBuildThreads - 1) / kBuildThreads;
    const int v0 = min(n, tid * chunk), v1 = min(n, v0 + chunk);
    int n_act = 0, n_edge = 0;
    for (int v = v0; v < v1; ++v) {
        const uint32_t m = pred[v];
        if (!m) continue;
        ++n_act; n_edge += __popc(m);
        uint32_t mm = m;                                                   // mark the sources of these edges
        while (mm) {
            const int slot = __ffs(mm) - 1; mm &= mm - 1;
            const int u = (v / K - (step - 1 - slot / K)) * K + slot % K;
            if (!pred[u]) idx[u] = 0xFFFE;                                 // benign race: everybody writes the same value
        }
    }
    int A, E, S;
    int a_at = block_scan(n_act, warp_sums, &A);
    int e_at = block_scan(n_edge, warp_sums, &E);
    int n_src = 0;
    for (int v = v0; v < v1; ++v) n_src += idx[v] == 0xFFFE;
    int s_at = block_scan(n_src, warp_sums, &S);
    if (A > kMaxActive || E > kMaxEdges || S > kMaxSources) {              // table overflow: general kernel
        if (tid == 0) {
            *hdr = GraphHeader{A, S, E, 0};
            if (atomicExch(&w.skip[pair], 1) == 0) out.list[atomicAdd(out.count, 1)] = pair;
        }
        return;
    }
    for (int v = v0; v < v1; ++v) {
        if (pred[v]) idx[v] = (uint16_t)a_at++;
        else if (idx[v] == 0xFFFE) idx[v] = (uint16_t)(A + s_at++);
    }
    __syncthreads();
    unsigned char *gimg = graphs + (size_t)pair * kGraphBytes;
    uint16_t *g_node = reinterpret_cast<uint16_t *>(gimg + kOffNode), *g_ref = reinterpret_cast<uint16_t *>(gimg + kOffRef);
    float *g_sim = reinterpret_cast<float *>(gimg + kOffSim);
    uint16_t *g_eoff = reinterpret_cast<uint16_t *>(gimg + kOffEoff), *g_esrc = reinterpret_cast<uint16_t *>(gimg + kOffEsrc);
    uint8_t *g_inner = gimg + kOffInner;
    for (int v = v0; v < v1; ++v) {
        const int i = idx[v];
        if (i == 0xFFFF) continue;
        g_node[i] = (uint16_t)v; g_ref[i] = refs[v] & kRefMask; g_sim[i] = node_sim(w, nb + v);
        uint32_t m = pred[v];
        if (!m) continue;
        g_eoff[i] = (uint16_t)e_at;
        bool inner = false;
        while (m) {
            const int slot = __ffs(m) - 1; m &= m - 1;
            const int u = (v / K - (step - 1 - slot / K)) * K + slot % K;
            inner = inner || pred[u] != 0;
            g_esrc[e_at++] = idx[u];
        }
        g_inner[i] = inner ? 1 : 0;
    }
    if (tid == 0) { g_eoff[A] = (uint16_t)E; *hdr = GraphHeader{A, S, E, 1}; }
}

// ------------------------------------------------------------------ paths
struct PathSmem {
    float sim[kEntries];
    float dist[kMaxActive];
    uint16_t node[kEntries], ref[kEntries];
    uint16_t eoff[kMaxActive + 2];
    uint16_t esrc[kMaxEdges];
    uint16_t order[kMaxActive];
    uint16_t gstart[kMaxGen + 3];
    uint16_t chain[kMaxChain];
    uint8_t inner[kMaxActive], gen[kMaxActive], reach[kMaxActive], flag[kMaxActive], ezero[kMaxEdges];
    int8_t best[kMaxActive];
    int hist[kMaxGen + 3];
};

// one node against the current distances: FIRST maximal predecessor in slot order; negative best -> (0, none).
// Returns true when the distance changed.
__device__ __forceinline__ bool relax(PathSmem &s, int A, int i) {
    float best = 0.0f; int bs = -1;
    const int e0 = s.eoff[i], e1 = s.eoff[i + 1];
    const float w = s.sim[i];
    for (int e = e0; e < e1; ++e) {
        const int u = s.esrc[e];
        const float cand = (u < A ? s.dist[u] : 0.0f) + (s.ezero[e] ? 0.0f : w);
        if (bs < 0 || cand > best) { best = cand; bs = e - e0; }
    }
    if (bs >= 0 && !(best >= 0.0f)) { best = 0.0f; bs = -1; }
    best += 0.0f;   // -0 -> +0: distances are compared by their bits
    const bool changed = __float_as_uint(best) != __float_as_uint(s.dist[i]);
    s.dist[i] = best; s.best[i] = (int8_t)bs;
    return changed;
}

__global__ void __launch_bounds__(32) tn_paths_kernel(const Batch b, const Workspace w, const WorkList out,
                                                      const unsigned char *graphs, int K) {
    __shared__ PathSmem s;
    const int pair = blockIdx.x, lane = threadIdx.x;
    if (w.skip[pair]) return;
    const unsigned char *gimg = graphs + (size_t)pair * kGraphBytes;
    const GraphHeader h = *reinterpret_cast<const GraphHeader *>(gimg);
    const int A = h.A, S = h.S, E = h.E;
    const int box_cap = b.max_path + 1;
    int4 *boxes = reinterpret_cast<int4 *>(b.boxes) + (size_t)pair * box_cap;
    if (A == 0) {   // no edge at all: every sweep of the reference finds the empty path
        if (lane == 0) { b.n_boxes[pair] = 0; if (b.status) b.status[pair] = 0; }
        return;
    }
    // ---- load the graph image (4-byte words; every array starts 4-byte aligned inside the image)
    {
        const uint32_t *g32 = reinterpret_cast<const uint32_t *>(gimg);
        auto copy = [&](void *dst, size_t off, int bytes) {
            uint32_t *d = static_cast<uint32_t *>(dst);
            const uint32_t *src = g32 + off / 4;
            for (int i = lane; i < (bytes + 3) / 4; i += 32) d[i] = src[i];
        };
        copy(s.node, kOffNode, 2 * (A + S));
        copy(s.ref, kOffRef, 2 * (A + S));
        copy(s.sim, kOffSim, 4 * (A + S));
        copy(s.eoff, kOffEoff, 2 * (A + 1));
        copy(s.esrc, kOffEsrc, 2 * E);
        copy(s.inner, kOffInner, A);
    }
    for (int e = lane; e < E; e += 32) s.ezero[e] = 0;
    for (int i = lane; i < A; i += 32) { s.dist[i] = 0.0f; s.flag[i] = 0; s.reach[i] = 0; s.gen[i] = 0; }
    for (int k = lane; k < kMaxGen + 3; k += 32) s.hist[k] = 0;
    __syncwarp();

    // ---- first sweep.  Generation 1 (no active predecessor) in one step; the inner nodes go to a work list.
    int n_inner = 0;
    for (int i0 = 0; i0 < A; i0 += 32) {
        const int i = i0 + lane;
        const bool in = i < A && s.inner[i];
        if (i < A && !in) { relax(s, A, i); s.gen[i] = 1; }
        const unsigned m = __ballot_sync(kFullMask, in);
        if (in) s.order[n_inner + __popc(m & ((1u << lane) - 1u))] = (uint16_t)i;
        n_inner += __popc(m);
    }
    __syncwarp();
    int G = 1;
    bool overflow = false;
    for (int done = 0, cur = 2; done < n_inner; ++cur) {
        if (cur > kMaxGen) { overflow = true; break; }
        int mine = 0;
        for (int k = lane; k < n_inner; k += 32) {
            const int i = s.order[k];
            if (s.gen[i]) continue;
            bool ready = true;
            const int e1 = s.eoff[i + 1];
            for (int e = s.eoff[i]; e < e1; ++e) {
                const int u = s.esrc[e];
                if (u < A) { const int gu = s.gen[u]; if (gu == 0 || gu >= cur) { ready = false; break; } }
            }
            if (!ready) continue;
            relax(s, A, i);
            s.gen[i] = (uint8_t)cur;
            for (int e = s.eoff[i]; e < e1; ++e) {
                const int u = s.esrc[e];
                if (u < A && s.reach[u] < cur) s.reach[u] = (uint8_t)cur;   // benign race: same value
            }
            ++mine;
        }
        __syncwarp();
        done += __reduce_add_sync(kFullMask, mine);
        G = cur;
    }
    if (overflow) {
        if (lane == 0 && atomicExch(&w.skip[pair], 1) == 0) out.list[atomicAdd(out.count, 1)] = pair;
        return;
    }
    // ---- inner nodes sorted by generation (counting sort): generation g occupies order[gstart[g] .. gstart[g+1])
    for (int k = lane; k < n_inner; k += 32) atomicAdd(&s.hist[s.gen[s.order[k]]], 1);
    __syncwarp();
    if (lane == 0) {
        int run = 0;
        for (int k = 0; k <= G + 1; ++k) { const int c = s.hist[k]; s.gstart[k] = (uint16_t)run; s.hist[k] = run; run += c; }
    }
    __syncwarp();
    for (int i = lane; i < A; i += 32) {   // re-derived from the `inner` flags, so `order` can be overwritten in place
        if (!s.inner[i]) continue;
        s.order[atomicAdd(&s.hist[s.gen[i]], 1)] = (uint16_t)i;   // the order inside a generation is irrelevant
    }
    __syncwarp();

    int n_boxes = 0;
    bool ambiguous = false;
    for (int round = 0; round <= b.max_path; ++round) {
        // ---- end node: largest distance, then smallest generation; an unresolved tie -> exact-order kernel
        uint32_t bk = 0; int bg = INT_MAX, bi = -1, cnt = 0;
        for (int i = lane; i < A; i += 32) {
            const uint32_t k = __float_as_uint(s.dist[i]);
            if (k == 0u) continue;
            const int g = s.gen[i];
            if (k > bk || (k == bk && g < bg)) { bk = k; bg = g; bi = i; cnt = 1; }
            else if (k == bk && g == bg) ++cnt;
        }
        const uint32_t mk = __reduce_max_sync(kFullMask, bk);
        if (mk == 0u) break;   // only zero-length paths left: networkx returns [source]
        const int g_min = __reduce_min_sync(kFullMask, bk == mk ? bg : INT_MAX);
        const bool mine_best = bk == mk && bg == g_min;
        if (__reduce_add_sync(kFullMask, mine_best ? cnt : 0) > 1) { ambiguous = true; break; }
        const int end = __shfl_sync(kFullMask, bi, __ffs(__ballot_sync(kFullMask, mine_best)) - 1);

        // ---- lane 0: walk the chain back (marks the spent edges), score, box, filter
        int len = 0, horizon = 0;
        bool too_long = false;
        if (lane == 0) {
            int first_entry = end;
            for (int i = end;;) {
                if (len >= kMaxChain - 1) { too_long = true; break; }
                s.chain[len++] = (uint16_t)i;
                const int bs = s.best[i];
                if (bs < 0) { first_entry = i; break; }
                const int e = s.eoff[i] + bs;
                s.ezero[e] = 1; s.flag[i] |= 1;
                const int u = s.esrc[e];
                if (u >= A) { first_entry = u; s.chain[len++] = (uint16_t)u; break; }
                i = u;
            }
            if (!too_long) {
                float score = 0.0f;
                for (int k = len - 1; k >= 0; --k) score += s.sim[s.chain[k]];   // float32 sum in path order
                int q_lo = 0, q_hi = 0, r_lo = 0, r_hi = 0;
                if (score > 0.0f) {   // q and (by C2) r increase strictly along a chain
                    q_lo = s.node[first_entry] / K; q_hi = s.node[end] / K;
                    r_lo = s.ref[first_entry]; r_hi = s.ref[end];
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
                // spent edges into generation-1 nodes (at most the chain's first destination) are relaxed right here;
                // inner nodes wait for their generation
                if (round != b.max_path) {
                    for (int k = len - 1; k >= 0; --k) {
                        const int i = s.chain[k];
                        if (i >= A || !(s.flag[i] & 1)) continue;
                        if (!s.inner[i]) {
                            s.flag[i] = 0;
                            if (relax(s, A, i)) { s.flag[i] = 2; horizon = max(horizon, (int)s.reach[i]); }
                        } else {
                            horizon = max(horizon, (int)s.gen[i]);
                        }
                    }
                }
            }
        }
        too_long = __shfl_sync(kFullMask, too_long, 0);
        if (too_long) { ambiguous = true; break; }
        n_boxes = __shfl_sync(kFullMask, n_boxes, 0);
        horizon = __shfl_sync(kFullMask, horizon, 0);
        len = __shfl_sync(kFullMask, len, 0);
        __syncwarp();
        if (round == b.max_path) break;

        // ---- relax again what the spent edges can change, generation by generation, as far as a changed node reaches
        for (int cur = 2; cur <= horizon; ++cur) {
            const int k1 = s.gstart[cur + 1];
            int reach_now = 0;
            for (int k = s.gstart[cur] + lane; k < k1; k += 32) {
                const int i = s.order[k];
                bool affected = s.flag[i] & 1;
                const int e1 = s.eoff[i + 1];
                for (int e = s.eoff[i]; e < e1 && !affected; ++e) {
                    const int u = s.esrc[e];
                    if (u < A && (s.flag[u] & 2)) affected = true;
                }
                uint8_t f = s.flag[i] & ~1;
                if (affected && relax(s, A, i)) { f |= 2; reach_now = max(reach_now, (int)s.reach[i]); }
                s.flag[i] = f;
            }
            __syncwarp();
            horizon = max(horizon, (int)__reduce_max_sync(kFullMask, (unsigned)reach_now));
        }
        // clear the marks of this round: the chain's nodes and whatever changed (scan of the active nodes)
        for (int i = lane; i < A; i += 32) s.flag[i] = 0;
        __syncwarp();
    }
    if (lane == 0) {
        if (ambiguous) {
            if (atomicExch(&w.skip[pair], 1) == 0) {
                if (w.exact_list) w.exact_list[atomicAdd(w.exact_count, 1)] = pair;
                else out.list[atomicAdd(out.count, 1)] = pair;
            }
        } else {
            b.n_boxes[pair] = n_boxes;
            if (b.status) b.status[pair] = 0;
        }
    }
}

}  // namespace

namespace vsc {
namespace tn {

bool graph_v2_supported(const Batch &b) {
    if (b.topk < 1 || b.topk > kMaxTop || (b.step - 1) * b.topk > 32) return false;
    if (b.max_nodes > 65000 || b.max_lr > kRefMask) return false;
    if ((reinterpret_cast<uintptr_t>(b.boxes) & 15u) != 0) return false;
    return build_smem_bytes(b.max_nodes, b.max_lq, b.max_lr, b.step) <= 160 * 1024;
}

size_t graph_v2_scratch_bytes(const Batch &b) { return (size_t)b.n_pairs * kGraphBytes; }

template <int K>
static int launch_build(const Batch &b, const Workspace &w, const WorkList &out, unsigned char *graphs, cudaStream_t stream) {
    const size_t smem = build_smem_bytes(b.max_nodes, b.max_lq, b.max_lr, b.step);
    VSC_CUDA_CHECK(cudaFuncSetAttribute(tn_build_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tn_build_kernel<K><<<b.n_pairs, kBuildThreads, smem, stream>>>(b, w, out, graphs);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

// edges + compaction, then the longest-path rounds.  `graphs`: graph_v2_scratch_bytes(b) bytes of device scratch.
// `mark_between` (may be null) is called between the two launches (stage timing).
int launch_graph_v2(const Batch &b, const Workspace &w, const WorkList &out, unsigned char *graphs, cudaStream_t stream,
                    void (*mark_between)(cudaStream_t)) {
    int rc;
    switch (b.topk) {
        case 1: rc = launch_build<1>(b, w, out, graphs, stream); break;
        case 2: rc = launch_build<2>(b, w, out, graphs, stream); break;
        case 3: rc = launch_build<3>(b, w, out, graphs, stream); break;
        case 4: rc = launch_build<4>(b, w, out, graphs, stream); break;
        case 5: rc = launch_build<5>(b, w, out, graphs, stream); break;
        case 6: rc = launch_build<6>(b, w, out, graphs, stream); break;
        case 7: rc = launch_build<7>(b, w, out, graphs, stream); break;
        default: rc = launch_build<8>(b, w, out, graphs, stream); break;
    }
    if (rc != VSC_OK) return rc;
    if (mark_between) mark_between(stream);
    VSC_CUDA_CHECK(cudaFuncSetAttribute(tn_paths_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    tn_paths_kernel<<<b.n_pairs, 32, 0, stream>>>(b, w, out, graphs, b.topk);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

}  // namespace tn
}  // namespace vsc
