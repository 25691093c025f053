/*
 * oracle/tn_fast.c -- plain-C array formulation of VCSL's temporal-network (TN)
 * alignment.  TEST INFRASTRUCTURE (oracle), never linked into the product.
 *
 * Authority: oracle/tn_networkx.py (literal restatement of alipay/VCSL
 * vcsl/vta.py `tn` @ c39269d5, the engine behind
 * /root/reference/vsc/baseline/localization.py:44-46,58).  This file restates
 * the same algorithm without a graph library so that full-size workloads
 * (thousands of 300x300 matrices) can be checked in seconds; tests/ validates it
 * against tn_networkx on thousands of seeded matrices, ties included.
 *
 * Node n = q*top + rank  (networkx id n+1; the isolated source node 0 is
 * implicit).  An edge into node (q_dst, b) from (q_src, a) is bit
 *     slot = (step-1 - (q_dst-q_src))*top + a
 * of pred_mask[n_dst]; ascending slot order == networkx predecessor insertion
 * order (q_src ascending, then src rank ascending), because VCSL adds edges for
 * q_src ascending, q_dst ascending, and np.where() row-major over
 * [dst rank, src rank].
 *
 * networkx.dag_longest_path semantics reproduced (networkx 3.6.1, dag.py):
 *   - topological order = Kahn generations == one FIFO queue seeded with the
 *     zero-in-degree nodes in id order, children visited in successor insertion
 *     order (q_dst ascending, then dst rank ascending);
 *   - dist[v] = FIRST maximum over predecessors of dist[u] + w(u,v), or (0, v)
 *     when there is no predecessor or the best is negative;
 *   - end node = FIRST maximum of dist in topological order (the source node 0
 *     is first, so an all-zero graph yields the empty path and stops the loop).
 * float32 throughout (numpy-2 weak scalars), sums taken in path order.
 *
 * Build: see oracle/Makefile  (gcc -O2 -shared -fPIC, no -ffast-math).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TN_MAX_SLOTS 64

typedef struct {
    int lq, lr, step, top, n_nodes;
    int32_t *ref_of;      /* [n_nodes] reference index of node          */
    float *sim_of;        /* [n_nodes] similarity at node               */
    uint64_t *pred_mask;  /* [n_nodes] incoming edges, by slot          */
    uint64_t *zero_mask;  /* [n_nodes] incoming edges whose weight is 0 */
    float *dist;          /* [n_nodes]                                  */
    int8_t *best_slot;    /* [n_nodes] -1 = no predecessor (self)       */
    int32_t *topo_pos;    /* [n_nodes] position in Kahn order           */
    int32_t *queue;       /* [n_nodes] scratch                          */
    int32_t *indeg;       /* [n_nodes] scratch                          */
    int32_t *chain;       /* [lq]      scratch                          */
} tn_state;

static inline int slot_of(const tn_state *s, int q_dst, int q_src, int a) {
    return (s->step - 1 - (q_dst - q_src)) * s->top + a;
}
static inline int src_node_of(const tn_state *s, int q_dst, int slot) {
    int o = s->step - 1 - slot / s->top;
    return (q_dst - o) * s->top + slot % s->top;
}

/* stable descending top-k of one row: ties -> lower index first */
static void row_topk(const float *row, int lr, int top, int32_t *idx, float *val) {
    int have = 0;
    for (int j = 0; j < lr; ++j) {
        float v = row[j];
        if (have == top && !(v > val[top - 1])) continue;
        int p = have < top ? have++ : top - 1;
        while (p > 0 && v > val[p - 1]) { val[p] = val[p - 1]; idx[p] = idx[p - 1]; --p; }
        val[p] = v; idx[p] = j;
    }
}

static void build_edges(tn_state *s, float min_sim) {
    const int top = s->top, step = s->step;
    uint32_t window[16];
    for (int q_src = 0; q_src < s->lq; ++q_src) {
        const int32_t *r_src = s->ref_of + q_src * top;
        for (int a = 0; a < top; ++a) window[a] = 0;  /* linked refs in [r_src[a], r_src[a]+step) */
        int q_end = q_src + step < s->lq ? q_src + step : s->lq;
        for (int q_dst = q_src + 1; q_dst < q_end; ++q_dst) {
            const int32_t *r_dst = s->ref_of + q_dst * top;
            uint32_t accepted = 0;
            for (int b = 0; b < top; ++b) {
                if (!(s->sim_of[q_dst * top + b] >= min_sim)) continue;        /* C4 */
                for (int a = 0; a < top; ++a) {
                    int d = r_dst[b] - r_src[a];
                    if (d <= 0 || d >= step) continue;                         /* C2 */
                    if (window[a] & ((2u << d) - 1u)) continue;                /* C3 */
                    s->pred_mask[q_dst * top + b] |= 1ull << slot_of(s, q_dst, q_src, a);
                    accepted |= 1u << b;
                }
            }
            for (int b = 0; b < top; ++b) {
                if (!(accepted >> b & 1)) continue;
                for (int a = 0; a < top; ++a) {
                    int d = r_dst[b] - r_src[a];
                    if (d >= 0 && d < step) window[a] |= 1u << d;
                }
            }
        }
    }
}

/* literal Kahn order with one FIFO queue (== networkx generations concatenated) */
static void kahn_positions(tn_state *s) {
    const int top = s->top, step = s->step, n = s->n_nodes;
    int tail = 0;
    for (int v = 0; v < n; ++v) {
        s->indeg[v] = __builtin_popcountll(s->pred_mask[v]);
        if (s->indeg[v] == 0) s->queue[tail++] = v;
    }
    for (int head = 0; head < tail; ++head) {
        int u = s->queue[head];
        s->topo_pos[u] = head;
        int q = u / top, a = u % top;
        for (int q_dst = q + 1; q_dst < q + step && q_dst < s->lq; ++q_dst) {
            uint64_t bit = 1ull << slot_of(s, q_dst, q, a);
            for (int b = 0; b < top; ++b) {
                int c = q_dst * top + b;
                if ((s->pred_mask[c] & bit) && --s->indeg[c] == 0) s->queue[tail++] = c;
            }
        }
    }
}

static void sweep(tn_state *s) {
    const int top = s->top;
    for (int v = 0; v < s->n_nodes; ++v) {
        uint64_t m = s->pred_mask[v];
        float best = 0.0f; int best_slot = -1;
        while (m) {
            int slot = __builtin_ctzll(m); m &= m - 1;
            float w = (s->zero_mask[v] >> slot & 1) ? 0.0f : s->sim_of[v];
            float cand = s->dist[src_node_of(s, v / top, slot)] + w;
            if (best_slot < 0 || cand > best) { best = cand; best_slot = slot; }
        }
        if (best_slot >= 0 && !(best >= 0.0f)) { best = 0.0f; best_slot = -1; }
        s->dist[v] = best; s->best_slot[v] = (int8_t)best_slot;
    }
}

int tn_fast(const float *sims, int lq, int lr, int step, int topk, int max_path,
            float min_sim, double min_length, double max_iou, int32_t *boxes_out) {
    tn_state s; memset(&s, 0, sizeof s);
    s.lq = lq; s.lr = lr; s.step = step; s.top = topk < lr ? topk : lr;
    if (lq <= 0 || s.top <= 0) return 0;
    if (s.top > 16 || (step - 1) * s.top > TN_MAX_SLOTS || step < 1 || step > 31) return -1;
    s.n_nodes = lq * s.top;
    const int n = s.n_nodes;
    s.ref_of = malloc(sizeof(int32_t) * n); s.sim_of = malloc(sizeof(float) * n);
    s.pred_mask = calloc(n, sizeof(uint64_t)); s.zero_mask = calloc(n, sizeof(uint64_t));
    s.dist = malloc(sizeof(float) * n); s.best_slot = malloc(n);
    s.topo_pos = malloc(sizeof(int32_t) * n); s.queue = malloc(sizeof(int32_t) * n);
    s.indeg = malloc(sizeof(int32_t) * n); s.chain = malloc(sizeof(int32_t) * (lq + 1));
    for (int q = 0; q < lq; ++q)
        row_topk(sims + (size_t)q * lr, lr, s.top, s.ref_of + q * s.top, s.sim_of + q * s.top);
    build_edges(&s, min_sim);
    kahn_positions(&s);

    int n_boxes = 0;
    for (int round = 0; round <= max_path; ++round) {
        sweep(&s);
        /* first maximum in topological order; the source (dist 0) precedes everything */
        int end = -1; float best = 0.0f;
        for (int v = 0; v < n; ++v) {
            float d = s.dist[v];
            if (d > best || (end >= 0 && d == best && s.topo_pos[v] < s.topo_pos[end])) {
                best = d; end = v;
            }
        }
        if (end < 0) break;  /* all zero: networkx returns [source] -> empty chain */
        int len = 0;
        for (int v = end;;) {
            s.chain[len++] = v;
            int slot = s.best_slot[v];
            if (slot < 0) break;
            s.zero_mask[v] |= 1ull << slot;
            v = src_node_of(&s, v / s.top, slot);
        }
        float score = 0.0f;
        for (int i = len - 1; i >= 0; --i) score += s.sim_of[s.chain[i]];
        int first = s.chain[len - 1], last = s.chain[0];
        int q_lo = 0, q_hi = 0, r_lo = 0, r_hi = 0;
        if (score > 0.0f) {
            /* q and (by C2) r strictly increase along a chain */
            q_lo = first / s.top; q_hi = last / s.top;
            r_lo = s.ref_of[first]; r_hi = s.ref_of[last];
        }
        double mean_extent = (double)(r_hi - r_lo + q_hi - q_lo) / 2.0;
        double worst = 0.0;
        for (int k = 0; k < n_boxes; ++k) {
            const int32_t *g = boxes_out + 4 * k;
            int64_t w = (int64_t)(q_hi < g[2] ? q_hi : g[2]) - (q_lo > g[0] ? q_lo : g[0]) + 1;
            int64_t h = (int64_t)(r_hi < g[3] ? r_hi : g[3]) - (r_lo > g[1] ? r_lo : g[1]) + 1;
            if (w < 0) w = 0;
            if (h < 0) h = 0;
            int64_t inter = w * h;
            int64_t a1 = (int64_t)(q_hi - q_lo + 1) * (r_hi - r_lo + 1);
            int64_t a2 = (int64_t)(g[2] - g[0] + 1) * (g[3] - g[1] + 1);
            double v = (double)inter / (double)(a1 + a2 - inter);
            if (k == 0 || v > worst) worst = v;
        }
        int shorter = (r_hi - r_lo) < (q_hi - q_lo) ? (r_hi - r_lo) : (q_hi - q_lo);
        if (mean_extent != 0.0 && score / (float)mean_extent > min_sim &&
            (double)shorter > min_length && worst < max_iou) {
            int32_t *o = boxes_out + 4 * n_boxes++;
            o[0] = q_lo; o[1] = r_lo; o[2] = q_hi; o[3] = r_hi;
        }
    }
    free(s.ref_of); free(s.sim_of); free(s.pred_mask); free(s.zero_mask); free(s.dist);
    free(s.best_slot); free(s.topo_pos); free(s.queue); free(s.indeg); free(s.chain);
    return n_boxes;
}

/* batch entry: pairs packed back to back at element offsets off[i] */
int tn_fast_batch(const float *sims, const int64_t *off, const int32_t *lq, const int32_t *lr,
                  int n_pairs, int step, int topk, int max_path, float min_sim,
                  double min_length, double max_iou, int32_t *boxes_out, int32_t *n_boxes_out) {
    int cap = max_path + 1;
    for (int i = 0; i < n_pairs; ++i) {
        int nb = tn_fast(sims + off[i], lq[i], lr[i], step, topk, max_path, min_sim,
                         min_length, max_iou, boxes_out + (size_t)i * cap * 4);
        if (nb < 0) return -1;
        n_boxes_out[i] = nb;
    }
    return 0;
}
