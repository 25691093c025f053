/*
 * oracle/tn_graph_model.c -- CPU model of the COMPACT-GRAPH formulation of VCSL's temporal network that
 * vsc2022_b200/csrc/tn_graph.cu runs on the GPU.  TEST INFRASTRUCTURE: it exists so that the algorithm of the kernels
 * (not just their output on a GPU box) is pinned against oracle/tn_fast.c on thousands of seeded matrices by the CPU
 * test suite.  Same data structures and the same order of operations as the kernels, executed sequentially.
 *
 * Authority for the results: oracle/tn_networkx.py (via tn_fast.c, which this file includes for row_topk and as the
 * comparison target).  Reference call site: vsc/baseline/localization.py:44-46,58.
 *
 * The formulation (what differs from tn_fast.c's layer-by-layer sweeps):
 *   E  edges by bitmap screening: every query row keeps a bitmap of the references of its top-K nodes; a source node
 *      (q, a) finds its C2 candidates in row q+o by extracting bits (r_src, r_src + step) of that row's bitmap.
 *      C3 / C4 and the predecessor-slot numbering are tn_fast.c's.
 *   C  compaction: ACTIVE nodes (at least one predecessor) in node order, their incoming edges as a CSR list in slot
 *      order; SOURCE nodes that feed an edge are appended after the active ones (distance 0 for ever).
 *      Typical 300x300 pair: 1500 nodes, ~350 active, ~380 edges, only ~100 edges between two active nodes.
 *   D  longest paths by Kahn GENERATION instead of by row layer: generation 1 (all predecessors are sources) is relaxed
 *      in one parallel step; the ~100 "inner" nodes follow generation by generation (G ~ 5 without a copy, ~40 with one,
 *      instead of 300 layers).  After a chain is extracted only nodes downstream of it are relaxed again, in generation
 *      order, up to the farthest generation a changed node can reach.  End node = largest distance, ties by smaller
 *      generation; a tie that generations cannot break (or a pair that exceeds the fixed-size tables) is reported with
 *      status 3 and goes to the exact-order kernel.
 */
#include "tn_fast.c"

#define GM_MAX_ACTIVE 448
#define GM_MAX_SOURCES 448
#define GM_MAX_EDGES 576
#define GM_MAX_GEN 255
#define GM_MAX_CHAIN 256

typedef struct {
    int A, S, E, I;                     /* active nodes, sources, edges, inner nodes */
    uint16_t node[GM_MAX_ACTIVE + GM_MAX_SOURCES];   /* node id = q*top + rank */
    uint16_t ref[GM_MAX_ACTIVE + GM_MAX_SOURCES];
    float sim[GM_MAX_ACTIVE + GM_MAX_SOURCES];
    uint16_t eoff[GM_MAX_ACTIVE + 1];
    uint16_t esrc[GM_MAX_EDGES];        /* entry index: < A active, >= A source */
    uint8_t inner[GM_MAX_ACTIVE];       /* 1 = has an active predecessor */
    /* D state */
    float dist[GM_MAX_ACTIVE];
    uint8_t gen[GM_MAX_ACTIVE];
    uint8_t reach[GM_MAX_ACTIVE];       /* largest generation among the node's successors (0: none) */
    int8_t best[GM_MAX_ACTIVE];         /* rank of the best edge inside the node's edge list, -1 = none */
    uint8_t flag[GM_MAX_ACTIVE];        /* bit 0: an incoming edge was zeroed this round, bit 1: distance changed this round */
    uint8_t ezero[GM_MAX_EDGES];
    uint16_t order[GM_MAX_ACTIVE];      /* inner nodes sorted by generation */
    uint16_t gstart[GM_MAX_GEN + 2];
} gm_graph;

/* ---- E + C: build the compact graph from the node table.  Returns 0, or 3 when a table overflows. */
static int gm_build(const tn_state *s, float min_sim, gm_graph *g) {
    const int top = s->top, step = s->step, lq = s->lq, n = s->n_nodes;
    const int words = (s->lr + step + 31) / 32 + 1;
    uint32_t *rowbits = calloc((size_t)lq * words, sizeof(uint32_t));
    uint32_t *pred = calloc(n, sizeof(uint32_t));
    for (int v = 0; v < n; ++v) {
        int r = s->ref_of[v];
        rowbits[(size_t)(v / top) * words + (r >> 5)] |= 1u << (r & 31);
    }
    for (int q_src = 0; q_src < lq; ++q_src) {
        uint32_t window[16];
        const int32_t *r_src = s->ref_of + q_src * top;
        for (int a = 0; a < top; ++a) window[a] = 0;
        for (int o = 1; o < step && q_src + o < lq; ++o) {
            const int q_dst = q_src + o;
            const uint32_t *bits = rowbits + (size_t)q_dst * words;
            uint32_t accepted = 0;   /* destination ranks linked in this step */
            for (int a = 0; a < top; ++a) {
                /* references r_src[a]+1 .. r_src[a]+step-1 present in the destination row (C2) */
                const int lo = r_src[a] + 1;
                const uint64_t two = (uint64_t)bits[lo >> 5] | ((uint64_t)bits[(lo >> 5) + 1] << 32);
                uint32_t hit = (uint32_t)(two >> (lo & 31)) & ((1u << (step - 1)) - 1u);
                while (hit) {
                    const int d = __builtin_ctz(hit) + 1; hit &= hit - 1;
                    const int rd = r_src[a] + d;
                    int b = 0;
                    while (s->ref_of[q_dst * top + b] != rd) ++b;              /* rank of that reference */
                    if (!(s->sim_of[q_dst * top + b] >= min_sim)) continue;    /* C4 */
                    if (window[a] & ((2u << d) - 1u)) continue;                /* C3 */
                    pred[q_dst * top + b] |= 1u << ((step - 1 - o) * top + a);
                    accepted |= 1u << b;
                }
            }
            while (accepted) {   /* references linked in this step constrain the later destination rows */
                const int b = __builtin_ctz(accepted); accepted &= accepted - 1;
                const int rd = s->ref_of[q_dst * top + b];
                for (int a = 0; a < top; ++a) {
                    const int d = rd - r_src[a];
                    if (d >= 0 && d < step) window[a] |= 1u << d;
                }
            }
        }
    }
    /* compaction */
    int *idx = malloc(sizeof(int) * n);
    int A = 0, E = 0, S = 0, rc = 0;
    for (int v = 0; v < n; ++v) idx[v] = -1;
    for (int v = 0; v < n; ++v)
        if (pred[v]) { if (A < GM_MAX_ACTIVE) idx[v] = A; ++A; E += __builtin_popcount(pred[v]); }
    if (A > GM_MAX_ACTIVE || E > GM_MAX_EDGES) rc = 3;
    if (!rc) {
        int e = 0;
        for (int v = 0; v < n && !rc; ++v) {
            if (!pred[v]) continue;
            const int i = idx[v];
            g->node[i] = (uint16_t)v; g->ref[i] = (uint16_t)s->ref_of[v]; g->sim[i] = s->sim_of[v];
            g->eoff[i] = (uint16_t)e; g->inner[i] = 0;
            uint32_t m = pred[v];
            while (m) {
                const int slot = __builtin_ctz(m); m &= m - 1;
                const int u = (v / top - (step - 1 - slot / top)) * top + slot % top;
                if (idx[u] < 0) {          /* a source seen for the first time */
                    if (S >= GM_MAX_SOURCES) { rc = 3; break; }
                    idx[u] = GM_MAX_ACTIVE + S; ++S;   /* provisional: sources are renumbered to A + k below */
                }
                if (idx[u] < GM_MAX_ACTIVE && pred[u]) g->inner[i] = 1;
                g->esrc[e++] = (uint16_t)idx[u];
            }
        }
        g->eoff[A] = (uint16_t)E;
        if (!rc) {
            for (int v = 0; v < n; ++v) {
                if (idx[v] >= GM_MAX_ACTIVE) {
                    const int k = A + (idx[v] - GM_MAX_ACTIVE);
                    g->node[k] = (uint16_t)v; g->ref[k] = (uint16_t)s->ref_of[v]; g->sim[k] = s->sim_of[v];
                }
            }
            for (int x = 0; x < E; ++x)
                if (g->esrc[x] >= GM_MAX_ACTIVE) g->esrc[x] = (uint16_t)(A + (g->esrc[x] - GM_MAX_ACTIVE));
        }
    }
    g->A = A; g->S = S; g->E = E;
    free(idx); free(pred); free(rowbits);
    return rc;
}

/* one node against the current distances: FIRST maximal predecessor in slot order; negative best -> (0, none) */
static inline int gm_relax(gm_graph *g, int i) {
    float best = 0.0f; int bs = -1;
    for (int e = g->eoff[i]; e < g->eoff[i + 1]; ++e) {
        const int u = g->esrc[e];
        const float cand = (u < g->A ? g->dist[u] : 0.0f) + (g->ezero[e] ? 0.0f : g->sim[i]);
        if (bs < 0 || cand > best) { best = cand; bs = e - g->eoff[i]; }
    }
    if (bs >= 0 && !(best >= 0.0f)) { best = 0.0f; bs = -1; }
    best += 0.0f;   /* -0 -> +0: distances are compared by their bits */
    const int changed = memcmp(&best, &g->dist[i], sizeof(float)) != 0;
    g->dist[i] = best; g->best[i] = (int8_t)bs;
    return changed;
}

int tn_graph_model(const float *sims, int lq, int lr, int step, int topk, int max_path,
                   float min_sim, double min_length, double max_iou, int32_t *boxes_out, int32_t *status_out) {
    tn_state s; memset(&s, 0, sizeof s);
    *status_out = 0;
    s.lq = lq; s.lr = lr; s.step = step; s.top = topk < lr ? topk : lr;
    if (lq <= 0 || s.top <= 0) return 0;
    if ((step - 1) * s.top > 32 || step < 1 || step > 31 || s.top > 8) return -1;   /* narrow masks only */
    s.n_nodes = lq * s.top;
    const int n = s.n_nodes, top = s.top;
    s.ref_of = malloc(sizeof(int32_t) * n); s.sim_of = malloc(sizeof(float) * n);
    for (int q = 0; q < lq; ++q)
        row_topk(sims + (size_t)q * lr, lr, top, s.ref_of + q * top, s.sim_of + q * top);
    gm_graph *g = calloc(1, sizeof(gm_graph));
    int n_boxes = 0;
    int status = gm_build(&s, min_sim, g);
    if (status) goto done;
    const int A = g->A;

    /* ---- first sweep.  Generation 1: no active predecessor. */
    int n_inner = 0, G = 1;
    for (int i = 0; i < A; ++i) {
        g->dist[i] = 0.0f; g->flag[i] = 0; g->reach[i] = 0;
        if (!g->inner[i]) { gm_relax(g, i); g->gen[i] = 1; }
        else { g->gen[i] = 0; ++n_inner; }
    }
    /* later generations: a node is ready when all its active predecessors are done (gen != 0) */
    for (int done = 0, cur = 2; done < n_inner; ++cur) {
        if (cur > GM_MAX_GEN) { status = 3; goto done; }
        int ready[GM_MAX_ACTIVE], nr = 0;
        for (int i = 0; i < A; ++i) {
            if (!g->inner[i] || g->gen[i]) continue;
            int ok = 1;
            for (int e = g->eoff[i]; e < g->eoff[i + 1]; ++e) {
                const int u = g->esrc[e];
                if (u < A && (g->gen[u] == 0 || g->gen[u] >= cur)) { ok = 0; break; }
            }
            if (ok) ready[nr++] = i;
        }
        for (int k = 0; k < nr; ++k) {   /* parallel on the GPU: nodes of one generation are independent */
            const int i = ready[k];
            gm_relax(g, i);
            g->gen[i] = (uint8_t)cur;
            for (int e = g->eoff[i]; e < g->eoff[i + 1]; ++e) {
                const int u = g->esrc[e];
                if (u < A && g->reach[u] < cur) g->reach[u] = (uint8_t)cur;
            }
        }
        done += nr; G = cur;
    }
    /* generation-1 nodes also need `reach`: done above only for predecessors of inner nodes, which covers them */
    /* inner nodes sorted by generation (counting sort) */
    for (int k = 0; k <= G + 1; ++k) g->gstart[k] = 0;
    for (int i = 0; i < A; ++i) if (g->inner[i]) ++g->gstart[g->gen[i] + 1];
    for (int k = 1; k <= G + 1; ++k) g->gstart[k] += g->gstart[k - 1];
    {
        uint16_t fill[GM_MAX_GEN + 2];
        memcpy(fill, g->gstart, sizeof(uint16_t) * (G + 2));
        for (int i = 0; i < A; ++i) if (g->inner[i]) g->order[fill[g->gen[i]]++] = (uint16_t)i;
    }

    for (int round = 0; round <= max_path; ++round) {
        /* ---- end node: largest distance, then smallest generation; unresolved tie -> exact-order kernel */
        uint32_t bk = 0; int bg = 1 << 30, bi = -1, cnt = 0;
        for (int i = 0; i < A; ++i) {
            uint32_t k; memcpy(&k, &g->dist[i], 4);
            if (k == 0) continue;
            if (k > bk || (k == bk && g->gen[i] < bg)) { bk = k; bg = g->gen[i]; bi = i; cnt = 1; }
            else if (k == bk && g->gen[i] == bg) ++cnt;
        }
        if (bi < 0) break;
        if (cnt > 1) { status = 3; goto done; }
        /* ---- walk the chain back, mark its edges spent */
        int chain[GM_MAX_CHAIN], len = 0, first_entry = -1;
        for (int i = bi;;) {
            if (len >= GM_MAX_CHAIN - 1) { status = 3; goto done; }
            chain[len++] = i;
            if (g->best[i] < 0) { first_entry = i; break; }
            const int e = g->eoff[i] + g->best[i];
            g->ezero[e] = 1; g->flag[i] |= 1;
            const int u = g->esrc[e];
            if (u >= A) { first_entry = u; chain[len++] = u; break; }
            i = u;
        }
        /* chain[] = end ... first (entry indices; the first may be a source) */
        float score = 0.0f;
        for (int k = len - 1; k >= 0; --k) score += g->sim[chain[k]];
        const int last_entry = chain[0];
        int q_lo = 0, q_hi = 0, r_lo = 0, r_hi = 0;
        if (score > 0.0f) {
            q_lo = g->node[first_entry] / top; q_hi = g->node[last_entry] / top;
            r_lo = g->ref[first_entry]; r_hi = g->ref[last_entry];
        }
        const double mean_extent = (double)(r_hi - r_lo + q_hi - q_lo) / 2.0;
        double worst = 0.0;
        for (int k = 0; k < n_boxes; ++k) {
            const int32_t *b = boxes_out + 4 * k;
            int64_t w = (int64_t)(q_hi < b[2] ? q_hi : b[2]) - (q_lo > b[0] ? q_lo : b[0]) + 1;
            int64_t h = (int64_t)(r_hi < b[3] ? r_hi : b[3]) - (r_lo > b[1] ? r_lo : b[1]) + 1;
            if (w < 0) w = 0;
            if (h < 0) h = 0;
            const int64_t inter = w * h;
            const int64_t a1 = (int64_t)(q_hi - q_lo + 1) * (r_hi - r_lo + 1);
            const int64_t a2 = (int64_t)(b[2] - b[0] + 1) * (b[3] - b[1] + 1);
            const double v = (double)inter / (double)(a1 + a2 - inter);
            if (k == 0 || v > worst) worst = v;
        }
        const int shorter = (r_hi - r_lo) < (q_hi - q_lo) ? (r_hi - r_lo) : (q_hi - q_lo);
        if (mean_extent != 0.0 && score / (float)mean_extent > min_sim && (double)shorter > min_length && worst < max_iou) {
            int32_t *o = boxes_out + 4 * n_boxes++;
            o[0] = q_lo; o[1] = r_lo; o[2] = q_hi; o[3] = r_hi;
        }
        if (round == max_path) break;

        /* ---- relax again what the spent edges can change, in generation order.
         * The chain's destination nodes carry flag bit 0; generation-1 nodes among them (at most the first one, plus
         * none else: a chain's generations strictly increase) are relaxed directly, inner nodes through `order`. */
        int horizon = 0;
        for (int k = len - 1; k >= 0; --k) {
            const int i = chain[k];
            if (i >= A || !(g->flag[i] & 1)) continue;
            if (!g->inner[i]) {
                g->flag[i] = 0;
                if (gm_relax(g, i)) { g->flag[i] = 2; if (g->reach[i] > horizon) horizon = g->reach[i]; }
            } else if (g->gen[i] > horizon) {
                horizon = g->gen[i];
            }
        }
        for (int cur = 2; cur <= horizon; ++cur) {
            for (int k = g->gstart[cur]; k < g->gstart[cur + 1]; ++k) {   /* parallel: one generation */
                const int i = g->order[k];
                int affected = g->flag[i] & 1;
                for (int e = g->eoff[i]; e < g->eoff[i + 1] && !affected; ++e) {
                    const int u = g->esrc[e];
                    if (u < A && (g->flag[u] & 2)) affected = 1;
                }
                g->flag[i] &= ~1;
                if (affected && gm_relax(g, i)) { g->flag[i] |= 2; if (g->reach[i] > horizon) horizon = g->reach[i]; }
            }
        }
        for (int i = 0; i < A; ++i) g->flag[i] = 0;   /* the GPU clears only what it set */
    }
done:
    *status_out = status;
    free(g); free(s.ref_of); free(s.sim_of);
    return status ? 0 : n_boxes;
}

int tn_graph_model_batch(const float *sims, const int64_t *off, const int32_t *lq, const int32_t *lr,
                         int n_pairs, int step, int topk, int max_path, float min_sim,
                         double min_length, double max_iou, int32_t *boxes_out, int32_t *n_boxes_out, int32_t *status_out) {
    const int cap = max_path + 1;
    for (int i = 0; i < n_pairs; ++i) {
        int nb = tn_graph_model(sims + off[i], lq[i], lr[i], step, topk, max_path, min_sim, min_length, max_iou,
                                boxes_out + (size_t)i * cap * 4, status_out + i);
        if (nb < 0) return -1;
        n_boxes_out[i] = nb;
    }
    return 0;
}
