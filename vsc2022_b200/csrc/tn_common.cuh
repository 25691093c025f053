// Shared declarations of the TN alignment engine (fast pipeline + general fused kernel).
#pragma once
#include "common.cuh"

namespace vsc {
namespace tn {

constexpr int kMaxTop = 8;     // tn_top_k limit
constexpr int kMaxBoxes = 32;  // max_path + 1 limit

// One batch of candidate pairs and where results go (all device pointers).
struct Batch {
    const float *sims;
    const int64_t *off;
    const int32_t *lq;
    const int32_t *lr;
    int n_pairs;
    int step, topk, max_path;
    float min_sim;
    double min_length, max_iou;
    int32_t *boxes;      // [n_pairs][max_path+1][4]
    int32_t *n_boxes;    // [n_pairs]
    float *box_maxsim;   // [n_pairs][max_path+1] or null
    int32_t *status;     // [n_pairs] or null: which kernel produced the result
    int max_nodes, max_lq, max_lr;
    // Optional node records of the fast pipeline (Workspace layout, node (q, rank) at [pair * max_nodes + q * topk + rank]).
    // When set, the general kernel reads its row top-K from them instead of scanning a similarity matrix.
    const uint16_t *node_ref;
    const void *node_rec;
    int node_rec_bytes, node_sim_off;
};

// A device-side work list: count[0] entries in list[].  in == nullptr means "all pairs".
struct WorkList {
    int32_t *count;
    int32_t *list;
};

// ------------------------------------------------------------------ workspace of the fast pipeline
// Written by the row top-K stage (tn_topk_kernel for similarity matrices in memory, pair_topk_kernel when the
// similarities come straight out of tensor memory), read by the graph stage.
struct Workspace {
    uint16_t *ref_of;   // [P][N] reference frame of node | kSimOk
    void *rec;          // [P][N] NodeRec: {predecessor mask, zeroed-edge mask, similarity, distance}
    uint16_t *gen;      // [P][N] Kahn generation
    uint8_t *last_parent;  // [P][N] predecessor slot of the node's last parent in Kahn order (filled on demand by the DP)
    int rec_bytes, sim_off;  // record size and byte offset of the similarity inside it
    int32_t *skip;      // [P] 1 = handed to the general kernel
    int32_t *cursor;    // T1 tile counter
    // pairs whose end-node tie only the full Kahn order can break go straight to the exact-order kernel's work list
    int32_t *exact_count, *exact_list;
};

template <typename MaskT>
struct alignas(16) NodeRec {
    MaskT pred, zero;
    float sim, dist;
};
static_assert(sizeof(NodeRec<uint32_t>) == 16 && sizeof(NodeRec<uint64_t>) == 32, "record layout");
// ref_of entry: reference frame (< 2^15: the pipeline takes rows up to 512 columns) | flag "similarity >= min_sim"
// (constraint C4, evaluated once by the top-K stage where the similarity is in a register)
constexpr uint16_t kRefMask = 0x7FFF, kSimOk = 0x8000;

// Descriptor panels of a batch whose similarities are computed on the fly (pair_gemm.cu): pair p multiplies rows
// [q_start[p], +lq[p]) of the query panel by rows [r_start[p], +lr[p]) of the reference panel and adds `bias`.
struct PairOperands {
    const void *q_panel, *r_panel;   // bf16 [rows][k], K-major (vsc_prepare_operand)
    int64_t q_rows, r_rows;
    int k;
    const int32_t *q_start, *r_start;
    float bias;
    int ab_f16;                      // vsc_gemm_format
    int64_t ldq, ldr;
    const float *out_scale;
};
int launch_pair_topk(const PairOperands &op, const Batch &b, const Workspace &w, const WorkList &out, float *sims,
                     const int64_t *off, int64_t pair_stride, cudaStream_t stream);
bool pair_topk_supported(const Batch &b);

// General kernel (tn_fused.cu): any shape / alignment.  exact_order=false breaks end-node ties by
// Kahn generation and appends unresolved pairs to `out`; exact_order=true computes full Kahn
// positions and always finishes.  status code written per finished pair.
int launch_fused(const Batch &b, bool exact_order, const WorkList *in, const WorkList *out,
                 int status_code, cudaStream_t stream);

// Fast pipeline (tn_pipeline.cu): aligned rows (lr % 4 == 0, 16-byte aligned start, lr <= 512,
// lr >= topk).  Pairs it cannot finish are appended to `out`.
int launch_pipeline(const Batch &b, const Workspace &w, const WorkList &out, cudaStream_t stream);
bool pipeline_supported(const Batch &b);
// Same pipeline with the row top-K taken straight from descriptor panels (pair_gemm.cu) instead of matrices in memory.
int launch_pipeline_from_features(const PairOperands &op, const Batch &b, const Workspace &w, const WorkList &out,
                                  cudaStream_t stream);
bool graph_supported(const Batch &b);
// Graph stage on a compact graph (tn_graph.cu): narrow predecessor masks ((step-1)*topk <= 32)
bool graph_v2_supported(const Batch &b);
size_t graph_v2_scratch_bytes(const Batch &b);
int launch_graph_v2(const Batch &b, const Workspace &w, const WorkList &out, unsigned char *graphs, cudaStream_t stream,
                    void (*mark_between)(cudaStream_t));
int workspace_alloc(const Batch &b, Workspace *w, void **base_out, cudaStream_t stream);

}  // namespace tn
}  // namespace vsc
