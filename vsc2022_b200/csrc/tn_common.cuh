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
};

// A device-side work list: count[0] entries in list[].  in == nullptr means "all pairs".
struct WorkList {
    int32_t *count;
    int32_t *list;
};

// General kernel (tn_fused.cu): any shape / alignment.  exact_order=false breaks end-node ties by
// Kahn generation and appends unresolved pairs to `out`; exact_order=true computes full Kahn
// positions and always finishes.  status code written per finished pair.
int launch_fused(const Batch &b, bool exact_order, const WorkList *in, const WorkList *out,
                 int status_code, cudaStream_t stream);

// Fast pipeline (tn_pipeline.cu): aligned rows (lr % 4 == 0, 16-byte aligned start, lr <= 512,
// lr >= topk).  Pairs it cannot finish are appended to `out`.
int launch_pipeline(const Batch &b, const WorkList &out, cudaStream_t stream);
bool pipeline_supported(const Batch &b);

}  // namespace tn
}  // namespace vsc
