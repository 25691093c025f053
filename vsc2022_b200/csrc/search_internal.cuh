// Device-side control block of the global-threshold search (search.cu, select.cu, gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace vsc {

struct SelectState {          // radix selection: lives in device memory between the passes
    uint32_t prefix;          // key bits decided so far (aligned to the top)
    uint32_t mask;            // which bits are decided
    unsigned long long k;     // rank still to find inside the matching elements (1 = best)
};

// FAISS's range_search_max_results bookkeeping, kept on the device so that the host can enqueue the whole
// exponential-batch schedule without reading anything back (vsc/index.py:142-165 drives it from Python over FAISS).
struct SearchControl {
    float thr[2];                      // [0] radius (count threshold), [1] emit threshold (the same value here)
    int32_t do_tighten, overflow;
    unsigned long long counters[2];    // [0] slots claimed in the survivor buffer, [1] hits counted by the last launch
    unsigned long long total;          // results FAISS would hold now
    unsigned long long kept;           // output count of the running re-filter
    unsigned long long n_tighten;
    SelectState sel;
    unsigned int hist[2048];
    // filtered batches (search.cu): candidates of the single-product pass, claimed slots / counted
    unsigned long long cand_counters[2];
    // query-sharded search over several GPUs (vsc_search_step): the survivor count over ALL ranks after a re-filter, written
    // by the caller (all-reduced copy of `kept`) before the finish step when use_global is set
    unsigned long long kept_global;
    int32_t use_global, pad_;
};

int launch_emit_device(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, const float *d_a_norm,
                       const float *d_b_norm, int32_t metric_l2, const float *d_thr, int64_t row_offset, float *d_score,
                       int32_t *d_row, int32_t *d_col, uint64_t capacity, unsigned long long *d_counters,
                       const vsc_gemm_format *fmt, cudaStream_t stream, const float *d_margin = nullptr);
// exact float32 scores of the candidates of a single-product pass; those beyond the thresholds of `ctl` are counted / appended
int search_rescore_append(SearchControl *ctl, const float *d_a_raw, int64_t lda, const float *d_b_raw, int64_t ldb, int32_t d,
                          const float *cand_s, const int32_t *cand_r, const int32_t *cand_c, uint64_t capacity, float *s,
                          int32_t *r, int32_t *c, cudaStream_t stream);
int search_after_batch(SearchControl *ctl, float *s, int32_t *r, int32_t *c, float *s2, int32_t *r2, int32_t *c2,
                       uint64_t capacity, int64_t max_results, int64_t min_results, int keep_max, cudaStream_t stream);
int search_phase(int phase, int arg, SearchControl *ctl, float *s, int32_t *r, int32_t *c, float *s2, int32_t *r2, int32_t *c2,
                 uint64_t capacity, int64_t max_results, int64_t min_results, int keep_max, cudaStream_t stream);
int search_final_filter(SearchControl *ctl, float *s, int32_t *r, int32_t *c, float *s2, int32_t *r2, int32_t *c2,
                        int keep_max, cudaStream_t stream);

}  // namespace vsc
