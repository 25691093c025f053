// Global-threshold search, whole schedule enqueued from C++ without a host round trip.
//
// Replaces VideoIndex._global_threshold_knn_search's engine call (vsc/index.py:142-165):
//   faiss.contrib.exhaustive_search.range_search_max_results(index, exponential_query_iterator(xq),
//                                                            radius, max_results=2K, min_results=K)
// FAISS runs a range search per exponential query batch (32, 64, ... rows), and whenever it holds more than max_results
// results it makes the (min_results+1)-th best held score the new radius and re-filters everything strictly.  The
// result is every pair beyond the FINAL radius; the radius trajectory (and with it the behaviour at exact ties) depends
// on the batch schedule, so the schedule is followed literally -- but the bookkeeping (result count, the decision to
// tighten, the radius) lives in a device-side control block: each batch is {tensor-core range search with the thresholds
// read from that block} + {decide, radix-select the radius, strict re-filter}, the last seven kernels doing nothing
// unless the batch pushed the total over max_results.  The host reads the control block once, at the end.
// If a batch emits more than the buffer holds the block says so and the caller repeats the search the host-driven way
// (vsc2022_b200/index.py), which can split batches and prune.
#include <stddef.h>

#include "search_internal.cuh"

extern "C" int vsc_search_global_topk(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k,
                                      const float *d_a_norm, const float *d_b_norm, int32_t metric_l2,
                                      int64_t max_results, int64_t min_results, float *d_score, int32_t *d_row,
                                      int32_t *d_col, float *d_score2, int32_t *d_row2, int32_t *d_col2,
                                      uint64_t capacity, void *d_control, int32_t a_row_bytes,
                                      const vsc_gemm_format *fmt, vsc_stream_t stream_) {
    using namespace vsc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m < 0 || n < 0 || !d_control || !d_score || !d_row || !d_col || !d_score2 || !d_row2 || !d_col2 ||
        max_results < min_results || min_results < 0 || a_row_bytes <= 0) {
        set_error("vsc_search_global_topk: bad arguments"); return VSC_ERR_INVALID;
    }
    if (metric_l2 && (!d_a_norm || !d_b_norm)) { set_error("vsc_search_global_topk: L2 metric needs squared norms"); return VSC_ERR_INVALID; }
    SearchControl *ctl = static_cast<SearchControl *>(d_control);
    VSC_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(SearchControl), stream));
    const float radius0 = metric_l2 ? 1e10f : -1e10f;
    const float thr0[2] = {radius0, radius0};
    VSC_CUDA_CHECK(cudaMemcpyAsync(ctl->thr, thr0, sizeof thr0, cudaMemcpyHostToDevice, stream));
    if (m == 0 || n == 0) return VSC_OK;
    const int keep_max = metric_l2 ? 0 : 1;
    // faiss.contrib.exhaustive_search.exponential_query_iterator: 32, 64, ... doubling while < 20000
    int64_t size = 32;
    for (int64_t at = 0; at < m;) {
        const int64_t rows = at + size < m ? size : m - at;
        int rc = launch_emit_device(static_cast<const char *>(d_a) + at * a_row_bytes, rows, d_b, n, k,
                                    d_a_norm ? d_a_norm + at : nullptr, d_b_norm, metric_l2, ctl->thr, at, d_score, d_row,
                                    d_col, capacity, ctl->counters, fmt, stream);
        if (rc != VSC_OK) return rc;
        rc = search_after_batch(ctl, d_score, d_row, d_col, d_score2, d_row2, d_col2, capacity, max_results, min_results,
                                keep_max, stream);
        if (rc != VSC_OK) return rc;
        at += rows;
        if (size < 20000) size *= 2;
    }
    return search_final_filter(ctl, d_score, d_row, d_col, d_score2, d_row2, d_col2, keep_max, stream);
}

// The same schedule with the large batches filtered: a batch of at least `filter_from_rows` query rows runs ONE tensor-core
// product per value pair (the hi parts: d_a_single / d_b_single over k_single) with both thresholds loosened by *d_margin --
// twice the bound of the single-product error, so nothing the exact scores would accept is missed -- into the twin buffer,
// and search_rescore_append takes the exact float32 inner products of those candidates from the original matrices
// (d_a_raw / d_b_raw, d dimensions) and counts / appends them against the real thresholds.  By then the radius is tight: a
// batch of 4096 x 200k pairs leaves ~1.5 M candidates.  The small early batches, whose radius is loose (every pair of the
// first batch is a hit), keep the three-product GEMM.  Inner product only.
extern "C" int vsc_search_global_topk_filtered(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k,
                                               const void *d_a_single, const void *d_b_single, int32_t k_single,
                                               const float *d_a_raw, int64_t lda_raw, const float *d_b_raw, int64_t ldb_raw,
                                               int32_t d, const float *d_margin, int64_t filter_from_rows,
                                               int64_t max_results, int64_t min_results, float *d_score, int32_t *d_row,
                                               int32_t *d_col, float *d_score2, int32_t *d_row2, int32_t *d_col2,
                                               uint64_t capacity, void *d_control, int32_t a_row_bytes,
                                               const vsc_gemm_format *fmt, vsc_stream_t stream_) {
    using namespace vsc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m < 0 || n < 0 || !d_control || !d_score || !d_row || !d_col || !d_score2 || !d_row2 || !d_col2 || !d_a_single ||
        !d_b_single || !d_a_raw || !d_b_raw || !d_margin || max_results < min_results || min_results < 0 || a_row_bytes <= 0) {
        set_error("vsc_search_global_topk_filtered: bad arguments"); return VSC_ERR_INVALID;
    }
    SearchControl *ctl = static_cast<SearchControl *>(d_control);
    VSC_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(SearchControl), stream));
    const float thr0[2] = {-1e10f, -1e10f};
    VSC_CUDA_CHECK(cudaMemcpyAsync(ctl->thr, thr0, sizeof thr0, cudaMemcpyHostToDevice, stream));
    if (m == 0 || n == 0) return VSC_OK;
    int64_t size = 32;
    for (int64_t at = 0; at < m;) {
        const int64_t rows = at + size < m ? size : m - at;
        int rc;
        if (rows >= filter_from_rows) {
            VSC_CUDA_CHECK(cudaMemsetAsync(ctl->cand_counters, 0, sizeof ctl->cand_counters, stream));
            rc = launch_emit_device(static_cast<const char *>(d_a_single) + at * a_row_bytes, rows, d_b_single, n, k_single,
                                    nullptr, nullptr, 0, ctl->thr, at, d_score2, d_row2, d_col2, capacity, ctl->cand_counters,
                                    fmt, stream, d_margin);
            if (rc == VSC_OK)
                rc = search_rescore_append(ctl, d_a_raw, lda_raw, d_b_raw, ldb_raw, d, d_score2, d_row2, d_col2, capacity,
                                           d_score, d_row, d_col, stream);
        } else {
            rc = launch_emit_device(static_cast<const char *>(d_a) + at * a_row_bytes, rows, d_b, n, k, nullptr, nullptr, 0,
                                    ctl->thr, at, d_score, d_row, d_col, capacity, ctl->counters, fmt, stream);
        }
        if (rc != VSC_OK) return rc;
        rc = search_after_batch(ctl, d_score, d_row, d_col, d_score2, d_row2, d_col2, capacity, max_results, min_results, 1,
                                stream);
        if (rc != VSC_OK) return rc;
        at += rows;
        if (size < 20000) size *= 2;
    }
    return search_final_filter(ctl, d_score, d_row, d_col, d_score2, d_row2, d_col2, 1, stream);
}

// ---- the schedule step by step, for the query-sharded search over several GPUs (vsc2022_b200/index.py): every rank emits ITS
// slice of a batch, the ranks all-reduce the hit count, the radix histograms and the survivor count between the phases
// (NCCL calls on views of the control block, enqueued on the same stream: still no host round trip), and every rank runs the
// same decide / pick kernels on the same numbers.
//   phase -1  begin: clear the control block, unbounded radius (arg = metric_l2), use_global = 1
//   phase 0-4 search_phase (decide; histogram / pick of radix pass `arg`; strict re-filter; copy back + finish)
//   phase 5   end: drop the emit epilogue's fillers with the final radius
extern "C" int vsc_search_step(int32_t phase, int32_t arg, void *d_control, float *d_score, int32_t *d_row, int32_t *d_col,
                               float *d_score2, int32_t *d_row2, int32_t *d_col2, uint64_t capacity, int64_t max_results,
                               int64_t min_results, int32_t keep_max, vsc_stream_t stream_) {
    using namespace vsc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!d_control) { set_error("vsc_search_step: null control block"); return VSC_ERR_INVALID; }
    SearchControl *ctl = static_cast<SearchControl *>(d_control);
    if (phase == -1) {
        VSC_CUDA_CHECK(cudaMemsetAsync(ctl, 0, sizeof(SearchControl), stream));
        const float radius0 = arg ? 1e10f : -1e10f;
        const float thr0[2] = {radius0, radius0};
        VSC_CUDA_CHECK(cudaMemcpyAsync(ctl->thr, thr0, sizeof thr0, cudaMemcpyHostToDevice, stream));
        const int32_t one = 1;
        VSC_CUDA_CHECK(cudaMemcpyAsync(&ctl->use_global, &one, sizeof one, cudaMemcpyHostToDevice, stream));
        return VSC_OK;
    }
    if (!d_score || !d_row || !d_col || !d_score2 || !d_row2 || !d_col2) { set_error("vsc_search_step: null buffer"); return VSC_ERR_INVALID; }
    if (phase == 5) return search_final_filter(ctl, d_score, d_row, d_col, d_score2, d_row2, d_col2, keep_max, stream);
    if ((phase == 1 || phase == 2) && (arg < 0 || arg > 2)) { set_error("vsc_search_step: radix pass 0..2"); return VSC_ERR_INVALID; }
    return search_phase(phase, arg, ctl, d_score, d_row, d_col, d_score2, d_row2, d_col2, capacity, max_results, min_results,
                        keep_max, stream);
}

// One range-search launch of the schedule for query rows [at, at + rows) of the panels (thresholds from the control block):
// the three-product GEMM, or -- filtered != 0, inner product -- the single-product candidate pass + exact re-score.
extern "C" int vsc_search_emit_batch(const void *d_a, const void *d_b, int64_t n, int32_t k, const void *d_a_single,
                                     const void *d_b_single, int32_t k_single, const float *d_a_raw, int64_t lda_raw,
                                     const float *d_b_raw, int64_t ldb_raw, int32_t d, const float *d_margin, int32_t filtered,
                                     const float *d_a_norm, const float *d_b_norm, int32_t metric_l2, int64_t at, int64_t rows,
                                     float *d_score, int32_t *d_row, int32_t *d_col, float *d_score2, int32_t *d_row2,
                                     int32_t *d_col2, uint64_t capacity, void *d_control, int32_t a_row_bytes,
                                     const vsc_gemm_format *fmt, vsc_stream_t stream_) {
    using namespace vsc;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (rows <= 0 || n <= 0) return VSC_OK;
    if (!d_control || !d_a || !d_b || !d_score || !d_row || !d_col) { set_error("vsc_search_emit_batch: null pointer"); return VSC_ERR_INVALID; }
    SearchControl *ctl = static_cast<SearchControl *>(d_control);
    if (filtered && !metric_l2) {
        if (!d_a_single || !d_b_single || !d_a_raw || !d_b_raw || !d_margin || !d_score2 || !d_row2 || !d_col2) {
            set_error("vsc_search_emit_batch: the filtered form needs the hi panels, the raw matrices and the twin buffer");
            return VSC_ERR_INVALID;
        }
        VSC_CUDA_CHECK(cudaMemsetAsync(ctl->cand_counters, 0, sizeof ctl->cand_counters, stream));
        int rc = launch_emit_device(static_cast<const char *>(d_a_single) + at * a_row_bytes, rows, d_b_single, n, k_single,
                                    nullptr, nullptr, 0, ctl->thr, at, d_score2, d_row2, d_col2, capacity, ctl->cand_counters,
                                    fmt, stream, d_margin);
        if (rc != VSC_OK) return rc;
        return search_rescore_append(ctl, d_a_raw, lda_raw, d_b_raw, ldb_raw, d, d_score2, d_row2, d_col2, capacity, d_score,
                                     d_row, d_col, stream);
    }
    return launch_emit_device(static_cast<const char *>(d_a) + at * a_row_bytes, rows, d_b, n, k,
                              d_a_norm ? d_a_norm + at : nullptr, d_b_norm, metric_l2, ctl->thr, at, d_score, d_row, d_col,
                              capacity, ctl->counters, fmt, stream);
}

// byte offsets inside the control block of what the sharded search all-reduces / reads:
// {hit count of the last launch (u64), radix histogram (2048 x u32), survivors of the re-filter (u64), their global count (u64),
//  overflow flag (i32), radius (f32), survivors held (u64)}
extern "C" int vsc_search_control_layout(int32_t *out7) {
    using vsc::SearchControl;
    if (!out7) return VSC_ERR_INVALID;
    out7[0] = (int32_t)offsetof(SearchControl, counters) + 8;
    out7[1] = (int32_t)offsetof(SearchControl, hist);
    out7[2] = (int32_t)offsetof(SearchControl, kept);
    out7[3] = (int32_t)offsetof(SearchControl, kept_global);
    out7[4] = (int32_t)offsetof(SearchControl, overflow);
    out7[5] = (int32_t)offsetof(SearchControl, thr);
    out7[6] = (int32_t)offsetof(SearchControl, counters);
    return VSC_OK;
}

extern "C" int vsc_search_control_bytes(void) { return (int)sizeof(vsc::SearchControl); }
