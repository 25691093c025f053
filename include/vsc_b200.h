/*
 * vsc_b200.h -- C ABI of libvsc_b200.so, the B200 (sm_100a) engine behind the
 * vsc2022 hot path.  Plain pointers and sizes only; every pointer whose name
 * starts with d_ is a DEVICE pointer, everything else is host memory.  All
 * entry points return 0 on success and a negative code on failure;
 * vsc_last_error() then describes the failure (thread-local string).
 *
 * Each entry point names the reference interface it replaces
 * (paths relative to the facebookresearch/vsc2022 tree).
 */
#ifndef VSC_B200_H
#define VSC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSC_OK 0
#define VSC_ERR_INVALID (-1)    /* bad argument / unsupported parameter combination */
#define VSC_ERR_CUDA (-2)       /* a CUDA runtime call failed                        */
#define VSC_ERR_CAPACITY (-3)   /* problem does not fit the kernel's on-chip budget  */

typedef void *vsc_stream_t;     /* cudaStream_t */

/* Operand format of the tensor-core GEMM entry points (see vsc_prepare_operand_f16); NULL = bf16 panels with row
 * stride k and no output scale. */
typedef struct {
    int32_t ab_f16;            /* 0: bf16 panels, 1: fp16 panels */
    int64_t lda, ldb;          /* row strides in elements (0: = k) */
    const float *d_out_scale;  /* device scalar, NULL = 1 */
} vsc_gemm_format;

const char *vsc_last_error(void);
int vsc_abi_version(void);

/* -------------------------------------------------------------------------
 * Stage C: temporal-network alignment.
 * Replaces vcsl.vta.build_vta_model("TN", **cfg).forward_sim(data), the engine
 * called from vsc/baseline/localization.py:44-46,58 (VCSL vcsl/vta.py `tn`).
 *
 * Pair p is the row-major float32 matrix d_sims[d_off[p] .. + lq[p]*lr[p]).
 * Output per pair: up to (max_path+1) boxes [q_min, r_min, q_max, r_max]
 * (inclusive frame indices) in discovery order, their count, and for every box
 * max(sims[q_min:q_max, r_min:r_max]) (EXCLUSIVE upper bounds, the slice
 * vsc/baseline/localization.py:88-91 scores with; bias NOT subtracted).
 * d_box_maxsim may be NULL.
 * d_status[p] (may be NULL) says which kernel finished pair p:
 *   0 = fast pipeline (rows 16-byte aligned, lr % 4 == 0, lr <= 512, lr >= tn_top_k),
 *   2 = general kernel (any shape/alignment, tie-heavy rows),
 *   1 = exact-order kernel (a tie between unrelated graph nodes needed the full Kahn order).
 * All three produce identical results; the split is a performance detail.
 * ------------------------------------------------------------------------- */
typedef struct {
    int32_t tn_max_step;  /* VCSL default 10; vsc passes 5 (sscd_baseline.py:121,131) */
    int32_t tn_top_k;     /* 5 */
    int32_t max_path;     /* 10 */
    float min_sim;        /* 0.2 (compared in float32, as numpy does) */
    double min_length;    /* VCSL default 5; vsc passes 4 */
    double max_iou;       /* 0.3 */
} vsc_tn_params;

int vcsl_tn_batch(const float *d_sims, const int64_t *d_off, const int32_t *d_lq,
                  const int32_t *d_lr, int32_t n_pairs, int32_t max_lq, int32_t max_lr,
                  const vsc_tn_params *params, int32_t *d_boxes, int32_t *d_n_boxes,
                  float *d_box_maxsim, int32_t *d_status, int32_t force_exact_order,
                  vsc_stream_t stream);

/* Stage C straight from frame descriptors: the batch form of VCSLLocalization.localize_all's
 *     sims = [(key, np.matmul(q.feature, r.feature.T) + similarity_bias) ...]; model.forward_sim(sims)
 * (vsc/baseline/localization.py:33-36,49-54,57-58).  d_q_panel / d_r_panel are K-major bf16 panels of ALL query /
 * reference frames (vsc_prepare_operand: k = kpad, or 3*kpad for the split panels); pair p multiplies panel rows
 * [d_q_start[p], +d_lq[p]) by [d_r_start[p], +d_lr[p]).
 *
 * vsc_pair_similarity writes the matrices (row-major lq x lr at element offset d_off[p]).
 * vcsl_tn_batch_from_features aligns them: when every pair has tn_top_k <= lr <= 512 (min_lr / max_lr, host values)
 * each 128-row accumulator tile is consumed out of tensor memory by the row top-K and the matrices are written only if
 * d_sims_out is given (then d_off is required) or MaxSim scores are requested (d_box_maxsim != NULL; scratch memory
 * when d_sims_out is NULL).  Other shapes compute the matrices first and run vcsl_tn_batch on them.  Outputs as
 * vcsl_tn_batch. */
int vsc_pair_similarity(const void *d_q_panel, int64_t q_rows, const void *d_r_panel, int64_t r_rows, int32_t k,
                        const int32_t *d_q_start, const int32_t *d_lq, const int32_t *d_r_start, const int32_t *d_lr,
                        int32_t n_pairs, int32_t max_lq, int32_t max_lr, float bias, float *d_sims, const int64_t *d_off,
                        const vsc_gemm_format *fmt, vsc_stream_t stream);
int vcsl_tn_batch_from_features(const void *d_q_panel, int64_t q_rows, const void *d_r_panel, int64_t r_rows, int32_t k,
                                const int32_t *d_q_start, const int32_t *d_lq, const int32_t *d_r_start,
                                const int32_t *d_lr, int32_t n_pairs, int32_t max_lq, int32_t max_lr, int32_t min_lr,
                                float similarity_bias, const vsc_tn_params *params, float *d_sims_out,
                                const int64_t *d_off, int32_t *d_boxes, int32_t *d_n_boxes, float *d_box_maxsim,
                                int32_t *d_status, int32_t force_exact_order, const vsc_gemm_format *fmt,
                                vsc_stream_t stream);

/* Per-stage device timing of the TN fast pipeline (CUDA events recorded on the caller's stream
 * around each stage).  vsc_tn_last_stage_ms fills {row top-K, edges, sweeps, MaxSim} in ms for the
 * most recent vcsl_tn_batch call; synchronise the stream first. */
int vsc_tn_set_profiling(int on);
/* Graph stage of the fast pipeline: 0 = layer-by-layer sweeps (default), 1 = compact graph relaxed by Kahn generation
 * (csrc/tn_graph.cu; parameter sets with (tn_max_step-1)*tn_top_k <= 32).  Identical results. */
/* Pairs per warp of the longest-path kernel: 0 = chosen from the batch size (default), 1 / 2 / 4 forced (tests, tuning). */
int vsc_tn_set_dp_pairs_per_warp(int pairs);
int vsc_tn_set_graph_variant(int variant);
int vsc_tn_last_stage_ms(float *out4);

/* Development aid: DP phase clocks summed over warps since the last call (reading resets them):
 * SM cycles in {first sweep, end-node search, chain walk, zero + score, box filter, incremental sweeps},
 * incremental layer steps, warps.  All zero unless the library was built with -DVSC_TN_COUNTERS. */
int vsc_tn_debug_counters(unsigned long long *out8);

/* -------------------------------------------------------------------------
 * Stage B: descriptor similarity on tensor cores (tcgen05 / TMEM / TMA).
 * Replaces the FAISS IndexFlat arithmetic behind vsc/index.py:142-177 and
 * vsc/baseline/score_normalization.py:87-96.
 *
 * Operands are K-major bf16 panels [rows][k] (k a multiple of 64, 16-byte aligned) produced by
 * vsc_prepare_operand from fp32 descriptors x[n][d] (row stride ld):
 *   mode 0  bf16(x), zero padded to kpad                      -> [n][kpad]
 *   mode 1  query side of the 3-term split  [hi | hi | lo]     -> [n][3*kpad]
 *   mode 2  reference side of the split     [hi | lo | hi]     -> [n][3*kpad]
 * so that one GEMM with k = 3*kpad yields hi.hi + hi.lo + lo.hi (fp32-class products).
 * *d_lo_flag (may be NULL) is OR-ed with 1 when some lo != 0, i.e. x is not bf16-representable.
 * ------------------------------------------------------------------------- */
/* Operand format of the GEMM entry points below (NULL = bf16 panels with row stride k, no output scale).
 * fp16 panels come from vsc_prepare_operand_f16: x is scaled by a power of two 2^e chosen from max|x| (device side, no
 * host round trip) so that hi = fp16(x 2^e) and lo = fp16((x 2^e - hi) 2^11) keep 22 significant bits of every value:
 *   side 0 (queries)     [hi 2^-11 | lo 2^-11 | hi]        side 1 (references)   [lo | hi | hi]
 * one GEMM over k = 3*kpad gives hi.lo + lo.hi + hi.hi with every product exact in the fp32 accumulator; what is left
 * is the accumulator itself, which TRUNCATES (measured: median |error| 7e-8 on unit-norm 512-d rows, up to ~1.5e-6 on
 * identical rows, where all products have one sign; an fp32 sgemm: 1e-8 / 6e-7).  The small cross products come first
 * for that reason.  k = kpad over the last third (pointer + 2*kpad elements, row stride lda / ldb = 3*kpad) is exact
 * whenever *d_lo_flag stays 0, i.e. every value has at most 11 significant bits.
 * *d_out_scale = 2^-(e_a + e_b) (the product of the two operands' *d_inv_scale) is multiplied into the accumulators. */
int vsc_prepare_operand_f16(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t kpad, int32_t side,
                            void *d_out_f16, float *d_inv_scale, int32_t *d_lo_flag, uint32_t *d_scratch,
                            vsc_stream_t stream);
/* More rows of an operand whose scale is fixed: *d_scratch / *d_inv_scale as an earlier vsc_prepare_operand_f16 call on
 * the same operand left them (no max|x| pass).  Lets a caller upload and convert a large descriptor collection piece by
 * piece while earlier pieces are already being multiplied (localization.py:56-79 at configs[3] size).  Bit 1 of
 * *d_lo_flag is set when a value does not fit the fp16 range under that scale: prepare the whole operand again. */
int vsc_prepare_operand_f16_more(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t kpad, int32_t side,
                                 void *d_out_f16, float *d_inv_scale, int32_t *d_lo_flag, uint32_t *d_scratch,
                                 vsc_stream_t stream);
/* The same for a list of rows: d_x / d_out_f16 address the WHOLE matrix / panel, d_rows[0..n) (device) are the absolute rows to
 * convert, in one launch; keep_scale != 0 behaves like _more, 0 like vsc_prepare_operand_f16 restricted to these rows. */
int vsc_prepare_operand_f16_rows(const float *d_x, const int32_t *d_rows, int64_t n, int32_t d, int64_t ld, int32_t kpad,
                                 int32_t side, void *d_out_f16, float *d_inv_scale, int32_t *d_lo_flag, uint32_t *d_scratch,
                                 int32_t keep_scale, vsc_stream_t stream);
int vsc_prepare_operand(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t kpad, int32_t mode,
                        void *d_out_bf16, int32_t *d_lo_flag, vsc_stream_t stream);
int vsc_row_sqnorm(const float *d_x, int64_t n, int32_t d, int64_t ld, float *d_out, vsc_stream_t stream);

/* *d_out = the k-th best of d_scores[0..n) (1-based; largest != 0: k-th largest, else k-th smallest) by radix selection:
 * the radius update of faiss.contrib.exhaustive_search.apply_maxres (vsc/index.py:147-154).  d_scratch: >= 8208 bytes. */
int vsc_kth_best(const float *d_scores, int64_t n, int64_t k, int32_t largest, float *d_out, void *d_scratch,
                 vsc_stream_t stream);

/* One radix-selection pass in two halves, for a k-th best over scores spread over several GPUs (query-sharded search):
 * vsc_select_hist counts this GPU's scores into the 2048-bin histogram at the start of d_state (8224 bytes: histogram,
 * then {prefix, mask, k}; pass 0 initialises it with k), the caller sums the histograms over the GPUs (all-reduce),
 * vsc_select_pick narrows the key prefix.  Passes 0, 1, 2; after the last pick *d_out is the k-th best. */
int vsc_select_hist(const float *d_scores, int64_t n, int64_t k, int32_t largest, int32_t pass, void *d_state,
                    vsc_stream_t stream);
int vsc_select_pick(int32_t largest, int32_t pass, void *d_state, float *d_out, vsc_stream_t stream);

/* The re-filter of apply_maxres: copy the entries of (score, row, col)[0..n) strictly beyond `radius` to the output
 * arrays (unspecified order, no overlap with the inputs); *d_count receives how many. */
int vsc_compact_hits(const float *d_score, const int32_t *d_row, const int32_t *d_col, int64_t n, float radius,
                     int32_t keep_max, float *d_score_out, int32_t *d_row_out, int32_t *d_col_out,
                     unsigned long long *d_count, vsc_stream_t stream);

/* The whole global-threshold search of VideoIndex._global_threshold_knn_search (vsc/index.py:142-165):
 *   faiss.contrib.exhaustive_search.range_search_max_results(index, exponential_query_iterator(xq), radius,
 *                                                            max_results, min_results)
 * enqueued without a host round trip: per exponential query batch (32, 64, ... rows) one vsc_gemm_emit-style launch whose
 * thresholds live in the device-side control block, followed by FAISS's bookkeeping there -- when the running total exceeds
 * max_results the (min_results+1)-th best held score becomes the radius and everything held is re-filtered strictly.
 * d_a: query panel (row stride a_row_bytes), d_b: reference panel.  Survivors end up in d_score / d_row / d_col[0 .. held)
 * (d_*2: scratch of the same capacity).  d_control: vsc_search_control_bytes() bytes; after the stream has drained it
 * holds {float radius; float; int32; int32 overflow; uint64 held; ...}.  overflow != 0: a batch emitted more than
 * `capacity` entries and the result is incomplete (repeat with more room or batch by batch with vsc_gemm_emit). */
int vsc_search_global_topk(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, const float *d_a_norm,
                           const float *d_b_norm, int32_t metric_l2, int64_t max_results, int64_t min_results,
                           float *d_score, int32_t *d_row, int32_t *d_col, float *d_score2, int32_t *d_row2,
                           int32_t *d_col2, uint64_t capacity, void *d_control, int32_t a_row_bytes,
                           const vsc_gemm_format *fmt, vsc_stream_t stream);
/* The same schedule with batches of at least filter_from_rows query rows filtered: one tensor-core product per value pair
 * (d_a_single / d_b_single: the hi columns of the panels, k_single) with the thresholds loosened by *d_margin (device scalar:
 * twice the single-product error bound 2^-10 |a| |b|), then the exact float32 inner products of the candidates from the
 * original matrices d_a_raw [m][lda_raw] / d_b_raw [n][ldb_raw] (d dimensions) against the real thresholds.  Inner product. */
int vsc_search_global_topk_filtered(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k,
                                    const void *d_a_single, const void *d_b_single, int32_t k_single,
                                    const float *d_a_raw, int64_t lda_raw, const float *d_b_raw, int64_t ldb_raw, int32_t d,
                                    const float *d_margin, int64_t filter_from_rows, int64_t max_results, int64_t min_results,
                                    float *d_score, int32_t *d_row, int32_t *d_col, float *d_score2, int32_t *d_row2,
                                    int32_t *d_col2, uint64_t capacity, void *d_control, int32_t a_row_bytes,
                                    const vsc_gemm_format *fmt, vsc_stream_t stream);
/* The schedule step by step, for the query-sharded search over several GPUs: every rank emits ITS rows of a batch
 * (vsc_search_emit_batch: three-product GEMM, or the filtered form), the ranks all-reduce the hit count, the radix histograms
 * and the survivor count between the phases of vsc_search_step (views of the control block at the offsets of
 * vsc_search_control_layout; NCCL on the same stream, no host round trip).  phase -1 begin (arg = metric_l2), 0 decide,
 * 1 / 2 histogram / pick of radix pass arg, 3 strict re-filter, 4 copy back + finish, 5 end (drop fillers). */
int vsc_search_step(int32_t phase, int32_t arg, void *d_control, float *d_score, int32_t *d_row, int32_t *d_col,
                    float *d_score2, int32_t *d_row2, int32_t *d_col2, uint64_t capacity, int64_t max_results,
                    int64_t min_results, int32_t keep_max, vsc_stream_t stream);
int vsc_search_emit_batch(const void *d_a, const void *d_b, int64_t n, int32_t k, const void *d_a_single,
                          const void *d_b_single, int32_t k_single, const float *d_a_raw, int64_t lda_raw,
                          const float *d_b_raw, int64_t ldb_raw, int32_t d, const float *d_margin, int32_t filtered,
                          const float *d_a_norm, const float *d_b_norm, int32_t metric_l2, int64_t at, int64_t rows,
                          float *d_score, int32_t *d_row, int32_t *d_col, float *d_score2, int32_t *d_row2, int32_t *d_col2,
                          uint64_t capacity, void *d_control, int32_t a_row_bytes, const vsc_gemm_format *fmt,
                          vsc_stream_t stream);
int vsc_search_control_layout(int32_t *out7);
int vsc_search_control_bytes(void);

/* Score normalisation around the row-max GEMM (vsc/baseline/score_normalization.py:68-104), one pass each:
 * vsc_lowvar_dim      *d_out = argmin over columns of the population variance of x[n][d] (:73-76); d_scratch: 16*d bytes.
 * vsc_l2norm_dropdim  out[i] = x[i] without column *d_drop (NULL: keep all), divided by its L2 norm when normalize != 0
 *                     (rows of norm 0 stay: sklearn.preprocessing.normalize) (:77-85); has_tail: out[i][kept columns] = tail
 *                     (the appended 1 of the references, :100-104).  ldo: row stride of out.
 * vsc_fill_column     out[i][col] = factor * src[i] (the appended -beta * 1-NN similarity of the queries, :97-99). */
int vsc_lowvar_dim(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t *d_out, void *d_scratch, vsc_stream_t stream);
int vsc_l2norm_dropdim(const float *d_x, int64_t n, int32_t d, int64_t ld, const int32_t *d_drop, int32_t normalize,
                       float *d_out, int64_t ldo, int32_t has_tail, float tail, vsc_stream_t stream);
int vsc_fill_column(float *d_out, int64_t n, int64_t ldo, int32_t col, const float *d_src, float factor, vsc_stream_t stream);
/* MaxScoreAggregation over the frame matches of every video pair (vsc/candidates.py:24-40): the n hits arrive sorted
 * best first (d_row / d_col: query / reference frame rows), d_q_vid / d_r_vid give the video of every row.  Writes the
 * first `limit` (query video, reference video, best score) triples in order of first appearance -- the reference's dict
 * order followed by its stable sort -- and the number of distinct pairs.  d_scratch: vsc_pair_max_scratch_bytes(n). */
int64_t vsc_pair_max_scratch_bytes(int64_t n);
int vsc_pair_max(const int64_t *d_row, const int64_t *d_col, const float *d_score, int64_t n, const int32_t *d_q_vid,
                 const int32_t *d_r_vid, int64_t n_ref_videos, int64_t limit, int32_t *d_out_q, int32_t *d_out_r,
                 float *d_out_score, unsigned long long *d_n_unique, void *d_scratch, vsc_stream_t stream);

/* C[m][n] = A . B^T in fp32 (tests, per-pair similarity matrices). */
int vsc_gemm_store(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, float *d_c, int64_t ldc,
                   const vsc_gemm_format *fmt, vsc_stream_t stream);
/* d_rowmax[i] = max_j A_i . B_j  -- FAISS index.search(x, 1) similarities
 * (score_normalization.py:93-96) without materialising the matrix. */
int vsc_gemm_rowmax(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, float *d_rowmax,
                    const vsc_gemm_format *fmt, vsc_stream_t stream);
/* FAISS index.search(x, 1) proper (index.py:169-174 with k = 1): best score and its column per row, lowest
 * column on exact ties.  d_scratch: m x 8 bytes of workspace. */
int vsc_gemm_rowargmax(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, float *d_score,
                       int64_t *d_col, unsigned long long *d_scratch, const vsc_gemm_format *fmt, vsc_stream_t stream);
/* FAISS range_search (index.py:147-154): counts the scores strictly beyond count_thr into
 * d_counters[1] and appends (score, row + row_offset, col + col_offset) of those strictly beyond
 * emit_thr into slots claimed from d_counters[0] (entries past `capacity` are dropped but still claimed).  Slots are
 * claimed per warp in blocks of 256; the unused tail of a warp's last block is filled with a score no threshold accepts
 * (-inf for inner product, +inf for L2), so d_counters[0] counts CLAIMED slots and a strict re-filter of the scores
 * removes the fillers.  At most 148 * 8 * 256 filler slots per call.
 * metric_l2 = 0: inner product, "beyond" = greater;  1: squared L2 = a_norm + b_norm - 2 a.b, "beyond" = smaller. */
int vsc_gemm_emit(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, const float *d_a_norm,
                  const float *d_b_norm, int32_t metric_l2, float count_thr, float emit_thr, int64_t row_offset,
                  int64_t col_offset, float *d_score, int32_t *d_row, int32_t *d_col, uint64_t capacity,
                  unsigned long long *d_counters, const vsc_gemm_format *fmt, vsc_stream_t stream);

/* -------------------------------------------------------------------------
 * Stage A: SSCD ResNet-50 frame descriptors (vsc/baseline/inference_impl.py:210-239 runs the TorchScript model;
 * architecture contract vsc/baseline/adapt_sscd_model.py:56-70).  Activations NHWC bf16; BN folded into the
 * convolution weights [cout][k] (bf16) and biases (fp32) by the host.
 * ------------------------------------------------------------------------- */
/* out[m][n] (bf16, row stride ldc) = relu?(A[m][:] . W[n][:] + bias[n] + residual[m][n]); n % 32 == 0 */
int vsc_gemm_conv(const void *d_a, int64_t m, const void *d_w, int64_t n, int32_t k, const float *d_bias,
                  const void *d_residual, int32_t relu, void *d_out_bf16, int64_t ldc, vsc_stream_t stream);
/* 3x3 / pad 1 / stride 1|2 convolution + folded-BN bias (+ residual) (+ ReLU) straight from the NHWC bf16 tensor
 * [n][h][w][c] (implicit GEMM: TMA im2col loads, no patch matrix).  c % 64 == 0, cout % 32 == 0; weights
 * [cout][9*c] with K index = (ky*3 + kx)*c + channel; out [n*ho*wo][cout] bf16. */
int vsc_conv3x3(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t stride, const void *d_w,
                int32_t cout, const float *d_bias, const void *d_residual, int32_t relu, void *d_out_bf16,
                vsc_stream_t stream);
/* 1x1 / stride 1|2 convolution (the ResNet downsample branch) through the same TMA im2col path: stride 2 reads every
 * second pixel of every second row without a subsampled copy.  Weights [cout][c]. */
int vsc_conv1x1(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t stride, const void *d_w,
                int32_t cout, const float *d_bias, const void *d_residual, int32_t relu, void *d_out_bf16,
                vsc_stream_t stream);
/* fp32 out[m][n] = A . W^T + bias[n] (projection head) */
int vsc_gemm_linear(const void *d_a, int64_t m, const void *d_w, int64_t n, int32_t k, const float *d_bias, float *d_out,
                    int64_t ldc, vsc_stream_t stream);
/* Stem: 7x7 / stride 2 / pad 3 convolution on 3 channels + folded-BN bias + ReLU.  mode 0: uint8 NHWC pixels
 * (normalised here), 1: float32 NCHW normalised.  Weights [64][256] with K index = ky2*64 + kx2*16 + (dy*2+dx)*3 + c
 * for filter tap (2*ky2+dy, 2*kx2+dx) (space-to-depth order; taps with row / column 7 and channels 12..15 of a
 * group are padding and must be zero).  Output bf16 [n][ho+3][wo+3][64]: rows < ho, columns < wo are the result. */
int vsc_conv_stem(const void *d_in, int32_t mode, int32_t n, int32_t h, int32_t w, const void *d_w, const float *d_bias,
                  void *d_out_bf16, vsc_stream_t stream);
/* the GEMM half of vsc_conv_stem on a prepared space-to-depth image (pixels cells of 16 bf16, +4 cells readable) */
int vsc_gemm_stem(const void *d_s2d, int64_t pixels, int64_t row_shift, const void *d_w, const float *d_bias,
                  void *d_out_bf16, vsc_stream_t stream);
/* 3x3 pad-1 patches [n*ho*wo][9*c] of an NHWC bf16 tensor, stride 1 or 2 */
int vsc_im2col3x3(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t stride, void *d_out,
                  vsc_stream_t stream);
int vsc_subsample2(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, void *d_out, vsc_stream_t stream);
int vsc_maxpool3x3s2(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t row_pitch, int32_t img_rows,
                     void *d_out, vsc_stream_t stream);
/* GeM pooling over hw pixels: (mean clamp(x, eps)^p)^(1/p), bf16 [n][c] out */
int vsc_gem_pool(const void *d_in, int32_t n, int32_t hw, int32_t c, float p, float eps, void *d_out, vsc_stream_t stream);

/* Host -> device copies of n_ranges row ranges ([first_row, n_rows] pairs, int64) of one array, same offsets on both sides,
 * one cudaMemcpyAsync each, in one call: the block-wise descriptor upload of localize_all (localization.py:56-79 keeps all
 * descriptors in host dicts; here only the rows a batch of candidates needs cross PCIe, once). */
int vsc_upload_rows(void *d_dst, const void *h_src, int64_t row_bytes, const int64_t *ranges, int32_t n_ranges,
                    vsc_stream_t stream);

/* Filtered row maximum (FAISS index.search(x, 1) of score_normalization.py:93-96 at one tensor-core product per value pair
 * instead of three).  vsc_gemm_emit_rows: inner-product range search with one threshold per query row -- appends every
 * (score, i, j) with score > d_row_thr[i] (counters as vsc_gemm_emit).  vsc_rowmax_rescore: float32 inner products of the
 * candidate (row, column) pairs from the ORIGINAL float32 matrices (fused multiply-adds, fixed order) and their maximum
 * per row into d_out[m] (-inf for a row without candidate); d_scratch_keys: m uint32.  With thresholds = single-product
 * row maxima minus twice the single-product error bound the candidates contain the true arg max of every row. */
int vsc_gemm_emit_rows(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, const float *d_row_thr,
                       float *d_score, int32_t *d_row, int32_t *d_col, uint64_t capacity, unsigned long long *d_counters,
                       const vsc_gemm_format *fmt, vsc_stream_t stream);
int vsc_rowmax_rescore(const float *d_a, int64_t m, int64_t lda, const float *d_b, int64_t n, int64_t ldb, int32_t d,
                       const float *d_cand_score, const int32_t *d_cand_row, const int32_t *d_cand_col, int64_t n_cand,
                       uint32_t *d_scratch_keys, float *d_out, vsc_stream_t stream);

/* Number of kernel launches issued by this library since load (all entry points). */
int64_t vsc_launch_count(void);

/* -------------------------------------------------------------------------
 * Stage A, ahead of the model: the PIL bilinear resize of torchvision.transforms.Resize (+ CenterCrop) in the
 * reference's transforms, vsc/baseline/inference_impl.py:39-69, on decoded uint8 RGB frames [n][h][w][3].
 * The frame is resized to rh x rw and the window [top, top+oh) x [left, left+ow) of it is written to d_out
 * [n][oh][ow][3].  d_xbounds / d_xk ([rw][2] / [rw][xksize]) and d_ybounds / d_yk ([rh][2] / [rh][yksize]) are Pillow's
 * (first input pixel, count) windows and 22-bit fixed-point weights (libImaging/Resample.c precompute_coeffs +
 * normalize_coeffs_8bpc; vsc2022_b200/preprocess.py computes them); weights behind a window's count must be ZERO, as Pillow
 * leaves them (the horizontal pass runs every window over all xksize taps).  d_tmp: n*h*ow*3 bytes of scratch.
 * The result equals PIL's Image.resize(..., BILINEAR) bit for bit.
 * ------------------------------------------------------------------------- */
int vsc_resize_u8(const uint8_t *d_in, int32_t n, int32_t h, int32_t w, int32_t rh, int32_t rw,
                  int32_t top, int32_t left, int32_t oh, int32_t ow,
                  const int32_t *d_xbounds, const int32_t *d_xk, int32_t xksize,
                  const int32_t *d_ybounds, const int32_t *d_yk, int32_t yksize,
                  uint8_t *d_tmp, uint8_t *d_out, vsc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VSC_B200_H */
