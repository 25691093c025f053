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

/* Per-stage device timing of the TN fast pipeline (CUDA events recorded on the caller's stream
 * around each stage).  vsc_tn_last_stage_ms fills {row top-K, edges, sweeps, MaxSim} in ms for the
 * most recent vcsl_tn_batch call; synchronise the stream first. */
int vsc_tn_set_profiling(int on);
int vsc_tn_last_stage_ms(float *out4);

/* Development aid: cumulative DP work counters {first-sweep layers, incremental layer steps,
 * chains found, chain nodes}; all zero unless the library was built with -DVSC_TN_COUNTERS. */
int vsc_tn_debug_counters(unsigned long long *out4);

/* Number of kernel launches issued by this library since load (all entry points). */
int64_t vsc_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* VSC_B200_H */
