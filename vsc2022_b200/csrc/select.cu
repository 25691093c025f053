// k-th best of a score array by radix selection (sm_100a).
//
// Replaces the partition step of faiss.contrib.exhaustive_search.apply_maxres behind vsc/index.py:147-154
// (`alldis.partition(len - target - 1); radius = alldis[-1 - target]`): the new radius is the k-th largest (inner
// product) or k-th smallest (L2) of the scores held so far, k = min_results + 1.  A sort-based top-k over the 3-6 M held
// scores cost 2-3 ms per tightening and there are five or six tightenings per 40k x 200k search; three histogram
// passes (11 + 11 + 10 bits of the order-preserving key) read the array three times instead.  HBM-bound.
#include "common.cuh"
#include "search_internal.cuh"

namespace {

constexpr int kBins = 2048;

using vsc::SearchControl;
using vsc::SelectState;

__device__ __forceinline__ uint32_t order_key(float v, int largest) {
    const uint32_t key = vsc::float_to_key(v);     // ascending with the value
    return largest ? key : ~key;                   // "best" is always the largest key
}

// histogram of `bits` key bits starting at `shift` over the elements that match the decided prefix
// `ctl` (may be null): device-driven search -- the element count is ctl->counters[0] and the kernel does nothing unless
// ctl->do_tighten is set.
__global__ void __launch_bounds__(512) select_hist_kernel(const float *__restrict__ x, int64_t n, int largest, int shift,
                                                          int bits, const SelectState *__restrict__ st,
                                                          unsigned int *__restrict__ hist,
                                                          const SearchControl *__restrict__ ctl) {
    __shared__ unsigned int h[kBins];
    if (ctl) {
        if (!ctl->do_tighten || ctl->overflow) return;
        n = (int64_t)ctl->counters[0];
    }
    for (int i = threadIdx.x; i < kBins; i += blockDim.x) h[i] = 0;
    __syncthreads();
    const uint32_t prefix = st->prefix, mask = st->mask, field = (1u << bits) - 1u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t key = order_key(x[i], largest);
        if ((key & mask) == prefix) atomicAdd(&h[(key >> shift) & field], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBins; i += blockDim.x)
        if (h[i]) atomicAdd(&hist[i], h[i]);
}

// one block of 256 threads: find the bin (counting from the best = highest key downwards) that holds the remaining
// rank.  Thread t owns bins [8t, 8t+8); an inclusive suffix sum over the threads' totals locates the owner.
__global__ void __launch_bounds__(256) select_pick_kernel(SelectState *st, unsigned int *hist, int shift, int bits,
                                                          int largest, int last, float *out,
                                                          const SearchControl *__restrict__ ctl) {
    __shared__ unsigned long long suffix[256];
    if (ctl && !ctl->do_tighten) return;
    const int t = threadIdx.x;
    unsigned int c[8];
    unsigned long long mine = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = hist[t * 8 + i]; mine += c[i]; hist[t * 8 + i] = 0; }   // leaves the histogram clear
    suffix[t] = mine;
    __syncthreads();
    for (int d = 1; d < 256; d <<= 1) {   // suffix[t] = sum of the totals of threads >= t
        const unsigned long long add = t + d < 256 ? suffix[t + d] : 0;
        __syncthreads();
        suffix[t] += add;
        __syncthreads();
    }
    const unsigned long long k = st->k;
    const unsigned long long above = t + 1 < 256 ? suffix[t + 1] : 0;   // elements in better bins than mine
    __syncthreads();
    // exactly one thread has above < k <= above + mine (k <= total matching elements by construction)
    if (above < k && k <= above + mine) {
        unsigned long long r = k - above;
        int b = 7;
        for (; b > 0; --b) {
            if (r <= c[b]) break;
            r -= c[b];
        }
        const uint32_t prefix = st->prefix | ((uint32_t)(t * 8 + b) << shift);
        st->prefix = prefix;
        st->mask |= ((1u << bits) - 1u) << shift;
        st->k = r;
        if (last) *out = vsc::key_to_float(largest ? prefix : ~prefix);
    }
}

// Unordered stream compaction of the survivors: keep entries strictly beyond `radius`.  One atomic per 256 entries.
__global__ void __launch_bounds__(256) compact_kernel(const float *__restrict__ s_in, const int32_t *__restrict__ r_in,
                                                      const int32_t *__restrict__ c_in, int64_t n, float radius,
                                                      int keep_max, float *__restrict__ s_out, int32_t *__restrict__ r_out,
                                                      int32_t *__restrict__ c_out, unsigned long long *__restrict__ count,
                                                      const SearchControl *__restrict__ ctl, int always) {
    __shared__ unsigned int warp_total[8];
    __shared__ unsigned long long block_base;
    if (ctl) {   // device-driven search: count and radius live in the control block
        if ((!always && !ctl->do_tighten) || ctl->overflow) return;   // after an overflow the count exceeds the buffer
        n = (int64_t)ctl->counters[0];
        radius = ctl->thr[0];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t i0 = (int64_t)blockIdx.x * 256; i0 < n; i0 += (int64_t)gridDim.x * 256) {
        const int64_t i = i0 + threadIdx.x;
        float sc = 0.0f;
        bool keep = false;
        if (i < n) {
            sc = s_in[i];
            keep = keep_max ? sc > radius : sc < radius;
        }
        const unsigned int vote = __ballot_sync(vsc::kFullMask, keep);
        if (lane == 0) warp_total[warp] = __popc(vote);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned int t = 0;
            for (int w = 0; w < 8; ++w) t += warp_total[w];
            block_base = t ? atomicAdd(count, (unsigned long long)t) : 0ull;
        }
        __syncthreads();
        if (keep) {
            unsigned long long at = block_base + __popc(vote & ((1u << lane) - 1u));
            for (int w = 0; w < warp; ++w) at += warp_total[w];
            s_out[at] = sc; r_out[at] = r_in[i]; c_out[at] = c_in[i];
        }
        __syncthreads();   // warp_total / block_base are rewritten by the next round
    }
}


// ---- device-driven FAISS schedule (search.cu): small control kernels
__global__ void search_decide_kernel(SearchControl *c, unsigned long long capacity, unsigned long long max_results,
                                     unsigned long long min_results) {
    c->total += c->counters[1];
    c->counters[1] = 0;
    c->kept = 0;
    if (c->counters[0] > capacity) c->overflow = 1;   // entries were dropped: the host repeats the search its own way
    c->do_tighten = !c->overflow && c->total > max_results;
    if (c->do_tighten) { c->sel.prefix = 0; c->sel.mask = 0; c->sel.k = min_results + 1; }
}
__global__ void __launch_bounds__(256) search_copy_back_kernel(const SearchControl *__restrict__ c, const float *__restrict__ s2,
                                                               const int32_t *__restrict__ r2, const int32_t *__restrict__ c2,
                                                               float *__restrict__ s, int32_t *__restrict__ r,
                                                               int32_t *__restrict__ cc, int always) {
    if ((!always && !c->do_tighten) || c->overflow) return;
    const int64_t n = (int64_t)c->kept;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        s[i] = s2[i]; r[i] = r2[i]; cc[i] = c2[i];
    }
}
__global__ void search_finish_kernel(SearchControl *c, int always) {
    if ((!always && !c->do_tighten) || c->overflow) return;
    c->counters[0] = c->kept;
    if (!always) { c->total = c->use_global ? c->kept_global : c->kept; c->thr[1] = c->thr[0]; c->n_tighten += 1; }
    c->do_tighten = 0;
}

}  // namespace

namespace vsc {

// Filtered batch: the candidates (approximate score beyond threshold - margin) get their float32 inner product from the
// original matrices -- one warp per candidate, lane l over k = l, l+32, ... with fused multiply-adds in ascending k, then a fixed
// xor-shuffle tree (the arithmetic of vsc_rowmax_rescore: a pair's score does not depend on its batch) -- and those beyond the
// thresholds of the control block are counted / appended to the survivor buffer exactly as the emit epilogue would have:
// strict comparisons, one slot claim per 64 candidates.
namespace {
constexpr int kRescoreGroup = 64;
__global__ void __launch_bounds__(256) rescore_append_kernel(SearchControl *ctl, const float *__restrict__ a, int64_t lda,
                                                             const float *__restrict__ b, int64_t ldb, int d,
                                                             const float *__restrict__ cand_s, const int32_t *__restrict__ cand_r,
                                                             const int32_t *__restrict__ cand_c, unsigned long long capacity,
                                                             float *__restrict__ s_out, int32_t *__restrict__ r_out,
                                                             int32_t *__restrict__ c_out) {
    __shared__ float sc[kRescoreGroup];
    __shared__ int32_t rw[kRescoreGroup], cl[kRescoreGroup];
    __shared__ unsigned long long base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long claimed = ctl->cand_counters[0];
    if (claimed > capacity) { if (blockIdx.x == 0 && threadIdx.x == 0) ctl->overflow = 1; return; }
    const float count_thr = ctl->thr[0], emit_thr = ctl->thr[1];
    for (unsigned long long g0 = (unsigned long long)blockIdx.x * kRescoreGroup; g0 < claimed; g0 += (unsigned long long)gridDim.x * kRescoreGroup) {
        for (int t = warp; t < kRescoreGroup; t += 8) {
            const unsigned long long ci = g0 + t;
            float v = -INFINITY;
            int32_t i = -1, j = -1;
            if (ci < claimed && cand_s[ci] > -INFINITY) {     // fillers of the emit epilogue's per-warp blocks carry -inf
                i = cand_r[ci]; j = cand_c[ci];
                const float *x = a + (int64_t)i * lda, *y = b + (int64_t)j * ldb;
                float acc = 0.0f;
                for (int k = lane; k < d; k += 32) acc = __fmaf_rn(x[k], y[k], acc);
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_xor_sync(vsc::kFullMask, acc, sft);
                v = acc;
            }
            if (lane == 0) { sc[t] = v; rw[t] = i; cl[t] = j; }
        }
        __syncthreads();
        if (warp == 0) {
            const float v0 = sc[lane], v1 = sc[lane + 32];
            const bool e0 = v0 > emit_thr, e1 = v1 > emit_thr;
            const unsigned m0 = __ballot_sync(vsc::kFullMask, e0), m1 = __ballot_sync(vsc::kFullMask, e1);
            const int n_emit = __popc(m0) + __popc(m1);
            const int n_count = __popc(__ballot_sync(vsc::kFullMask, v0 > count_thr)) + __popc(__ballot_sync(vsc::kFullMask, v1 > count_thr));
            if (lane == 0) {
                base = n_emit ? atomicAdd(&ctl->counters[0], (unsigned long long)n_emit) : 0ull;
                if (n_count) atomicAdd(&ctl->counters[1], (unsigned long long)n_count);
            }
            __syncwarp();
            const unsigned long long b0 = base;
            if (e0) {
                const unsigned long long at = b0 + __popc(m0 & ((1u << lane) - 1u));
                if (at < capacity) { s_out[at] = v0; r_out[at] = rw[lane]; c_out[at] = cl[lane]; }
            }
            if (e1) {
                const unsigned long long at = b0 + __popc(m0) + __popc(m1 & ((1u << lane) - 1u));
                if (at < capacity) { s_out[at] = v1; r_out[at] = rw[lane + 32]; c_out[at] = cl[lane + 32]; }
            }
        }
        __syncthreads();
    }
}
}  // namespace

int search_rescore_append(SearchControl *ctl, const float *d_a_raw, int64_t lda, const float *d_b_raw, int64_t ldb, int32_t d,
                          const float *cand_s, const int32_t *cand_r, const int32_t *cand_c, uint64_t capacity, float *s,
                          int32_t *r, int32_t *c, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    rescore_append_kernel<<<sms * 8, 256, 0, stream>>>(ctl, d_a_raw, lda, d_b_raw, ldb, d, cand_s, cand_r, cand_c,
                                                        (unsigned long long)capacity, s, r, c);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

// After one range-search launch of the device-driven schedule: FAISS's bookkeeping, and -- when the running total
// exceeds max_results -- the new radius ((min_results+1)-th best held score, radix selection) and the strict re-filter.
// One phase of the bookkeeping after a batch (the query-sharded search interleaves them with all-reduces, vsc_search_step):
// 0 decide, 1 histogram of radix pass `arg`, 2 pick of pass `arg`, 3 strict re-filter into the twin buffer, 4 copy back + finish.
int search_phase(int phase, int arg, SearchControl *ctl, float *s, int32_t *r, int32_t *c, float *s2, int32_t *r2, int32_t *c2,
                 uint64_t capacity, int64_t max_results, int64_t min_results, int keep_max, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    switch (phase) {
        case 0:
            search_decide_kernel<<<1, 1, 0, stream>>>(ctl, capacity, (unsigned long long)max_results, (unsigned long long)min_results);
            vsc::count_launch(1);
            break;
        case 1:
            select_hist_kernel<<<sms * 4, 512, 0, stream>>>(s, 0, keep_max, shifts[arg], bits[arg], &ctl->sel, ctl->hist, ctl);
            vsc::count_launch(1);
            break;
        case 2:
            select_pick_kernel<<<1, 256, 0, stream>>>(&ctl->sel, ctl->hist, shifts[arg], bits[arg], keep_max, arg == 2, &ctl->thr[0], ctl);
            vsc::count_launch(1);
            break;
        case 3:
            compact_kernel<<<sms * 8, 256, 0, stream>>>(s, r, c, 0, 0.0f, keep_max, s2, r2, c2, &ctl->kept, ctl, 0);
            vsc::count_launch(1);
            break;
        case 4:
            search_copy_back_kernel<<<sms * 4, 256, 0, stream>>>(ctl, s2, r2, c2, s, r, c, 0);
            search_finish_kernel<<<1, 1, 0, stream>>>(ctl, 0);
            vsc::count_launch(2);
            break;
        default:
            vsc::set_error("search_phase: unknown phase %d", phase);
            return VSC_ERR_INVALID;
    }
    VSC_CUDA_CHECK(cudaGetLastError());
    return VSC_OK;
}

int search_after_batch(SearchControl *ctl, float *s, int32_t *r, int32_t *c, float *s2, int32_t *r2, int32_t *c2,
                       uint64_t capacity, int64_t max_results, int64_t min_results, int keep_max, cudaStream_t stream) {
    int rc = search_phase(0, 0, ctl, s, r, c, s2, r2, c2, capacity, max_results, min_results, keep_max, stream);
    for (int p = 0; p < 3 && rc == VSC_OK; ++p) {
        rc = search_phase(1, p, ctl, s, r, c, s2, r2, c2, capacity, max_results, min_results, keep_max, stream);
        if (rc == VSC_OK) rc = search_phase(2, p, ctl, s, r, c, s2, r2, c2, capacity, max_results, min_results, keep_max, stream);
    }
    if (rc == VSC_OK) rc = search_phase(3, 0, ctl, s, r, c, s2, r2, c2, capacity, max_results, min_results, keep_max, stream);
    if (rc == VSC_OK) rc = search_phase(4, 0, ctl, s, r, c, s2, r2, c2, capacity, max_results, min_results, keep_max, stream);
    return rc;
}

// End of the schedule: drop the never-accepted fillers of the emit epilogue (strict re-filter with the final radius).
int search_final_filter(SearchControl *ctl, float *s, int32_t *r, int32_t *c, float *s2, int32_t *r2, int32_t *c2,
                        int keep_max, cudaStream_t stream) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    VSC_CUDA_CHECK(cudaMemsetAsync(&ctl->kept, 0, sizeof(unsigned long long), stream));
    compact_kernel<<<sms * 8, 256, 0, stream>>>(s, r, c, 0, 0.0f, keep_max, s2, r2, c2, &ctl->kept, ctl, 1);
    search_copy_back_kernel<<<sms * 4, 256, 0, stream>>>(ctl, s2, r2, c2, s, r, c, 1);
    search_finish_kernel<<<1, 1, 0, stream>>>(ctl, 1);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch(3);
    return VSC_OK;
}

}  // namespace vsc

// Copies the entries of (score, row, col)[0..n) whose score is strictly beyond `radius` (greater for keep_max != 0,
// smaller otherwise) to the output arrays in unspecified order; *d_count (zeroed here) receives how many.  The
// re-filter of faiss.contrib.exhaustive_search.apply_maxres.  In and out must not overlap.
extern "C" int vsc_compact_hits(const float *d_score, const int32_t *d_row, const int32_t *d_col, int64_t n, float radius,
                                int32_t keep_max, float *d_score_out, int32_t *d_row_out, int32_t *d_col_out,
                                unsigned long long *d_count, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    VSC_CUDA_CHECK(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), stream));
    if (n <= 0) return VSC_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + 255) / 256;
    const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
    compact_kernel<<<grid, 256, 0, stream>>>(d_score, d_row, d_col, n, radius, keep_max, d_score_out, d_row_out,
                                             d_col_out, d_count, nullptr, 0);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

// *d_out = the k-th best (1-based; largest != 0: k-th largest, else k-th smallest) of d_scores[0..n).  NaNs are not
// supported.  d_scratch: at least 8208 bytes of device memory (histogram + state), contents irrelevant.
extern "C" int vsc_kth_best(const float *d_scores, int64_t n, int64_t k, int32_t largest, float *d_out, void *d_scratch,
                            vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0 || k < 1 || k > n || !d_scores || !d_out || !d_scratch) {
        vsc::set_error("vsc_kth_best: need 1 <= k <= n and non-null buffers (n=%lld, k=%lld)", (long long)n, (long long)k);
        return VSC_ERR_INVALID;
    }
    unsigned int *hist = static_cast<unsigned int *>(d_scratch);
    SelectState *st = reinterpret_cast<SelectState *>(hist + kBins);
    VSC_CUDA_CHECK(cudaMemsetAsync(d_scratch, 0, sizeof(unsigned int) * kBins + sizeof(SelectState), stream));
    const SelectState init = {0u, 0u, (unsigned long long)k};
    VSC_CUDA_CHECK(cudaMemcpyAsync(st, &init, sizeof init, cudaMemcpyHostToDevice, stream));
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + 511) / 512;
    const int grid = (int)(want < (int64_t)sms * 4 ? want : (int64_t)sms * 4);
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    for (int p = 0; p < 3; ++p) {
        select_hist_kernel<<<grid, 512, 0, stream>>>(d_scores, n, largest, shifts[p], bits[p], st, hist, nullptr);
        select_pick_kernel<<<1, 256, 0, stream>>>(st, hist, shifts[p], bits[p], largest, p == 2, d_out, nullptr);
    }
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch(6);
    return VSC_OK;
}

// The two halves of one radix-selection pass, for selections whose histogram is summed over several GPUs between them
// (query-sharded search: every rank counts its own held scores, the histograms are all-reduced, every rank picks the
// same bin).  d_state: 8208 + 16 bytes = 2048 histogram bins followed by {prefix, mask, k}; pass 0 initialises it with k.
// pass in 0..2; after the pick of pass 2, *d_out holds the k-th best score.
extern "C" int vsc_select_hist(const float *d_scores, int64_t n, int64_t k, int32_t largest, int32_t pass, void *d_state,
                               vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (pass < 0 || pass > 2 || !d_state || n < 0) { vsc::set_error("vsc_select_hist: bad arguments"); return VSC_ERR_INVALID; }
    unsigned int *hist = static_cast<unsigned int *>(d_state);
    SelectState *st = reinterpret_cast<SelectState *>(hist + kBins);
    if (pass == 0) {
        VSC_CUDA_CHECK(cudaMemsetAsync(d_state, 0, sizeof(unsigned int) * kBins + sizeof(SelectState), stream));
        const SelectState init = {0u, 0u, (unsigned long long)k};
        VSC_CUDA_CHECK(cudaMemcpyAsync(st, &init, sizeof init, cudaMemcpyHostToDevice, stream));
    }
    if (n == 0) return VSC_OK;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = (n + 511) / 512;
    const int grid = (int)(want < (int64_t)sms * 4 ? want : (int64_t)sms * 4);
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    select_hist_kernel<<<grid, 512, 0, stream>>>(d_scores, n, largest, shifts[pass], bits[pass], st, hist, nullptr);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

extern "C" int vsc_select_pick(int32_t largest, int32_t pass, void *d_state, float *d_out, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (pass < 0 || pass > 2 || !d_state || !d_out) { vsc::set_error("vsc_select_pick: bad arguments"); return VSC_ERR_INVALID; }
    unsigned int *hist = static_cast<unsigned int *>(d_state);
    SelectState *st = reinterpret_cast<SelectState *>(hist + kBins);
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    select_pick_kernel<<<1, 256, 0, stream>>>(st, hist, shifts[pass], bits[pass], largest, pass == 2, d_out, nullptr);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
