// Light kernels around the descriptor search (stage B), all one pass over their data, HBM-bound:
//   vsc_lowvar_dim        lowest-variance column of the noise set          score_normalization.py:73-76 (np.var(...).argmin())
//   vsc_l2norm_dropdim    delete that column, L2-normalise the rows        score_normalization.py:77-85 (np.delete, sklearn normalize)
//   vsc_fill_column       out[i][col] = factor * src[i]                     score_normalization.py:97-99 (append -beta * 1-NN similarity)
//   vsc_pair_max          best frame score per (query video, ref video)     candidates.py:24-40 (MaxScoreAggregation over PairMatches)
#include "common.cuh"

namespace {

using vsc::kFullMask;

// ---- column moments in float64: sum and sum of squares per column, one block per 256-row slab, atomics per column
__global__ void __launch_bounds__(256) column_moments_kernel(const float *__restrict__ x, int64_t n, int d, int64_t ld,
                                                             double *__restrict__ sum, double *__restrict__ sumsq) {
    const int64_t r0 = (int64_t)blockIdx.x * 256;
    const int64_t r1 = r0 + 256 < n ? r0 + 256 : n;
    for (int c = threadIdx.x; c < d; c += blockDim.x) {   // a warp reads 32 consecutive columns of a row: coalesced
        double s = 0.0, q = 0.0;
        for (int64_t r = r0; r < r1; ++r) {
            const double v = (double)x[r * ld + c];
            s += v; q += v * v;
        }
        atomicAdd(&sum[c], s);
        atomicAdd(&sumsq[c], q);
    }
}
__global__ void lowvar_pick_kernel(const double *__restrict__ sum, const double *__restrict__ sumsq, int64_t n, int d,
                                   int32_t *__restrict__ out) {
    // one warp: population variance per column, first minimum (numpy argmin)
    const int lane = threadIdx.x;
    double best = INFINITY; int arg = 0x7fffffff;
    for (int c = lane; c < d; c += 32) {
        const double m = sum[c] / (double)n;
        const double v = sumsq[c] / (double)n - m * m;
        if (v < best) { best = v; arg = c; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double ob = __shfl_xor_sync(kFullMask, best, s);
        const int oa = __shfl_xor_sync(kFullMask, arg, s);
        if (ob < best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) *out = arg;
}

// ---- one warp per row: out[row][0 .. d-2] = x[row] without column *drop, divided by its L2 norm (rows of norm 0 are
// left alone, like sklearn.preprocessing.normalize); out[row][d-1 ..] untouched unless `tail` says otherwise
__global__ void __launch_bounds__(256) l2norm_dropdim_kernel(const float *__restrict__ x, int64_t n, int d, int64_t ld,
                                                             const int32_t *__restrict__ drop_ptr, int normalize,
                                                             float *__restrict__ out, int64_t ldo, int has_tail,
                                                             float tail) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const int drop = drop_ptr ? *drop_ptr : -1;
    const float *src = x + row * ld;
    float acc = 0.0f;
    for (int k = lane; k < d; k += 32) {
        const float v = k == drop ? 0.0f : src[k];
        acc = fmaf(v, v, acc);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(kFullMask, acc, s);
    float norm = sqrtf(acc);
    if (!normalize || norm == 0.0f) norm = 1.0f;
    float *dst = out + row * ldo;
    for (int k = lane; k < d; k += 32) {
        if (k == drop) continue;
        dst[k - (drop >= 0 && k > drop ? 1 : 0)] = src[k] / norm;
    }
    if (has_tail && lane == 0) dst[d - (drop >= 0 ? 1 : 0)] = tail;
}

__global__ void __launch_bounds__(256) fill_column_kernel(float *__restrict__ out, int64_t n, int64_t ldo, int col,
                                                          const float *__restrict__ src, float factor) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i * ldo + col] = factor * src[i];
}

// ---- pair max: hits arrive sorted best first, so a video pair's FIRST hit carries its maximum, and the order of first
// appearances is the reference's order (dict insertion order, then a stable sort by score).
constexpr unsigned long long kEmptyKey = ~0ull;
struct Slot { unsigned long long key; unsigned int first; unsigned int pad; };

__global__ void __launch_bounds__(256) pairmax_clear_kernel(Slot *__restrict__ table, int64_t slots) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < slots) { table[i].key = kEmptyKey; table[i].first = 0xFFFFFFFFu; }
}
__device__ __forceinline__ unsigned long long mix(unsigned long long k) {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return k;
}
__device__ __forceinline__ int64_t probe(Slot *table, int64_t mask, unsigned long long key, bool insert) {
    int64_t at = (int64_t)(mix(key) & (unsigned long long)mask);
    for (;;) {
        unsigned long long cur = table[at].key;
        if (cur == key) return at;
        if (cur == kEmptyKey) {
            if (!insert) return -1;
            cur = atomicCAS(&table[at].key, kEmptyKey, key);
            if (cur == kEmptyKey || cur == key) return at;
        }
        at = (at + 1) & mask;
    }
}
__global__ void __launch_bounds__(256) pairmax_insert_kernel(const int64_t *__restrict__ row, const int64_t *__restrict__ col,
                                                             const int32_t *__restrict__ q_vid, const int32_t *__restrict__ r_vid,
                                                             int64_t n, int64_t n_ref_videos, Slot *__restrict__ table,
                                                             int64_t mask) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long key = (unsigned long long)q_vid[row[i]] * (unsigned long long)n_ref_videos + (unsigned long long)r_vid[col[i]];
    atomicMin(&table[probe(table, mask, key, true)].first, (unsigned int)i);
}
// flags + per-block counts (1024 hits per block), then the scan of the block counts, then the ordered scatter
__global__ void __launch_bounds__(256) pairmax_count_kernel(const int64_t *__restrict__ row, const int64_t *__restrict__ col,
                                                            const int32_t *__restrict__ q_vid, const int32_t *__restrict__ r_vid,
                                                            int64_t n, int64_t n_ref_videos, Slot *__restrict__ table,
                                                            int64_t mask, unsigned int *__restrict__ block_count) {
    __shared__ unsigned int total;
    if (threadIdx.x == 0) total = 0;
    __syncthreads();
    unsigned int mine = 0;
    for (int j = 0; j < 4; ++j) {
        const int64_t i = (int64_t)blockIdx.x * 1024 + j * 256 + threadIdx.x;
        if (i < n) {
            const unsigned long long key = (unsigned long long)q_vid[row[i]] * (unsigned long long)n_ref_videos + (unsigned long long)r_vid[col[i]];
            mine += table[probe(table, mask, key, false)].first == (unsigned int)i;
        }
    }
    mine = __reduce_add_sync(kFullMask, mine);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&total, mine);
    __syncthreads();
    if (threadIdx.x == 0) block_count[blockIdx.x] = total;
}
__global__ void __launch_bounds__(1024) pairmax_scan_kernel(unsigned int *__restrict__ block_count, int64_t n_blocks,
                                                            unsigned long long *__restrict__ n_unique) {
    // one block; sequential over chunks of 1024 block counts (a few thousand blocks at most)
    __shared__ unsigned int warp_sum[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t base = 0; base < n_blocks; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const unsigned int v = i < n_blocks ? block_count[i] : 0;
        unsigned int incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int up = __shfl_up_sync(kFullMask, incl, d);
            if ((threadIdx.x & 31) >= d) incl += up;
        }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned int before = carry;
        for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) before += warp_sum[w];
        if (i < n_blocks) block_count[i] = before + incl - v;   // exclusive prefix
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_unique = carry;
}
__global__ void __launch_bounds__(256) pairmax_scatter_kernel(const int64_t *__restrict__ row, const int64_t *__restrict__ col,
                                                              const int32_t *__restrict__ q_vid, const int32_t *__restrict__ r_vid,
                                                              const float *__restrict__ score, int64_t n, int64_t n_ref_videos,
                                                              Slot *__restrict__ table, int64_t mask,
                                                              const unsigned int *__restrict__ block_base, int64_t limit,
                                                              int32_t *__restrict__ out_q, int32_t *__restrict__ out_r,
                                                              float *__restrict__ out_s) {
    // ordered: thread t of the block handles hits [4t, 4t+4) of the block's 1024, so a block-wide exclusive scan of the
    // per-thread flag counts gives positions in hit order
    __shared__ unsigned int warp_sum[8];
    const int64_t i0 = (int64_t)blockIdx.x * 1024 + threadIdx.x * 4;
    int qv[4], rv[4]; bool flag[4];
    unsigned int mine = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int64_t i = i0 + j;
        flag[j] = false; qv[j] = rv[j] = 0;
        if (i < n) {
            qv[j] = q_vid[row[i]]; rv[j] = r_vid[col[i]];
            const unsigned long long key = (unsigned long long)qv[j] * (unsigned long long)n_ref_videos + (unsigned long long)rv[j];
            flag[j] = table[probe(table, mask, key, false)].first == (unsigned int)i;
            mine += flag[j];
        }
    }
    unsigned int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned int up = __shfl_up_sync(kFullMask, incl, d);
        if ((threadIdx.x & 31) >= d) incl += up;
    }
    if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned int at = block_base[blockIdx.x] + incl - mine;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) at += warp_sum[w];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!flag[j]) continue;
        if ((int64_t)at < limit) { out_q[at] = qv[j]; out_r[at] = rv[j]; out_s[at] = score[i0 + j]; }
        ++at;
    }
}

}  // namespace

extern "C" int vsc_lowvar_dim(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t *d_out, void *d_scratch,
                              vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0 || d <= 0 || !d_x || !d_out || !d_scratch) { vsc::set_error("vsc_lowvar_dim: bad arguments"); return VSC_ERR_INVALID; }
    double *sum = static_cast<double *>(d_scratch), *sumsq = sum + d;
    VSC_CUDA_CHECK(cudaMemsetAsync(d_scratch, 0, sizeof(double) * 2 * d, stream));
    column_moments_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_x, n, d, ld, sum, sumsq);
    lowvar_pick_kernel<<<1, 32, 0, stream>>>(sum, sumsq, n, d, d_out);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch(2);
    return VSC_OK;
}

extern "C" int vsc_l2norm_dropdim(const float *d_x, int64_t n, int32_t d, int64_t ld, const int32_t *d_drop,
                                  int32_t normalize, float *d_out, int64_t ldo, int32_t has_tail, float tail,
                                  vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0) return VSC_OK;
    if (!d_x || !d_out || d <= 0) { vsc::set_error("vsc_l2norm_dropdim: bad arguments"); return VSC_ERR_INVALID; }
    l2norm_dropdim_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(d_x, n, d, ld, d_drop, normalize, d_out, ldo,
                                                                       has_tail, tail);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

extern "C" int vsc_fill_column(float *d_out, int64_t n, int64_t ldo, int32_t col, const float *d_src, float factor,
                               vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0) return VSC_OK;
    fill_column_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_out, n, ldo, col, d_src, factor);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

// d_row / d_col / d_score: n hits sorted best first (VideoIndex.global_topk_device); d_q_vid / d_r_vid: video index of
// every query / reference row.  Writes the first `limit` video pairs in order of first appearance with their best
// score; *d_n_unique = number of distinct pairs.  d_table: vsc_pair_max_scratch_bytes(n) bytes.
extern "C" int64_t vsc_pair_max_scratch_bytes(int64_t n) {
    int64_t slots = 1024;
    while (slots < 2 * n) slots <<= 1;
    return slots * (int64_t)sizeof(Slot) + ((n + 1023) / 1024 + 1) * (int64_t)sizeof(unsigned int) + 64;
}
extern "C" int vsc_pair_max(const int64_t *d_row, const int64_t *d_col, const float *d_score, int64_t n,
                            const int32_t *d_q_vid, const int32_t *d_r_vid, int64_t n_ref_videos, int64_t limit,
                            int32_t *d_out_q, int32_t *d_out_r, float *d_out_score, unsigned long long *d_n_unique,
                            void *d_scratch, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n < 0 || n >= (1ll << 32) - 1 || !d_n_unique || !d_scratch) { vsc::set_error("vsc_pair_max: bad arguments"); return VSC_ERR_INVALID; }
    VSC_CUDA_CHECK(cudaMemsetAsync(d_n_unique, 0, sizeof(unsigned long long), stream));
    if (n == 0) return VSC_OK;
    int64_t slots = 1024;
    while (slots < 2 * n) slots <<= 1;
    Slot *table = static_cast<Slot *>(d_scratch);
    unsigned int *block_count = reinterpret_cast<unsigned int *>(table + slots);
    const int64_t n_blocks = (n + 1023) / 1024;
    pairmax_clear_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, stream>>>(table, slots);
    pairmax_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(d_row, d_col, d_q_vid, d_r_vid, n, n_ref_videos,
                                                                          table, slots - 1);
    pairmax_count_kernel<<<(unsigned)n_blocks, 256, 0, stream>>>(d_row, d_col, d_q_vid, d_r_vid, n, n_ref_videos, table,
                                                                 slots - 1, block_count);
    pairmax_scan_kernel<<<1, 1024, 0, stream>>>(block_count, n_blocks, d_n_unique);
    pairmax_scatter_kernel<<<(unsigned)n_blocks, 256, 0, stream>>>(d_row, d_col, d_q_vid, d_r_vid, d_score, n, n_ref_videos,
                                                                   table, slots - 1, block_count, limit, d_out_q, d_out_r,
                                                                   d_out_score);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch(5);
    return VSC_OK;
}

// ------------------------------------------------------------------ exact re-score of row-maximum candidates
// vsc_rowmax_rescore: for every candidate (row i, column j) the float32 inner product <a_i, b_j> over d dimensions -- one
// warp per candidate, lane l accumulates k = l, l+32, ... with fused multiply-adds in ascending k, then a fixed xor-shuffle
// tree: deterministic, independent of the batch -- and out[i] = max over the row's candidates (order-preserving integer
// keys + atomicMax).  Fillers of the emit epilogue (score = -inf, row = col = -1 or stale) are skipped by their score.
namespace {
__global__ void rowmax_init_kernel(uint32_t *__restrict__ key, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = vsc::float_to_key(-INFINITY);
}
__global__ void __launch_bounds__(256) rowmax_rescore_kernel(const float *__restrict__ a, int64_t lda, const float *__restrict__ b,
                                                             int64_t ldb, int d, const float *__restrict__ cand_score,
                                                             const int32_t *__restrict__ cand_row,
                                                             const int32_t *__restrict__ cand_col, int64_t n_cand,
                                                             int64_t m, int64_t n, uint32_t *__restrict__ key) {
    const int64_t c = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (c >= n_cand) return;
    if (!(cand_score[c] > -INFINITY)) return;                 // filler slot of a per-warp emit block
    const int64_t i = cand_row[c], j = cand_col[c];
    if (i < 0 || i >= m || j < 0 || j >= n) return;
    const float *x = a + i * lda, *y = b + j * ldb;
    float acc = 0.0f;
    for (int k = lane; k < d; k += 32) acc = __fmaf_rn(x[k], y[k], acc);
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(kFullMask, acc, s);
    if (lane == 0) atomicMax(&key[i], vsc::float_to_key(acc));
}
__global__ void rowmax_finish_kernel(const uint32_t *__restrict__ key, int64_t n, float *__restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = vsc::key_to_float(key[i]);
}
}  // namespace

extern "C" int vsc_rowmax_rescore(const float *d_a, int64_t m, int64_t lda, const float *d_b, int64_t n, int64_t ldb, int32_t d,
                                  const float *d_cand_score, const int32_t *d_cand_row, const int32_t *d_cand_col,
                                  int64_t n_cand, uint32_t *d_scratch_keys, float *d_out, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m <= 0) return VSC_OK;
    if (!d_a || !d_b || !d_scratch_keys || !d_out || (n_cand > 0 && (!d_cand_score || !d_cand_row || !d_cand_col))) {
        vsc::set_error("vsc_rowmax_rescore: null pointer"); return VSC_ERR_INVALID;
    }
    rowmax_init_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(d_scratch_keys, m);
    vsc::count_launch();
    if (n_cand > 0) {
        rowmax_rescore_kernel<<<(unsigned)((n_cand + 7) / 8), 256, 0, stream>>>(d_a, lda, d_b, ldb, d, d_cand_score, d_cand_row,
                                                                                 d_cand_col, n_cand, m, n, d_scratch_keys);
        vsc::count_launch();
    }
    rowmax_finish_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(d_scratch_keys, m, d_out);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
