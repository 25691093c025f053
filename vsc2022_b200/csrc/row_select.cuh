// Exact top-K of one similarity row by ONE thread (host + device).
//
// Contract (oracle/tn_networkx.py row_topk): the K largest values, best first;
// equal values keep the lower column first.  NaNs are not supported.
//
// Two passes over the row, which sits in shared memory on the device:
//   pass 1  per 16-column block: block maximum; a branch-free sorted insert
//           keeps the K largest block maxima.  The K-th of them, t, is a lower
//           bound for the K-th largest element (K distinct elements are >= t),
//           so every top-K element is >= t and lives in a block whose maximum
//           is >= t ("hot" block; exactly K of them unless maxima tie).
//   pass 2  only hot blocks are re-read; elements >= t (about K..K+3 of them)
//           go to a small candidate list in ascending column order.
//   select  stable insertion of the candidates into the sorted result.
// A row with more than kMaxCand candidates (e.g. a constant row) reports
// overflow and the caller routes the pair to the general kernel.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace vsc {

constexpr int kBlockCols = 16;  // columns per pass-1 block (4 x float4)
constexpr int kMaxCand = 16;    // candidate slots per row
constexpr int kMaxRowBlocks = 32;  // hot-block bitmask width -> rows up to 512 columns

__host__ __device__ __forceinline__ float max4(const float4 &v) {
    return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
}

// Streaming form used by the TN kernel: pass 1 is fed panel by panel while the row streams
// through shared memory; pass 2 re-reads the K hot blocks from `row` (global memory / L2).
template <int K>
struct RowTopK {
    float top[K];   // K largest block maxima so far, descending
    __host__ __device__ __forceinline__ void reset() {
#pragma unroll
        for (int i = 0; i < K; ++i) top[i] = -INFINITY;
    }
    // one 16-column block held in 4 float4 registers (n_chunks valid, 1..4)
    __host__ __device__ __forceinline__ float add_block(const float4 *chunk, int n_chunks) {
        float m = max4(chunk[0]);
#pragma unroll
        for (int c = 1; c < 4; ++c)
            if (c < n_chunks) m = fmaxf(m, max4(chunk[c]));
        float x = m;  // sorted insert (multiset), branch-free
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float hi = fmaxf(top[i], x);
            x = fminf(top[i], x);
            top[i] = hi;
        }
        return m;
    }
    // After every block went through add_block (block maxima in bm[b*stride]): finish the row.
    // `row` may be any 16-byte aligned pointer to the lr columns (global memory on the device: the
    // K hot blocks are L2 hits).  `stash`: 16 floats of per-thread scratch; candidate j lives at
    // cand_val[j*cstride] / cand_col[j*cstride].  Returns false on overflow.
    __host__ __device__ __forceinline__ bool finish(const float *row, int lr, const float *bm, int stride,
                                                    float *cand_val, int *cand_col, int cstride, float *stash,
                                                    float (&val)[K], int (&col)[K]) const {
        const float4 *row4 = reinterpret_cast<const float4 *>(row);
        float4 *stash4 = reinterpret_cast<float4 *>(stash);
        const int n_chunks = lr >> 2;
        const int n_blocks = (lr + kBlockCols - 1) / kBlockCols;
        const float t = top[K - 1];
        const float4 none = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        uint32_t hot = 0;
        for (int b = 0; b < n_blocks; ++b)
            if (bm[b * stride] >= t) hot |= 1u << b;
        int n_cand = 0;
        bool overflow = false;
        float4 nxt[4];
        int b_next = -1;
        auto fetch = [&]() {  // start loading the next hot block (exactly K of them unless maxima tie)
            if (!hot) { b_next = -1; return; }
#if defined(__CUDA_ARCH__)
            b_next = __ffs(hot) - 1;
#else
            b_next = __builtin_ctz(hot);
#endif
            hot &= hot - 1;
#pragma unroll
            for (int c = 0; c < 4; ++c) nxt[c] = b_next * 4 + c < n_chunks ? row4[b_next * 4 + c] : none;
        };
        fetch();
        while (b_next >= 0) {
            const int c0 = b_next * 4;
            uint32_t hits = 0;  // branch-free hit mask over the block's 16 columns
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                stash4[c] = nxt[c];
                hits |= ((nxt[c].x >= t ? 1u : 0u) | (nxt[c].y >= t ? 2u : 0u) | (nxt[c].z >= t ? 4u : 0u) |
                         (nxt[c].w >= t ? 8u : 0u)) << (4 * c);
            }
            fetch();  // overlaps the next block's L2 latency with the (short) hit loop
            while (hits) {
#if defined(__CUDA_ARCH__)
                const int k = __ffs(hits) - 1;
#else
                const int k = __builtin_ctz(hits);
#endif
                hits &= hits - 1;
                if (n_cand < kMaxCand) {
                    cand_val[n_cand * cstride] = stash[k];
                    cand_col[n_cand * cstride] = c0 * 4 + k;
                    ++n_cand;
                } else {
                    overflow = true;
                }
            }
        }
#pragma unroll
        for (int i = 0; i < K; ++i) { val[i] = -INFINITY; col[i] = 0x7fffffff; }
        for (int j = 0; j < n_cand; ++j) {  // candidates arrive in ascending column order
            float x = cand_val[j * cstride];
            int xc = cand_col[j * cstride];
            bool placed = false;
#pragma unroll
            for (int i = 0; i < K; ++i) {
                placed = placed || (x > val[i]);  // strict: an equal earlier column stays ahead
                if (placed) {
                    const float tv = val[i]; const int tc = col[i];
                    val[i] = x; col[i] = xc;
                    x = tv; xc = tc;
                }
            }
        }
        return !overflow;
    }
};

// Whole row at once (host tests, and rows that already sit in one buffer).
// row: 16-byte aligned, lr % 4 == 0, lr <= kBlockCols * kMaxRowBlocks, lr >= K.
// bm / cand_val / cand_col: per-thread scratch, element i at [i * stride].
template <int K>
__host__ __device__ __forceinline__ bool select_row(const float *row, int lr, float *bm, float *cand_val,
                                                    int *cand_col, int stride, float (&val)[K],
                                                    int (&col)[K]) {
    const float4 *row4 = reinterpret_cast<const float4 *>(row);
    const int n_chunks = lr >> 2;
    const int n_blocks = (lr + kBlockCols - 1) / kBlockCols;
    RowTopK<K> sel;
    sel.reset();
    for (int b = 0; b < n_blocks; ++b) {
        float4 chunk[4];
        const int left = n_chunks - b * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < left) chunk[c] = row4[b * 4 + c];
        bm[b * stride] = sel.add_block(chunk, left < 4 ? left : 4);
    }
    alignas(16) float stash[16];
    return sel.finish(row, lr, bm, stride, cand_val, cand_col, stride, stash, val, col);
}

}  // namespace vsc
