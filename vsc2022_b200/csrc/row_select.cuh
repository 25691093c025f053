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
//           is >= t ("hot" block; ~K of them).
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

// row: 16-byte aligned, lr % 4 == 0, lr <= kBlockCols * kMaxRowBlocks, lr >= K.
// bm / cand_val / cand_col: per-thread scratch, element i at [i * stride].
template <int K>
__host__ __device__ __forceinline__ bool select_row(const float *row, int lr, float *bm, float *cand_val,
                                                    int *cand_col, int stride, float (&val)[K],
                                                    int (&col)[K]) {
    const float4 *row4 = reinterpret_cast<const float4 *>(row);
    const int n_chunks = lr >> 2;
    const int n_blocks = (lr + kBlockCols - 1) / kBlockCols;
    float top[K];
#pragma unroll
    for (int i = 0; i < K; ++i) top[i] = -INFINITY;
    uint32_t hot = 0;
    for (int b = 0; b < n_blocks; ++b) {
        const int c0 = b * 4;
        float m = max4(row4[c0]);
#pragma unroll
        for (int c = 1; c < 4; ++c)
            if (c0 + c < n_chunks) m = fmaxf(m, max4(row4[c0 + c]));
        bm[b * stride] = m;
        if (m >= top[K - 1]) hot |= 1u << b;  // may still be hot once t is final
        float x = m;                          // sorted insert (multiset), branch-free
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const float hi = fmaxf(top[i], x);
            x = fminf(top[i], x);
            top[i] = hi;
        }
    }
    const float t = top[K - 1];
    int n_cand = 0;
    bool overflow = false;
    while (hot) {
#if defined(__CUDA_ARCH__)
        const int b = __ffs(hot) - 1;
#else
        const int b = __builtin_ctz(hot);
#endif
        hot &= hot - 1;
        if (!(bm[b * stride] >= t)) continue;
        const int c0 = b * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            if (c0 + c >= n_chunks) break;
            const float4 v = row4[c0 + c];
            const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (e[k] >= t) {
                    if (n_cand < kMaxCand) {
                        cand_val[n_cand * stride] = e[k];
                        cand_col[n_cand * stride] = (c0 + c) * 4 + k;
                        ++n_cand;
                    } else {
                        overflow = true;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < K; ++i) { val[i] = -INFINITY; col[i] = 0x7fffffff; }
    for (int j = 0; j < n_cand; ++j) {  // candidates arrive in ascending column order
        float x = cand_val[j * stride];
        int xc = cand_col[j * stride];
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

}  // namespace vsc
