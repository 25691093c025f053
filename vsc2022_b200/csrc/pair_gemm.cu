// Per-pair frame similarity  S_p = Q_p . R_p^T + bias  on tcgen05 tensor cores, for a whole batch of candidate
// pairs in one launch (grouped GEMM), with the row top-K of the temporal network fused into the epilogue.
//
// Replaces, for the batch, LocalizationWithMetadata.similarity / VCSLLocalization.similarity
// (vsc/baseline/localization.py:33-36,49-54: np.matmul(q.feature, r.feature.T) + similarity_bias, called per pair at
// :57) and -- in TOPK mode -- the first step of VCSL's `tn` (np.argsort(-sims)[:, :top], vcsl/vta.py), so that the
// Lq x Lr similarity matrix never has to exist in memory: each 128 x bn accumulator tile is consumed straight out
// of tensor memory.
//
// Operands: K-major bf16 panels of ALL query frames [q_rows][K] and ALL reference frames [r_rows][K]
// (vsc_prepare_operand; K = kpad, or 3*kpad for the split that recovers fp32-class products); pair p multiplies rows
// [q_start[p], +lq[p]) by rows [r_start[p], +lr[p]).  One tensor map per panel: a tile that reaches past the end of its
// video simply loads rows of the next one (or zeros past the end of the panel), the epilogue masks them.
//
// Work unit = (pair, 128-row block); the unit's ceil(lr / bn) accumulator tiles live in different tensor-memory
// buffers (512 columns = up to 512 / bn buffers).  Persistent kernel, one CTA per SM, 192 threads:
//   warp 0      TMA producer, `stages`-deep ring of (128 x 64 | bn x 64) bf16 tiles, SWIZZLE_128B
//   warp 1      tcgen05.mma issuer (M128 x N bn x K16), commits free the ring slots / publish accumulators
//   warps 2-5   epilogue, one thread = one query frame (row):
//       STORE   fp32 tile + bias -> packed similarity buffer (when the caller wants the matrices, e.g. MaxSim scores)
//       TOPK    pass 1: 16-column block maxima, the K largest of them give a threshold t with >= K elements >= t;
//               pass 2 re-reads the accumulators from tensor memory and lists the elements >= t (about K of them),
//               a stable insertion (value descending, column ascending on ties: oracle/tn_networkx.py row_topk)
//               yields the exact top-K.  Rows with more than 16 candidates (heavy ties) and narrow matrices
//               (lr <= 128) take the plain exact insertion over every element instead.  Output = the node records the
//               graph stage reads (tn_common.cuh Workspace), identical to what tn_topk_kernel writes.
//   The unit's buffers are released after pass 2, so the MMAs of the next unit overlap this epilogue when the unit
//   leaves a buffer free (lr <= 320 with bn = 160: three buffers, two per unit).
#include "tc_common.cuh"
#include "tn_common.cuh"

namespace {

using namespace vsc::tc;
using vsc::kFullMask;
using vsc::tn::kMaxTop;
using vsc::tn::kRefMask;
using vsc::tn::kSimOk;
using vsc::tn::Workspace;

constexpr int kSub = 4;          // epilogue warps per tensor-memory lane quadrant (they share 32 rows, interleave the chunks)
constexpr int kEpiWarps = 4 * kSub;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kMaxBufs = 8;
constexpr int kMaxStages = 6;
constexpr int kCand = 32;        // candidate slots per row (all warps of the quadrant together)
constexpr int kSmallLr = 128;    // up to this many columns every element goes through the exact insertion
constexpr uint32_t kStageBytesA = BM * BK * 2;

struct PairArgs {
    const int32_t *q_start, *lq, *r_start, *lr;
    int n_pairs;
    int K;                 // inner dimension of the panels, multiple of 64
    int bn;                // accumulator tile width, multiple of 32, <= 256
    int m_tiles;           // ceil(max_lq / 128): units per pair
    int n_bufs;            // tensor-memory accumulators in use (n_bufs * bn <= 512)
    int stages;            // operand ring depth
    float bias;
    int ab_f16;            // fp16 instead of bf16 panels (vsc_gemm_format)
    int64_t ldq, ldr;      // panel row strides in elements (0: = K)
    const float *out_scale;   // device scalar multiplied into the accumulators (null: 1)
    // STORE
    float *sims;
    const int64_t *off;    // element offset of pair p (null: p * pair_stride)
    int64_t pair_stride;
    // TOPK
    int topk, max_nodes;
    float min_sim;
    Workspace w;
    int32_t *out_count, *out_list;   // pairs handed back (lr < topk)
};

struct Barriers {
    alignas(8) uint64_t full[kMaxStages], empty[kMaxStages], tmem_full[kMaxBufs], tmem_empty[kMaxBufs];
    uint32_t tmem_base;
};

// Per-CTA scratch of the TOPK epilogue; row index = quadrant * 32 + lane
struct TopkScratch {
    float part[kSub][kMaxTop][128];   // pass 1: every warp's K largest block maxima
    float thr[128];                   // merged threshold
    float cand_val[kCand][128];       // pass 2: elements >= threshold
    uint16_t cand_col[kCand][128];
    int n_cand[128];
};

__host__ __device__ inline size_t pair_smem_bytes(int bn, int stages, bool topk) {
    size_t b = (size_t)stages * (kStageBytesA + (size_t)bn * BK * 2);
    b += sizeof(Barriers);
    if (topk) b += sizeof(TopkScratch);
    return b + 1024 + 64;   // alignment slack
}

template <int K>
__device__ __forceinline__ void exact_insert(float x, int xc, float (&val)[K], int (&col)[K]) {
    bool placed = false;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        placed = placed || (x > val[i]);   // strict: an equal earlier (lower) column stays ahead
        if (placed) {
            const float tv = val[i]; const int tc = col[i];
            val[i] = x; col[i] = xc;
            x = tv; xc = tc;
        }
    }
}

// order-independent form (candidates of different warps interleave): value descending, then column ascending
template <int K>
__device__ __forceinline__ void exact_insert_any(float x, int xc, float (&val)[K], int (&col)[K]) {
    bool placed = false;
#pragma unroll
    for (int i = 0; i < K; ++i) {
        placed = placed || (x > val[i]) || (x == val[i] && xc < col[i]);
        if (placed) {
            const float tv = val[i]; const int tc = col[i];
            val[i] = x; col[i] = xc;
            x = tv; xc = tc;
        }
    }
}

__device__ __forceinline__ void quad_sync(int quad) {   // the kSub warps of one lane quadrant
    asm volatile("bar.sync %0, %1;" ::"r"(quad + 1), "r"(kSub * 32) : "memory");
}

__device__ __forceinline__ float max16(const float *v) {
    float m0 = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
    float m1 = fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7]));
    float m2 = fmaxf(fmaxf(v[8], v[9]), fmaxf(v[10], v[11]));
    float m3 = fmaxf(fmaxf(v[12], v[13]), fmaxf(v[14], v[15]));
    return fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

template <int K>
__device__ __forceinline__ void insert_max(float x, float (&top)[K]) {   // sorted insert into a multiset, branch-free
#pragma unroll
    for (int i = 0; i < K; ++i) {
        const float hi = fmaxf(top[i], x);
        x = fminf(top[i], x);
        top[i] = hi;
    }
}

// one 32-column chunk of the accumulator row of this thread: + bias, columns past the row end -> -inf
// (the output scale is a power of two, so acc * scale is exact and the fused multiply-add rounds once: fl(dot + bias))
__device__ __forceinline__ void load_chunk(uint32_t taddr, float osc, float bias, int valid, float (&v)[32]) {
    uint32_t raw[32];
    tmem_ld32(taddr, raw);
    if (valid >= 32) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __fmaf_rn(__uint_as_float(raw[j]), osc, bias);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < valid ? __fmaf_rn(__uint_as_float(raw[j]), osc, bias) : -INFINITY;
    }
}

__device__ __forceinline__ void store_chunk(float *dst, bool aligned, int valid, const float (&v)[32]) {
    if (aligned && valid >= 32) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            reinterpret_cast<float4 *>(dst)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < valid) dst[j] = v[j];
    }
}

template <int K, bool TOPK, bool STORE>
__global__ void __launch_bounds__(kThreads, 1) pair_gemm_kernel(const __grid_constant__ CUtensorMap tma_q,
                                                                 const __grid_constant__ CUtensorMap tma_r,
                                                                 const PairArgs g) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t stage_b_bytes = (uint32_t)g.bn * BK * 2;
    uint8_t *sm_a = base;                                        // [stages][16 KB]
    uint8_t *sm_b = sm_a + (size_t)g.stages * kStageBytesA;      // [stages][bn * 128 B]
    Barriers &bar = *reinterpret_cast<Barriers *>(sm_b + (size_t)g.stages * stage_b_bytes);
    TopkScratch &ts = *reinterpret_cast<TopkScratch *>((reinterpret_cast<uintptr_t>(&bar + 1) + 15) & ~(uintptr_t)15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t n_units = (int64_t)g.n_pairs * g.m_tiles;
    const int k_blocks = g.K / BK;
    const int chunks_per_tile = g.bn / 32;

    if (threadIdx.x == 0) {
        for (int s = 0; s < g.stages; ++s) { mbar_init(&bar.full[s], 1); mbar_init(&bar.empty[s], 1); }
        // TOPK: one warp per quadrant releases a unit's accumulators; STORE only: every warp, tile by tile
        for (int s = 0; s < g.n_bufs; ++s) { mbar_init(&bar.tmem_full[s], 1); mbar_init(&bar.tmem_empty[s], TOPK ? 4 : kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (TOPK && threadIdx.x >= 64 && threadIdx.x < 64 + 128) ts.n_cand[threadIdx.x - 64] = 0;
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bar.tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = bar.tmem_base;

    // Every role walks the same unit sequence and derives the same tile count from lq / lr.
    auto tiles_of = [&](int pair, int mt) -> int {
        const int lq = g.lq[pair], lr = g.lr[pair];
        if (mt * BM >= lq || lr <= 0) return 0;
        if (TOPK && lr < g.topk) return 0;
        return (lr + g.bn - 1) / g.bn;
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_q) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_r) : "memory");
            int stage = 0; uint32_t phase = 0;
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int pair = (int)(u / g.m_tiles), mt = (int)(u - (int64_t)pair * g.m_tiles);
                const int n_t = tiles_of(pair, mt);
                if (n_t == 0) continue;
                const int q0 = g.q_start[pair] + mt * BM, r0 = g.r_start[pair];
                for (int i = 0; i < n_t; ++i) {
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&bar.empty[stage], phase ^ 1);
                        mbar_expect_tx(&bar.full[stage], kStageBytesA + stage_b_bytes);
                        tma_load_2d(sm_a + (size_t)stage * kStageBytesA, &tma_q, kb * BK, q0, &bar.full[stage]);
                        tma_load_2d(sm_b + (size_t)stage * stage_b_bytes, &tma_r, kb * BK, r0 + i * g.bn, &bar.full[stage]);
                        if (++stage == g.stages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(BM, g.bn, !g.ab_f16);
            int stage = 0; uint32_t phase = 0;
            int buf = 0; uint32_t buf_phase = 0;   // accumulator ring position
            for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
                const int pair = (int)(u / g.m_tiles), mt = (int)(u - (int64_t)pair * g.m_tiles);
                const int n_t = tiles_of(pair, mt);
                for (int i = 0; i < n_t; ++i) {
                    mbar_wait(&bar.tmem_empty[buf], buf_phase ^ 1);   // epilogue has drained this accumulator
                    tcgen05_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * g.bn);
                    for (int kb = 0; kb < k_blocks; ++kb) {
                        mbar_wait(&bar.full[stage], phase);
                        tcgen05_fence_after();
                        const uint64_t da = umma_desc_k_major_sw128(smem_u32(sm_a + (size_t)stage * kStageBytesA));
                        const uint64_t db = umma_desc_k_major_sw128(smem_u32(sm_b + (size_t)stage * stage_b_bytes));
#pragma unroll
                        for (int k = 0; k < BK / UMMA_K; ++k)
                            tcgen05_mma_bf16(tmem_d, da + (uint64_t)(k * UMMA_K * 2 >> 4), db + (uint64_t)(k * UMMA_K * 2 >> 4),
                                             idesc, (kb | k) != 0);
                        tcgen05_commit(&bar.empty[stage]);
                        if (++stage == g.stages) { stage = 0; phase ^= 1; }
                    }
                    tcgen05_commit(&bar.tmem_full[buf]);
                    if (++buf == g.n_bufs) { buf = 0; buf_phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue warps: warp w may touch TMEM lanes [32*(w%4), +32).  The kSub warps of a quadrant own the
        // same 32 rows and take the 32-column chunks of a unit round-robin (chunk index % kSub == sub). =====
        const int quad = warp & 3, sub = (warp - 2) >> 2;
        const float osc = g.out_scale ? *g.out_scale : 1.0f;
        const int qrow = quad * 32 + lane;                  // row slot in the scratch arrays
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        int buf = 0; uint32_t buf_phase = 0;
        for (int64_t u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int pair = (int)(u / g.m_tiles), mt = (int)(u - (int64_t)pair * g.m_tiles);
            const int lq = g.lq[pair], lr = g.lr[pair];
            const int n_t = tiles_of(pair, mt);
            if (n_t == 0) {
                if (TOPK && mt == 0 && lq > 0 && lr < g.topk && warp == 2 && lane == 0 &&
                    atomicExch(&g.w.skip[pair], 1) == 0)
                    g.out_list[atomicAdd(g.out_count, 1)] = pair;
                continue;
            }
            const int row = mt * BM + quad * 32 + lane;
            const bool row_ok = row < lq;
            const int64_t pair_base = STORE ? (g.off ? g.off[pair] : (int64_t)pair * g.pair_stride) : 0;
            float *out_row = STORE ? g.sims + pair_base + (int64_t)row * lr : nullptr;
            const bool aligned = STORE && (lr & 3) == 0 && (pair_base & 3) == 0 &&
                                 (reinterpret_cast<uintptr_t>(g.sims) & 15u) == 0;
            const int buf0 = buf;

            float top[K], val[K]; int col[K];
#pragma unroll
            for (int i = 0; i < K; ++i) { top[i] = -INFINITY; val[i] = -INFINITY; col[i] = 0x7fffffff; }
            const bool small = TOPK && lr <= kSmallLr;

            // ---- pass 1 over this warp's chunks of the unit's accumulators (as they complete)
            for (int i = 0; i < n_t; ++i) {
                mbar_wait(&bar.tmem_full[buf], buf_phase);
                tcgen05_fence_after();
#pragma unroll 1
                for (int c = 0; c < chunks_per_tile; ++c) {
                    const int col0 = i * g.bn + c * 32;
                    if (col0 >= lr) break;
                    if (((i * chunks_per_tile + c) & (kSub - 1)) != sub) continue;
                    const int valid = min(32, lr - col0);
                    float v[32];
                    load_chunk(lane_base + (uint32_t)(buf * g.bn + c * 32), osc, g.bias, valid, v);
                    if (STORE && row_ok) store_chunk(out_row + col0, aligned, valid, v);
                    if (TOPK) {
                        if (small) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) exact_insert<K>(v[j], col0 + j, val, col);
                        } else {
                            insert_max<K>(max16(v), top);
                            insert_max<K>(max16(v + 16), top);
                        }
                    }
                }
                if (!TOPK) {   // STORE only: this warp is done with the accumulator
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar.tmem_empty[buf]);
                }
                if (++buf == g.n_bufs) { buf = 0; buf_phase ^= 1; }
            }
            if (!TOPK) continue;

            if (!small) {
#pragma unroll
                for (int i = 0; i < K; ++i) ts.part[sub][i][qrow] = top[i];
            }
            quad_sync(quad);   // B1: partial maxima visible; the previous unit's candidates have been consumed
            if (small) {
                // every warp's exact top-K of its own chunks goes to the merge as its candidates
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    if (row_ok && val[i] > -INFINITY) {
                        const int s = atomicAdd(&ts.n_cand[qrow], 1);
                        ts.cand_val[s][qrow] = val[i];
                        ts.cand_col[s][qrow] = (uint16_t)col[i];
                    }
                }
            } else {
                if (sub == 0) {   // threshold = K-th largest block maximum of the whole row: at least K elements are >= it
                    float m[K];
#pragma unroll
                    for (int i = 0; i < K; ++i) m[i] = -INFINITY;
#pragma unroll
                    for (int s = 0; s < kSub; ++s)
#pragma unroll
                        for (int i = 0; i < K; ++i) insert_max<K>(ts.part[s][i][qrow], m);
                    ts.thr[qrow] = row_ok ? m[K - 1] : INFINITY;
                }
                quad_sync(quad);   // B2: threshold visible
                // ---- pass 2: list the elements >= t of this warp's chunks
                const float t = ts.thr[qrow];
                int b2 = buf0;
                for (int i = 0; i < n_t; ++i) {
#pragma unroll 1
                    for (int c = 0; c < chunks_per_tile; ++c) {
                        const int col0 = i * g.bn + c * 32;
                        if (col0 >= lr) break;
                        if (((i * chunks_per_tile + c) & (kSub - 1)) != sub) continue;
                        const int valid = min(32, lr - col0);
                        float v[32];
                        load_chunk(lane_base + (uint32_t)(b2 * g.bn + c * 32), osc, g.bias, valid, v);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (v[j] >= t) {
                                const int s = atomicAdd(&ts.n_cand[qrow], 1);
                                if (s < kCand) {
                                    ts.cand_val[s][qrow] = v[j];
                                    ts.cand_col[s][qrow] = (uint16_t)(col0 + j);
                                }
                            }
                        }
                    }
                    if (++b2 == g.n_bufs) b2 = 0;
                }
            }
            tcgen05_fence_before();
            quad_sync(quad);   // B3: candidates complete, nobody reads the accumulators any more (except a redo below)
            if (sub != 0) continue;

            // ---- first warp of the quadrant: exact top-K of the candidates, release, node records
            {
                const int n_all = ts.n_cand[qrow];
                const bool redo = n_all > kCand;   // heavy ties: more candidates than slots
                const int n_c = redo ? 0 : n_all;
#pragma unroll
                for (int i = 0; i < K; ++i) { val[i] = -INFINITY; col[i] = 0x7fffffff; }
                const int n_max = __reduce_max_sync(kFullMask, n_c);
                for (int j = 0; j < n_max; ++j)
                    if (j < n_c) exact_insert_any<K>(ts.cand_val[j][qrow], (int)ts.cand_col[j][qrow], val, col);
                ts.n_cand[qrow] = 0;
                if (__any_sync(kFullMask, redo)) {   // plain exact insertion over the whole row
                    int b2 = buf0;
                    for (int i = 0; i < n_t; ++i) {
#pragma unroll 1
                        for (int c = 0; c < chunks_per_tile; ++c) {
                            const int col0 = i * g.bn + c * 32;
                            if (col0 >= lr) break;
                            const int valid = min(32, lr - col0);
                            float v[32];
                            load_chunk(lane_base + (uint32_t)(b2 * g.bn + c * 32), osc, g.bias, valid, v);
                            if (redo) {
#pragma unroll
                                for (int j = 0; j < 32; ++j) exact_insert<K>(v[j], col0 + j, val, col);
                            }
                        }
                        if (++b2 == g.n_bufs) b2 = 0;
                    }
                }
            }
            // the unit's accumulators are free (the other warps of the quadrant finished reading before B3)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
                int b3 = buf0;
                for (int i = 0; i < n_t; ++i) { mbar_arrive(&bar.tmem_empty[b3]); if (++b3 == g.n_bufs) b3 = 0; }
            }
            // ---- node records (same layout as tn_topk_kernel)
            if (row_ok) {
                const size_t node = (size_t)pair * g.max_nodes + (size_t)row * K;
                unsigned char *rec = static_cast<unsigned char *>(g.w.rec) + node * g.w.rec_bytes;
#pragma unroll
                for (int i = 0; i < K; ++i) {
                    g.w.ref_of[node + i] = (uint16_t)col[i] | (val[i] >= g.min_sim ? kSimOk : (uint16_t)0);
                    if (g.w.rec_bytes == 16) {
                        *reinterpret_cast<float4 *>(rec + i * 16) = make_float4(0.f, 0.f, val[i], 0.f);
                    } else {
                        *reinterpret_cast<float4 *>(rec + i * 32) = make_float4(0.f, 0.f, 0.f, 0.f);
                        *reinterpret_cast<float4 *>(rec + i * 32 + 16) = make_float4(val[i], 0.f, 0.f, 0.f);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// N tile for rows of up to max_lr columns: as few tiles as possible, equal widths, multiple of 32
void choose_tiles(int max_lr, bool topk, int *bn, int *n_bufs) {
    const int n = (max_lr + 255) / 256;
    int w = ((max_lr + n - 1) / n + 31) / 32 * 32;
    if (w < 32) w = 32;
    *bn = w;
    int b = 512 / w;
    if (b > kMaxBufs) b = kMaxBufs;
    if (!topk && b > 2) b = 2;   // STORE releases tile by tile: two accumulators are enough
    *n_bufs = b;
}

template <int K, bool TOPK, bool STORE>
int launch_k(const void *q_panel, int64_t q_rows, const void *r_panel, int64_t r_rows, PairArgs &g, int max_lr,
             cudaStream_t stream) {
    if (g.n_pairs <= 0) return VSC_OK;
    if (g.K <= 0 || g.K % BK != 0) { vsc::set_error("pair similarity: K=%d must be a positive multiple of %d", g.K, BK); return VSC_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(q_panel) & 15) || (reinterpret_cast<uintptr_t>(r_panel) & 15)) {
        vsc::set_error("pair similarity: panels must be 16-byte aligned"); return VSC_ERR_INVALID;
    }
    choose_tiles(max_lr, TOPK, &g.bn, &g.n_bufs);
    if (TOPK && (max_lr + g.bn - 1) / g.bn > g.n_bufs) {
        vsc::set_error("pair top-K: %d columns exceed the tensor-memory budget (512)", max_lr); return VSC_ERR_CAPACITY;
    }
    g.stages = 4;
    while (g.stages > 2 && pair_smem_bytes(g.bn, g.stages, TOPK) > 200 * 1024) --g.stages;
    CUtensorMap mq, mr;
    int rc = make_map(&mq, q_panel, q_rows, g.K, BM, g.ldq);
    if (rc != VSC_OK) return rc;
    rc = make_map(&mr, r_panel, r_rows, g.K, g.bn, g.ldr);
    if (rc != VSC_OK) return rc;
    const size_t smem = pair_smem_bytes(g.bn, g.stages, TOPK);
    VSC_CUDA_CHECK(cudaFuncSetAttribute(pair_gemm_kernel<K, TOPK, STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    VSC_CUDA_CHECK(cudaGetDevice(&dev));
    VSC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t units = (int64_t)g.n_pairs * g.m_tiles;
    const int grid = (int)(units < sms ? units : sms);
    pair_gemm_kernel<K, TOPK, STORE><<<grid, kThreads, smem, stream>>>(mq, mr, g);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

template <bool STORE>
int launch_topk(const void *q_panel, int64_t q_rows, const void *r_panel, int64_t r_rows, PairArgs &g, int max_lr,
                cudaStream_t stream) {
    switch (g.topk) {
        case 1: return launch_k<1, true, STORE>(q_panel, q_rows, r_panel, r_rows, g, max_lr, stream);
        case 2: return launch_k<2, true, STORE>(q_panel, q_rows, r_panel, r_rows, g, max_lr, stream);
        case 3: return launch_k<3, true, STORE>(q_panel, q_rows, r_panel, r_rows, g, max_lr, stream);
        case 4: return launch_k<4, true, STORE>(q_panel, q_rows, r_panel, r_rows, g, max_lr, stream);
        case 5: return launch_k<5, true, STORE>(q_panel, q_rows, r_panel, r_rows, g, max_lr, stream);
        case 6: return launch_k<6, true, STORE>(q_panel, q_rows, r_panel, r_rows, g, max_lr, stream);
        case 7: return launch_k<7, true, STORE>(q_panel, q_rows, r_panel, r_rows, g, max_lr, stream);
        default: return launch_k<8, true, STORE>(q_panel, q_rows, r_panel, r_rows, g, max_lr, stream);
    }
}

}  // namespace

namespace vsc {
namespace tn {

// Row top-K of every pair straight from the descriptor panels (the similarity matrices are written to `sims` only when
// it is non-null).  Fills the node records of `w`; pairs it cannot take (fewer than topk columns) go to `out`.
int launch_pair_topk(const PairOperands &op, const Batch &b, const Workspace &w, const WorkList &out, float *sims,
                     const int64_t *off, int64_t pair_stride, cudaStream_t stream) {
    PairArgs g = {};
    g.q_start = op.q_start; g.lq = b.lq; g.r_start = op.r_start; g.lr = b.lr; g.n_pairs = b.n_pairs;
    g.K = op.k; g.m_tiles = (b.max_lq + BM - 1) / BM; g.bias = op.bias;
    g.ab_f16 = op.ab_f16; g.ldq = op.ldq; g.ldr = op.ldr; g.out_scale = op.out_scale;
    g.sims = sims; g.off = off; g.pair_stride = pair_stride;
    g.topk = b.topk; g.max_nodes = b.max_nodes; g.min_sim = b.min_sim; g.w = w;
    g.out_count = out.count; g.out_list = out.list;
    if (sims) return launch_topk<true>(op.q_panel, op.q_rows, op.r_panel, op.r_rows, g, b.max_lr, stream);
    return launch_topk<false>(op.q_panel, op.q_rows, op.r_panel, op.r_rows, g, b.max_lr, stream);
}

bool pair_topk_supported(const Batch &b) { return b.max_lr <= 512 && b.topk >= 1 && b.topk <= kMaxTop; }

}  // namespace tn
}  // namespace vsc

// sims[off[p] + i * lr[p] + j] = Q[q_start[p] + i] . R[r_start[p] + j] + bias
extern "C" int vsc_pair_similarity(const void *d_q_panel, int64_t q_rows, const void *d_r_panel, int64_t r_rows, int32_t k,
                                   const int32_t *d_q_start, const int32_t *d_lq, const int32_t *d_r_start,
                                   const int32_t *d_lr, int32_t n_pairs, int32_t max_lq, int32_t max_lr, float bias,
                                   float *d_sims, const int64_t *d_off, const vsc_gemm_format *fmt, vsc_stream_t stream) {
    if (n_pairs <= 0 || max_lq <= 0 || max_lr <= 0) return VSC_OK;
    if (!d_q_panel || !d_r_panel || !d_q_start || !d_lq || !d_r_start || !d_lr || !d_sims || !d_off) {
        vsc::set_error("vsc_pair_similarity: null pointer"); return VSC_ERR_INVALID;
    }
    PairArgs g = {};
    g.q_start = d_q_start; g.lq = d_lq; g.r_start = d_r_start; g.lr = d_lr; g.n_pairs = n_pairs;
    g.K = k; g.m_tiles = (max_lq + BM - 1) / BM; g.bias = bias;
    if (fmt) { g.ab_f16 = fmt->ab_f16; g.ldq = fmt->lda; g.ldr = fmt->ldb; g.out_scale = fmt->d_out_scale; }
    g.sims = d_sims; g.off = d_off;
    return launch_k<1, false, true>(d_q_panel, q_rows, d_r_panel, r_rows, g, max_lr, static_cast<cudaStream_t>(stream));
}
