// Descriptor similarity GEMM on 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
//
//   S[i][j] = sum_k A[i][k] * B[j][k]      A: queries [M][K] bf16, B: references [N][K] bf16,
//                                           both K-major (row-major, K contiguous), fp32 accumulate.
//
// Replaces the FAISS flat-index arithmetic behind vsc/index.py:142-177 (range / kNN search) and
// vsc/baseline/score_normalization.py:93-96 (1-NN against the noise set).  The similarity matrix is
// never written to memory: each 128x256 accumulator tile is consumed straight out of tensor memory
// by one of three fused epilogues
//   STORE   write the fp32 tile (tests, small per-pair matrices)
//   ROWMAX  per-row maximum over all references (score normalisation's 1-NN similarity)
//   EMIT    count scores beyond `count_thr`, append (score, i, j) of those beyond `emit_thr`
//           (strict comparisons, inner product or squared-L2 metric)
//
// Kernel anatomy (persistent, one CTA per SM, 192 threads; 320 for the convolution epilogue):
//   warp 0      TMA producer: cp.async.bulk.tensor 2D tiles (SWIZZLE_128B) into a 4-stage ring
//   warp 1      allocates TMEM; one lane issues tcgen05.mma (M128 x N256 x K16, cta_group::1)
//   warps 2-5   epilogue: tcgen05.ld 32 columns at a time from one of two accumulator buffers
//   (warps 6-9  CONV only: a second epilogue warp per TMEM lane quadrant taking every other 32-column chunk --
//               short-K convolutions are bound by the epilogue's dependent chain, not by the MMAs)
//               (512 TMEM columns), so the epilogue of tile t overlaps the MMAs of tile t+1
// Tile order: m fastest, so concurrently running CTAs share the same reference tile in L2 while the
// whole query matrix stays L2-resident; references stream from HBM once.
#include "tc_common.cuh"

namespace {

using vsc::kFullMask;
using namespace vsc::tc;   // BM = 128, BK = 64 (the N tile, 64 / 128 / 256, is a template parameter), PTX wrappers

constexpr int STAGES = 4;
#ifndef VSC_CONV_EPI_WARPS
#define VSC_CONV_EPI_WARPS 8   // 12 (three per lane quadrant, 448 threads) measured: no gain, `down` GEMMs slower
#endif
// CONV and EMIT epilogues are long dependent chains per 32-column chunk (transposes / hit compaction): two warps per
// TMEM lane quadrant, alternating chunks.  ROWMAX-type epilogues are short and keep four warps.
__host__ __device__ constexpr int epi_warps(int epi) { return epi == 4 /* EPI_CONV */ ? VSC_CONV_EPI_WARPS : epi == 2 /* EPI_EMIT */ ? 8 : 4; }
__host__ __device__ constexpr int cta_threads(int epi) { return 64 + 32 * epi_warps(epi); }   // producer + MMA + epilogue warps
constexpr uint32_t kStageBytesA = BM * BK * 2;

enum ALoad { A_TILED = 0, A_IM2COL = 1, A_SHIFT = 2 };  // how the producer fetches the A tile of a k-block
enum Epilogue { EPI_STORE = 0, EPI_ROWMAX = 1, EPI_EMIT = 2, EPI_ROWARGMAX = 3, EPI_CONV = 4 };

struct GemmArgs {
    int64_t M, N;
    int K;  // multiple of BK
    // STORE (fp32, optional per-column bias) and CONV (bf16 out = relu?(acc + bias[col] + residual))
    float *c; int64_t ldc;
    const float *bias;
    const __nv_bfloat16 *residual;
    __nv_bfloat16 *out_bf16;
    int relu;
    // ROWMAX: order-preserving keys (vsc::float_to_key), combined with atomicMax
    uint32_t *rowmax_key;
    // ROWARGMAX: (key << 32) | (0xFFFFFFFF - column): atomicMax keeps the best score, lowest column on ties
    unsigned long long *rowbest;
    // EMIT
    const float *a_norm, *b_norm;  // squared norms (L2 metric) or null
    int metric_l2;
    float count_thr, emit_thr;
    const float *thr_ptr;            // when set: {count_thr, emit_thr} are read from device memory (device-driven search)
    const float *row_thr;            // when set (inner product only): row i emits (and counts) the scores beyond row_thr[i]
    const float *thr_margin;         // when set (inner product only): both thresholds are loosened by *thr_margin (candidate pass)
    int64_t row_offset, col_offset;  // added to the emitted indices
    float *out_score; int32_t *out_row, *out_col;
    unsigned long long capacity;
    unsigned long long *counters;  // [0] entries claimed (may exceed capacity), [1] hits beyond count_thr
    // implicit 3x3 convolution (A operand loaded by TMA in im2col mode from the NHWC activation tensor):
    // output extent, traversal stride and 64-channel blocks per filter tap
    int conv_ho, conv_wo, conv_stride, conv_cblocks, conv_ksize, conv_pad;
    // A_SHIFT (stem): k-block kb reads rows m + kb * a_row_shift of an overlapping-row view (see vsc_conv_stem)
    int64_t a_row_shift;
    // operand format (vsc_gemm_format): fp16 instead of bf16 panels, panel row strides, accumulator scale
    int ab_f16;
    int64_t lda, ldb;
    const float *out_scale;
};

template <int BN>
struct SharedStorage {
    alignas(1024) uint8_t a[STAGES][kStageBytesA];
    alignas(1024) uint8_t b[STAGES][BN * BK * 2];
    alignas(8) uint64_t full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    // CONV epilogue, per epilogue warp: a 32-row x 64-byte transpose buffer (16-byte units, XOR-swizzled: see
    // conv_unit) used first for the residual chunk coming in, then for the bf16 chunk going out; and the folded
    // BatchNorm bias of the warp's (up to four) 32-column chunks
    alignas(16) uint4 stage[12][128];
    alignas(16) float bias_s[12][4 * 32];
};

// EMIT output slots are claimed per WARP in blocks of kEmitBlock entries: one global atomic per block instead of one
// per chunk with a hit (measured: ~3 M same-address atomics per launch cost 0.7-1.4 ms, more than the GEMM of a 2048-row
// batch).  Unused slots of a warp's last block are filled with a score no threshold accepts (-inf / +inf), so the
// consumer's strict re-filter drops them; counters[0] therefore counts CLAIMED slots, counters[1] true hits.
constexpr int kEmitBlock = 256;
struct EmitState {
    unsigned long long pos;     // next free slot of this warp's block
    int left;                   // free slots in the block
    unsigned long long counted; // hits beyond count_thr seen by this warp (flushed once at the end)
};
// squared L2 distance from the norms and the inner product; one expression for every place that needs the value
__device__ __forceinline__ float l2_score(float an, float bn, float ip) { return __fmaf_rn(-2.0f, ip, __fadd_rn(an, bn)); }
__device__ __forceinline__ void emit_pad(const GemmArgs &g, EmitState &e, int lane) {   // retire the current block
    const float never = g.metric_l2 ? INFINITY : -INFINITY;
    for (int i = lane; i < e.left; i += 32)
        if (e.pos + i < g.capacity) g.out_score[e.pos + i] = never;
    e.left = 0;
}

// ---------------------------------------------------------------- epilogues (one thread = one accumulator row)
// `valid` = number of in-range columns of this 32-column chunk (1..32): bounds are applied to the
// bit masks, not per element.
template <int EPI>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs &g, float &best, int64_t &best_col, int64_t row,
                                               int64_t col0, int valid, const uint32_t (&acc)[32], int lane,
                                               EmitState &emit, uint32_t taddr, float osc) {
    const bool row_ok = row < g.M;
    const uint32_t valid_mask = valid >= 32 ? 0xFFFFFFFFu : ((1u << valid) - 1u);
    if (EPI == EPI_STORE) {
        if (row_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < valid) g.c[row * g.ldc + col0 + j] = __fmaf_rn(__uint_as_float(acc[j]), osc, g.bias ? g.bias[col0 + j] : 0.0f);
        }
    } else if (EPI == EPI_ROWARGMAX) {
        int arg = -1;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(acc[j]);
            if (j < valid && v > best) { best = v; arg = j; }  // strict: the first (lowest) column wins ties
        }
        if (arg >= 0) best_col = col0 + arg;
    } else if (EPI == EPI_ROWMAX) {
        if (valid >= 32) {
#pragma unroll
            for (int j = 0; j < 32; ++j) best = fmaxf(best, __uint_as_float(acc[j]));
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < valid) best = fmaxf(best, __uint_as_float(acc[j]));
        }
    } else {
        // EMIT: strict comparisons; IP keeps larger scores, squared-L2 keeps smaller distances
        uint32_t hits = 0, counted = 0;
        float s[32];
        // inner product: the comparisons run on the raw accumulators against thresholds divided by the (power-of-two,
        // hence exact) output scale; only the emitted scores are scaled
        const float inv = g.metric_l2 ? 1.0f : 1.0f / osc;
        const float loosen = g.thr_margin ? *g.thr_margin : 0.0f;
        float emit_thr = ((g.thr_ptr ? g.thr_ptr[1] : g.emit_thr) - loosen) * inv, count_thr = ((g.thr_ptr ? g.thr_ptr[0] : g.count_thr) - loosen) * inv;
        const bool two = !g.row_thr && emit_thr != count_thr;  // uniform: the common case has one threshold
        if (g.row_thr) emit_thr = count_thr = (row_ok ? g.row_thr[row] : INFINITY) * inv;   // one threshold per query row
        float an = 0.0f;
        if (!g.metric_l2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                s[j] = __uint_as_float(acc[j]);
                hits |= (s[j] > emit_thr ? 1u : 0u) << j;
            }
            counted = hits;
            if (two) {
                counted = 0;
#pragma unroll
                for (int j = 0; j < 32; ++j) counted |= (s[j] > count_thr ? 1u : 0u) << j;
            }
        } else {
            an = row_ok ? g.a_norm[row] : 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float bn = j < valid ? g.b_norm[col0 + j] : 0.0f;
                s[j] = l2_score(an, bn, __uint_as_float(acc[j]) * osc);
                hits |= (s[j] < emit_thr ? 1u : 0u) << j;
                counted |= (s[j] < count_thr ? 1u : 0u) << j;
            }
        }
        const uint32_t keep = row_ok ? valid_mask : 0u;
        hits &= keep; counted &= keep;
        if (!__any_sync(kFullMask, (hits | counted) != 0)) return;  // the common case once the radius is tight
        // warp-aggregated claim of output slots out of the warp's private block
        const int n_hit = __popc(hits);
        int incl = n_hit, total_cnt = __popc(counted);
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int up = __shfl_up_sync(kFullMask, incl, d);
            if (lane >= d) incl += up;
            total_cnt += __shfl_xor_sync(kFullMask, total_cnt, d);
        }
        const int total_hit = __shfl_sync(kFullMask, incl, 31);
        emit.counted += (unsigned long long)total_cnt;
        if (total_hit == 0) return;
        if (total_hit > emit.left) {
            emit_pad(g, emit, lane);
            const int want = total_hit > kEmitBlock ? total_hit : kEmitBlock;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&g.counters[0], (unsigned long long)want);
            emit.pos = __shfl_sync(kFullMask, base, 0);
            emit.left = want;
        }
        unsigned long long at = emit.pos + (unsigned long long)(incl - n_hit);
        emit.pos += (unsigned long long)total_hit;
        emit.left -= total_hit;
        if (total_hit <= 8) {
            // a handful of hits in the chunk (the usual case once the radius is tight): walk the hit COLUMNS
            // (warp-uniform) and fetch each one again from tensor memory instead of running the 32-way predicated
            // store sequence below (~250 instructions per chunk with a single hit)
            uint32_t cols = __reduce_or_sync(kFullMask, hits);
            while (cols) {
                const int j = __ffs(cols) - 1;
                cols &= cols - 1;
                uint32_t raw;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(raw) : "r"(taddr + (uint32_t)j));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if ((hits >> j) & 1) {
                    const float v = g.metric_l2 ? l2_score(an, g.b_norm[col0 + j], __uint_as_float(raw) * osc) : __uint_as_float(raw) * osc;
                    if (at < g.capacity) {
                        g.out_score[at] = v;
                        g.out_row[at] = (int32_t)(row + g.row_offset);
                        g.out_col[at] = (int32_t)(col0 + j + g.col_offset);
                    }
                    ++at;
                }
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if ((hits >> j) & 1) {
                if (at < g.capacity) {
                    g.out_score[at] = g.metric_l2 ? s[j] : s[j] * osc;
                    g.out_row[at] = (int32_t)(row + g.row_offset);
                    g.out_col[at] = (int32_t)(col0 + j + g.col_offset);
                }
                ++at;
            }
        }
    }
}

// CONV epilogue data path.  A thread owns one accumulator ROW (tcgen05.ld 32x32b), so direct global accesses would
// touch 32 different 128-byte lines per warp instruction (measured: the K=64 -> N=256 expansion convolutions ran at
// 2.3 TB/s, LSU-wavefront-bound).  Instead the 32 x 64-byte chunk is transposed through shared memory: global
// accesses are issued with 4 lanes per row (64 contiguous bytes, 8 rows per instruction), the row-per-thread
// accesses go to shared memory.  Unit (row, q) lives at row*4 + (q ^ ((row >> 1) & 3)): conflict-free both ways.
__device__ __forceinline__ int conv_unit(int row, int q) { return row * 4 + (q ^ ((row >> 1) & 3)); }
// explicit shared-space accesses: through a generic pointer these compile to LD.E / ST.E, which take the generic
// address path and wait on the long scoreboard like global loads
__device__ __forceinline__ void sts128(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}

// Everything of the CONV epilogue's addressing that does not change from chunk to chunk: the shared-memory byte
// addresses of the two access patterns of the transpose buffer (lane-invariant: computed once per kernel) and, per tile,
// the byte offsets of this lane's four (row, 16-byte segment) pieces in the residual / output matrices.  The first
// version recomputed all of it per chunk: ~690 integer instructions in the kernel, more than its floating-point work.
struct ConvCtx {
    uint32_t own[4];       // smem address of unit (lane, q): the row-per-thread pattern
    uint32_t tr[4];        // smem address of unit (it*8 + lane/4, lane%4): the 4-lanes-per-row pattern
    uint32_t bias;         // smem address of this warp's bias block
    int64_t piece[4];      // byte offset of (row0 + it*8 + lane/4, column colb + (lane%4)*8) in a [M][ldc] bf16 matrix
    uint32_t row_ok;       // bit it: that row is < M
};
__device__ __forceinline__ void conv_ctx_init(ConvCtx &cx, uint4 *stage, const float *bias_s, int lane) {
    const uint32_t base = smem_u32(stage);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        cx.own[q] = base + conv_unit(lane, q) * 16;
        cx.tr[q] = base + conv_unit(q * 8 + (lane >> 2), lane & 3) * 16;
    }
    cx.bias = smem_u32(bias_s);
}
__device__ __forceinline__ void conv_ctx_tile(ConvCtx &cx, const GemmArgs &g, int64_t row0, int64_t colb, int lane) {
    cx.row_ok = 0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int64_t row = row0 + it * 8 + (lane >> 2);
        cx.piece[it] = (row * g.ldc + colb + (lane & 3) * 8) * 2;
        cx.row_ok |= (row < g.M ? 1u : 0u) << it;
    }
}
// chunk c of the tile starts c * 64 bytes into every piece
__device__ __forceinline__ void conv_residual_fetch(const GemmArgs &g, const ConvCtx &cx, int c, uint4 (&r)[4]) {
    const char *base = reinterpret_cast<const char *>(g.residual) + c * 64;
#pragma unroll
    for (int it = 0; it < 4; ++it)
        r[it] = (cx.row_ok >> it) & 1 ? __ldg(reinterpret_cast<const uint4 *>(base + cx.piece[it])) : make_uint4(0, 0, 0, 0);
}

// Pull the residual rows of a whole tile (this warp's chunks) from HBM into L2 one tile ahead: the register
// prefetch above only reaches one chunk (~0.5 us) ahead, less than an HBM round trip under load.
__device__ __forceinline__ void conv_residual_prefetch_l2(const GemmArgs &g, int64_t row0, int64_t colb, int first,
                                                          int chunks, int lane) {
    const int64_t row = row0 + lane;
    if (row >= g.M) return;
    const __nv_bfloat16 *p = g.residual + row * g.ldc + colb;
    for (int c = first; c < chunks; c += epi_warps(4) / 4)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p + c * 32));
}

__device__ __forceinline__ void conv_epilogue_chunk(const GemmArgs &g, const ConvCtx &cx, int c, int j,
                                                    const uint32_t (&acc)[32], const uint4 (&res)[4]) {
    float v[32];
    const uint32_t bias_s = cx.bias + j * 128;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const uint4 bq = lds128(bias_s + q * 16);   // same address in every lane: broadcast
        v[q * 4 + 0] = __uint_as_float(acc[q * 4 + 0]) + __uint_as_float(bq.x);
        v[q * 4 + 1] = __uint_as_float(acc[q * 4 + 1]) + __uint_as_float(bq.y);
        v[q * 4 + 2] = __uint_as_float(acc[q * 4 + 2]) + __uint_as_float(bq.z);
        v[q * 4 + 3] = __uint_as_float(acc[q * 4 + 3]) + __uint_as_float(bq.w);
    }
    if (g.residual) {
#pragma unroll
        for (int it = 0; it < 4; ++it) sts128(cx.tr[it], res[it]);
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 r = lds128(cx.own[q]);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                v[q * 8 + 2 * t] += __uint_as_float(w[t] << 16);
                v[q * 8 + 2 * t + 1] += __uint_as_float(w[t] & 0xFFFF0000u);
            }
        }
        __syncwarp();   // every lane has its residual row: the buffer may be overwritten
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            // ReLU after the rounding (identical result: rounding keeps the sign), on the packed pair
            __nv_bfloat162 p = __floats2bfloat162_rn(v[q * 8 + 2 * t], v[q * 8 + 2 * t + 1]);
            if (g.relu) p = __hmax2(p, __floats2bfloat162_rn(0.0f, 0.0f));
            w[t] = *reinterpret_cast<const uint32_t *>(&p);
        }
        sts128(cx.own[q], make_uint4(w[0], w[1], w[2], w[3]));
    }
    __syncwarp();
    char *out = reinterpret_cast<char *>(g.out_bf16) + c * 64;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const uint4 o = lds128(cx.tr[it]);
        if ((cx.row_ok >> it) & 1) *reinterpret_cast<uint4 *>(out + cx.piece[it]) = o;
    }
    __syncwarp();   // the buffer is free again
}

template <int EPI, int BN, int ALOAD = A_TILED>
__global__ void __launch_bounds__(cta_threads(EPI), 1) gemm_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                           const __grid_constant__ CUtensorMap tma_b,
                                                           const GemmArgs g) {
    extern __shared__ uint8_t smem_raw[];
    using Storage = SharedStorage<BN>;
    Storage &sm = *reinterpret_cast<Storage *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr uint32_t kStageBytesB = BN * BK * 2;
    constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;  // two fp32 accumulators of BN columns
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t m_tiles = (g.M + BM - 1) / BM, n_tiles = (g.N + BN - 1) / BN;
    const int64_t tiles = m_tiles * n_tiles;
    const int k_blocks = g.K / BK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&sm.tmem_full[s], 1); mbar_init(&sm.tmem_empty[s], epi_warps(EPI)); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // one full warp allocates all 512 TMEM columns and publishes the base address
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(kTmemCols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = sm.tmem_base;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_a) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tma_b) : "memory");
            int stage = 0; uint32_t phase = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
                const int m_blk = (int)(t % m_tiles), n_blk = (int)(t / m_tiles);
                int px = 0, py = 0, pn = 0, tap = 0, cb = 0;
                if (ALOAD == A_IM2COL) {   // first output pixel of the tile -> base pixel of its filter window
                    const int64_t m0 = (int64_t)m_blk * BM, row = m0 / g.conv_wo;
                    px = (int)(m0 - row * g.conv_wo) * g.conv_stride - g.conv_pad;
                    pn = (int)(row / g.conv_ho);
                    py = (int)(row - (int64_t)pn * g.conv_ho) * g.conv_stride - g.conv_pad;
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&sm.empty[stage], phase ^ 1);
                    mbar_expect_tx(&sm.full[stage], kStageBytesA + kStageBytesB);
                    if (ALOAD == A_SHIFT) {
                        tma_load_2d(sm.a[stage], &tma_a, 0, (int)((int64_t)m_blk * BM + kb * g.a_row_shift), &sm.full[stage]);
                    } else if (ALOAD == A_IM2COL) {   // K index = (ky*ksize + kx)*C + channel
                        tma_load_im2col(sm.a[stage], &tma_a, cb * BK, px, py, pn, (uint16_t)(tap % g.conv_ksize), (uint16_t)(tap / g.conv_ksize),
                                        &sm.full[stage]);
                        if (++cb == g.conv_cblocks) { cb = 0; ++tap; }
                    } else
                    tma_load_2d(sm.a[stage], &tma_a, kb * BK, m_blk * BM, &sm.full[stage]);
                    tma_load_2d(sm.b[stage], &tma_b, kb * BK, n_blk * BN, &sm.full[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (single thread) =====
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16(BM, BN, !g.ab_f16);
            int stage = 0; uint32_t phase = 0;
            uint32_t it = 0;
            for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
                const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
                mbar_wait(&sm.tmem_empty[acc], acc_phase ^ 1);  // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&sm.full[stage], phase);          // TMA bytes have landed
                    tcgen05_fence_after();
                    const uint64_t da = umma_desc_k_major_sw128(smem_u32(sm.a[stage]));
                    const uint64_t db = umma_desc_k_major_sw128(smem_u32(sm.b[stage]));
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; ++k)       // +32 bytes along K inside the swizzle row
                        tcgen05_mma_bf16(tmem_d, da + (uint64_t)(k * UMMA_K * 2 >> 4), db + (uint64_t)(k * UMMA_K * 2 >> 4),
                                         idesc, (kb | k) != 0);
                    tcgen05_commit(&sm.empty[stage]);           // frees the smem slot once these MMAs retire
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(&sm.tmem_full[acc]);             // accumulator complete -> epilogue
            }
        }
    } else {
        // ===== epilogue warps: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) =====
        const int quad = warp & 3;
        const float osc = g.out_scale ? *g.out_scale : 1.0f;
        int64_t bias_blk = -1;
        EmitState emit = {0ull, 0, 0ull};
        ConvCtx cx;
        if (EPI == EPI_CONV) conv_ctx_init(cx, sm.stage[warp - 2], sm.bias_s[warp - 2], lane);
        uint32_t it = 0;
        for (int64_t t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
            const uint32_t acc = it & 1, acc_phase = (it >> 1) & 1;
            const int64_t m_blk = t % m_tiles, n_blk = t / m_tiles;
            const int64_t row = m_blk * BM + quad * 32 + lane;
            float best = -INFINITY;
            int64_t best_col = -1;
            mbar_wait(&sm.tmem_full[acc], acc_phase);
            tcgen05_fence_after();
            if (EPI == EPI_CONV) {
                // N is a multiple of 32 (checked on the host), so a chunk is all-or-nothing.  The residual of chunk
                // c+1 is requested before chunk c is processed, so its HBM latency hides behind the TMEM read,
                // the arithmetic and the stores of chunk c.
                const int64_t row0 = m_blk * BM + quad * 32, colb = n_blk * BN;
                const int chunks = (int)((g.N - colb < BN ? g.N - colb : BN) / 32);
                constexpr int kPer = epi_warps(EPI_CONV) / 4;   // warps per lane quadrant, interleaved over the chunks
                const int epi = warp - 2, first = epi >> 2;
                if (g.residual && t + gridDim.x < tiles) {    // next tile of this CTA: its residual goes to L2 now
                    const int64_t t2 = t + gridDim.x, m2 = t2 % m_tiles, n2 = t2 / m_tiles;
                    conv_residual_prefetch_l2(g, m2 * BM + quad * 32, n2 * BN, first,
                                              (int)((g.N - n2 * BN < BN ? g.N - n2 * BN : BN) / 32), lane);
                }
                if (n_blk != bias_blk) {   // tiles run m-fastest: the bias columns change only every m_tiles tiles
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int c = first + kPer * j;
                        if (c < chunks) sm.bias_s[epi][j * 32 + lane] = g.bias[colb + c * 32 + lane];
                    }
                    bias_blk = n_blk;
                    __syncwarp();
                }
                // This warp's chunks are c = first + kPer * j.  Residuals ping-pong between two register sets, each
                // requested one chunk ahead; no register set is ever copied (a copy behind the load waits for it and
                // exposes one memory round trip per chunk -- measured on the first version of this loop).
                const int n_mine = chunks > first ? (chunks - first + kPer - 1) / kPer : 0;
                const uint32_t tacc = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
                conv_ctx_tile(cx, g, row0, colb, lane);
                uint4 res_a[4] = {}, res_b[4] = {};
                if (g.residual && n_mine > 0) conv_residual_fetch(g, cx, first, res_a);
#pragma unroll 1
                for (int j = 0; j < n_mine; j += 2) {
                    const int c0 = first + kPer * j, c1 = c0 + kPer;
                    if (g.residual && j + 1 < n_mine) conv_residual_fetch(g, cx, c1, res_b);
                    {
                        uint32_t v[32];
                        tmem_ld32(tacc + c0 * 32, v);
                        conv_epilogue_chunk(g, cx, c0, j, v, res_a);
                    }
                    if (g.residual && j + 2 < n_mine) conv_residual_fetch(g, cx, c1 + kPer, res_a);
                    if (j + 1 < n_mine) {
                        uint32_t v[32];
                        tmem_ld32(tacc + c1 * 32, v);
                        conv_epilogue_chunk(g, cx, c1, j + 1, v, res_b);
                    }
                }
            } else {
#pragma unroll 1
                for (int c = (warp - 2) >> 2; c < BN / 32; c += epi_warps(EPI) / 4) {
                    const int64_t col0 = n_blk * BN + c * 32;
                    if (col0 >= g.N) break;
                    const int valid = (int)(g.N - col0 < 32 ? g.N - col0 : 32);
                    uint32_t v[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN + c * 32;
                    tmem_ld32(taddr, v);
                    epilogue_chunk<EPI>(g, best, best_col, row, col0, valid, v, lane, emit, taddr, osc);
                }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.tmem_empty[acc]);
            if (EPI == EPI_ROWMAX && row < g.M)
                atomicMax(&g.rowmax_key[row], vsc::float_to_key(best * osc));
            if (EPI == EPI_ROWARGMAX && row < g.M && best_col >= 0)
                atomicMax(&g.rowbest[row], ((unsigned long long)vsc::float_to_key(best * osc) << 32) |
                                               (unsigned long long)(0xFFFFFFFFu - (uint32_t)best_col));
        }
        if (EPI == EPI_EMIT) {   // retire the warp's last block, publish its hit count
            emit_pad(g, emit, lane);
            if (lane == 0 && emit.counted) atomicAdd(&g.counters[1], emit.counted);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols));
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeIm2colFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const int *, const int *, cuuint32_t, cuuint32_t, const cuuint32_t *,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

// NHWC bf16 activation tensor [n][h][w][c] for a k x k / pad p / stride s convolution: one load = 128 output pixels x
// 64 channels of one filter tap, 128-byte swizzle (the same shared-memory image as a tiled 128 x 64 box).
// Bounding box corners (cuTensorMapEncodeIm2col): lower = -pad, upper = pad - (filter - 1).
int make_im2col_map(CUtensorMap *map, const void *ptr, int n, int h, int w, int c, int stride, int ksize, int pad) {
    static EncodeIm2colFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeIm2colFn>(p);
    }
    if (!fn) { vsc::set_error("cuTensorMapEncodeIm2col entry point not available"); return VSC_ERR_CUDA; }
    cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
    cuuint64_t strides[3] = {(cuuint64_t)c * 2, (cuuint64_t)w * c * 2, (cuuint64_t)h * w * c * 2};
    int lower[2] = {-pad, -pad}, upper[2] = {pad - (ksize - 1), pad - (ksize - 1)};
    cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void *>(ptr), dims, strides, lower, upper,
                    (cuuint32_t)BK, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vsc::set_error("cuTensorMapEncodeIm2col failed (%d)", (int)r); return VSC_ERR_CUDA; }
    return VSC_OK;
}

template <int EPI, int BN, int ALOAD = A_TILED>
int launch(const void *a, const void *b, const GemmArgs &g, cudaStream_t stream, const CUtensorMap *map_a = nullptr) {
    if (g.M <= 0 || g.N <= 0) return VSC_OK;
    if (g.K <= 0 || g.K % BK != 0) { vsc::set_error("gemm: K=%d must be a positive multiple of %d", g.K, BK); return VSC_ERR_INVALID; }
    if ((reinterpret_cast<uintptr_t>(a) & 15) || (reinterpret_cast<uintptr_t>(b) & 15)) {
        vsc::set_error("gemm: operands must be 16-byte aligned"); return VSC_ERR_INVALID;
    }
    CUtensorMap ma, mb;
    int rc = VSC_OK;
    if (map_a) ma = *map_a;
    else rc = make_map(&ma, a, g.M, g.K, BM, g.lda);
    if (rc != VSC_OK) return rc;
    rc = make_map(&mb, b, g.N, g.K, BN, g.ldb);
    if (rc != VSC_OK) return rc;
    const size_t smem = sizeof(SharedStorage<BN>) + 1024;
    VSC_CUDA_CHECK(cudaFuncSetAttribute(gemm_kernel<EPI, BN, ALOAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int dev = 0, sms = 148;
    VSC_CUDA_CHECK(cudaGetDevice(&dev));
    VSC_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int64_t tiles = ((g.M + BM - 1) / BM) * ((g.N + BN - 1) / BN);
    const int grid = (int)(tiles < sms ? tiles : sms);
    gemm_kernel<EPI, BN, ALOAD><<<grid, cta_threads(EPI), smem, stream>>>(ma, mb, g);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

__global__ void fill_u32(uint32_t *p, int64_t n, uint32_t v) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void unpack_rowbest(const unsigned long long *packed, float *score, int64_t *col, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long p = packed[i];
    score[i] = p ? vsc::key_to_float((uint32_t)(p >> 32)) : -INFINITY;
    col[i] = p ? (int64_t)(0xFFFFFFFFu - (uint32_t)(p & 0xFFFFFFFFull)) : -1;
}
__global__ void keys_to_float(const uint32_t *k, float *out, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = vsc::key_to_float(k[i]);
}

}  // namespace

static void apply_format(GemmArgs &g, const vsc_gemm_format *fmt) {
    if (!fmt) return;
    g.ab_f16 = fmt->ab_f16; g.lda = fmt->lda; g.ldb = fmt->ldb; g.out_scale = fmt->d_out_scale;
}

extern "C" int vsc_gemm_store(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, float *d_c,
                              int64_t ldc, const vsc_gemm_format *fmt, vsc_stream_t stream) {
    GemmArgs g = {};
    apply_format(g, fmt);
    g.M = m; g.N = n; g.K = k; g.c = d_c; g.ldc = ldc;
    return launch<EPI_STORE, 256>(d_a, d_b, g, static_cast<cudaStream_t>(stream));
}

extern "C" int vsc_gemm_rowmax(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, float *d_rowmax,
                               const vsc_gemm_format *fmt, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m <= 0) return VSC_OK;
    GemmArgs g = {};
    apply_format(g, fmt);
    g.M = m; g.N = n; g.K = k;
    // the fp32 output buffer doubles as the key buffer: keys first, converted in place afterwards
    g.rowmax_key = reinterpret_cast<uint32_t *>(d_rowmax);
    fill_u32<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(g.rowmax_key, m, 0x007FFFFFu);  // key of -inf
    vsc::count_launch();
    int rc = launch<EPI_ROWMAX, 256>(d_a, d_b, g, stream);
    if (rc != VSC_OK) return rc;
    keys_to_float<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(g.rowmax_key, d_rowmax, m);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

// N tile for a convolution GEMM.  (Measured: 128-column tiles for the 18x18 / 9x9 layers, which have only 1.1 - 2.2
// rounds of 128x256 tiles on 148 SMs, were SLOWER -- 57 -> 80 us at 2304 -> 256 channels: the A tile is re-read per N
// tile and the per-tile pipeline fill dominates.  So: the widest tile that N allows.)
static int conv_bn(int64_t, int64_t n) { return n <= 64 ? 64 : n <= 128 ? 128 : 256; }

// Convolution as GEMM: out[m][n] (bf16, row stride ldc) = relu?(A[m][:] . W[n][:] + bias[n] + residual[m][n]).
extern "C" int vsc_gemm_conv(const void *d_a, int64_t m, const void *d_w, int64_t n, int32_t k, const float *d_bias,
                             const void *d_residual, int32_t relu, void *d_out_bf16, int64_t ldc,
                             vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n % 32 != 0 || ldc % 8 != 0 || !d_bias) { vsc::set_error("vsc_gemm_conv: need n %% 32 == 0, ldc %% 8 == 0, a bias"); return VSC_ERR_INVALID; }
    GemmArgs g = {};
    g.M = m; g.N = n; g.K = k; g.bias = d_bias; g.residual = static_cast<const __nv_bfloat16 *>(d_residual);
    g.relu = relu; g.out_bf16 = static_cast<__nv_bfloat16 *>(d_out_bf16); g.ldc = ldc;
    const int bn = conv_bn(m, n);
    if (bn == 64) return launch<EPI_CONV, 64>(d_a, d_w, g, stream);
    if (bn == 128) return launch<EPI_CONV, 128>(d_a, d_w, g, stream);
    return launch<EPI_CONV, 256>(d_a, d_w, g, stream);
}

// k x k convolution (3x3 / pad 1 or 1x1 / pad 0, stride 1|2) straight from the NHWC bf16 activation tensor (implicit
// GEMM: no patch matrix and, for the strided 1x1 downsample, no subsampled copy in memory).
static int conv_implicit(const char *who, const void *d_in, int n, int h, int w, int c, int ksize, int pad, int stride,
                         const void *d_w, int cout, const float *d_bias, const void *d_residual, int relu,
                         void *d_out_bf16, cudaStream_t stream) {
    if (c % BK != 0 || cout % 32 != 0 || (stride != 1 && stride != 2) || !d_bias) {
        vsc::set_error("%s: need c %% 64 == 0, cout %% 32 == 0, stride in {1,2}, a bias", who); return VSC_ERR_INVALID;
    }
    if (n <= 0) return VSC_OK;
    const int ho = (h + 2 * pad - ksize) / stride + 1, wo = (w + 2 * pad - ksize) / stride + 1;
    if ((reinterpret_cast<uintptr_t>(d_in) & 15) != 0) { vsc::set_error("%s: input must be 16-byte aligned", who); return VSC_ERR_INVALID; }
    CUtensorMap ma;
    int rc = make_im2col_map(&ma, d_in, n, h, w, c, stride, ksize, pad);
    if (rc != VSC_OK) return rc;
    GemmArgs g = {};
    g.M = (int64_t)n * ho * wo; g.N = cout; g.K = ksize * ksize * c; g.bias = d_bias;
    g.residual = static_cast<const __nv_bfloat16 *>(d_residual); g.relu = relu;
    g.out_bf16 = static_cast<__nv_bfloat16 *>(d_out_bf16); g.ldc = cout;
    g.conv_ho = ho; g.conv_wo = wo; g.conv_stride = stride; g.conv_cblocks = c / BK; g.conv_ksize = ksize; g.conv_pad = pad;
    const int bn = conv_bn(g.M, cout);
    if (bn == 64) return launch<EPI_CONV, 64, A_IM2COL>(d_in, d_w, g, stream, &ma);
    if (bn == 128) return launch<EPI_CONV, 128, A_IM2COL>(d_in, d_w, g, stream, &ma);
    return launch<EPI_CONV, 256, A_IM2COL>(d_in, d_w, g, stream, &ma);
}

// 3x3 / pad 1: weights [cout][9*c] with K index = (ky*3 + kx)*c + channel; out[(n*ho + oy)*wo + ox][cout].
extern "C" int vsc_conv3x3(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t stride, const void *d_w,
                           int32_t cout, const float *d_bias, const void *d_residual, int32_t relu, void *d_out_bf16,
                           vsc_stream_t stream) {
    return conv_implicit("vsc_conv3x3", d_in, n, h, w, c, 3, 1, stride, d_w, cout, d_bias, d_residual, relu, d_out_bf16,
                         static_cast<cudaStream_t>(stream));
}
// 1x1 / stride 1|2 (the ResNet downsample branch): weights [cout][c]; stride 2 reads every second pixel of every
// second row through the tensor map's traversal strides.
extern "C" int vsc_conv1x1(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t stride, const void *d_w,
                           int32_t cout, const float *d_bias, const void *d_residual, int32_t relu, void *d_out_bf16,
                           vsc_stream_t stream) {
    return conv_implicit("vsc_conv1x1", d_in, n, h, w, c, 1, 0, stride, d_w, cout, d_bias, d_residual, relu, d_out_bf16,
                         static_cast<cudaStream_t>(stream));
}

// The stem GEMM (vsc_conv_stem in sscd_ops.cu): A rows are 64-element windows that start every 16 elements of the
// space-to-depth image, i.e. a 2D view whose row stride (32 B) is smaller than its row length (128 B).
extern "C" int vsc_gemm_stem(const void *d_s2d, int64_t pixels, int64_t row_shift, const void *d_w, const float *d_bias,
                             void *d_out_bf16, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    EncodeTiledFn fn = encode_fn();
    if (!fn) { vsc::set_error("cuTensorMapEncodeTiled entry point not available"); return VSC_ERR_CUDA; }
    CUtensorMap ma;
    cuuint64_t dims[2] = {64, (cuuint64_t)pixels};
    cuuint64_t strides[1] = {32};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(d_s2d), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vsc::set_error("cuTensorMapEncodeTiled (overlapping stem rows) failed (%d)", (int)r); return VSC_ERR_CUDA; }
    GemmArgs g = {};
    g.M = pixels; g.N = 64; g.K = 256; g.bias = d_bias; g.relu = 1;
    g.out_bf16 = static_cast<__nv_bfloat16 *>(d_out_bf16); g.ldc = 64; g.a_row_shift = row_shift;
    return launch<EPI_CONV, 64, A_SHIFT>(d_s2d, d_w, g, stream, &ma);
}

// fp32 out[m][n] = A . W^T + bias[n]  (the SSCD projection head)
extern "C" int vsc_gemm_linear(const void *d_a, int64_t m, const void *d_w, int64_t n, int32_t k, const float *d_bias,
                               float *d_out, int64_t ldc, vsc_stream_t stream) {
    GemmArgs g = {};
    g.M = m; g.N = n; g.K = k; g.c = d_out; g.ldc = ldc; g.bias = d_bias;
    return launch<EPI_STORE, 256>(d_a, d_w, g, static_cast<cudaStream_t>(stream));
}

extern "C" int vsc_gemm_rowargmax(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, float *d_score,
                                  int64_t *d_col, unsigned long long *d_scratch, const vsc_gemm_format *fmt,
                                  vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (m <= 0) return VSC_OK;
    GemmArgs g = {};
    apply_format(g, fmt);
    g.M = m; g.N = n; g.K = k; g.rowbest = d_scratch;
    VSC_CUDA_CHECK(cudaMemsetAsync(d_scratch, 0, sizeof(unsigned long long) * (size_t)m, stream));
    int rc = launch<EPI_ROWARGMAX, 256>(d_a, d_b, g, stream);
    if (rc != VSC_OK) return rc;
    unpack_rowbest<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(d_scratch, d_score, d_col, m);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

namespace vsc {
// range-search launch with the thresholds in device memory (search.cu): d_thr = {count threshold, emit threshold}
int launch_emit_device(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, const float *d_a_norm,
                       const float *d_b_norm, int32_t metric_l2, const float *d_thr, int64_t row_offset, float *d_score,
                       int32_t *d_row, int32_t *d_col, uint64_t capacity, unsigned long long *d_counters,
                       const vsc_gemm_format *fmt, cudaStream_t stream, const float *d_margin) {
    GemmArgs g = {};
    apply_format(g, fmt);
    g.M = m; g.N = n; g.K = k;
    g.thr_margin = metric_l2 ? nullptr : d_margin;
    g.a_norm = d_a_norm; g.b_norm = d_b_norm; g.metric_l2 = metric_l2; g.thr_ptr = d_thr;
    g.row_offset = row_offset; g.out_score = d_score; g.out_row = d_row; g.out_col = d_col;
    g.capacity = capacity; g.counters = d_counters;
    return launch<EPI_EMIT, 256>(d_a, d_b, g, stream);
}
}  // namespace vsc

// Inner-product range search with one threshold PER QUERY ROW: appends every (score, i, j) with score > d_row_thr[i].
// The second pass of the filtered row maximum (vsc_rowmax_rescore below / index.py max_similarity): thresholds are the
// single-product row maxima minus the error margin, so the few columns that can hold the true maximum come out.
extern "C" int vsc_gemm_emit_rows(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, const float *d_row_thr,
                                  float *d_score, int32_t *d_row, int32_t *d_col, uint64_t capacity,
                                  unsigned long long *d_counters, const vsc_gemm_format *fmt, vsc_stream_t stream) {
    if (!d_row_thr) { vsc::set_error("vsc_gemm_emit_rows: null thresholds"); return VSC_ERR_INVALID; }
    GemmArgs g = {};
    apply_format(g, fmt);
    g.M = m; g.N = n; g.K = k;
    g.row_thr = d_row_thr;
    g.out_score = d_score; g.out_row = d_row; g.out_col = d_col; g.capacity = capacity; g.counters = d_counters;
    return launch<EPI_EMIT, 256>(d_a, d_b, g, static_cast<cudaStream_t>(stream));
}

extern "C" int vsc_gemm_emit(const void *d_a, int64_t m, const void *d_b, int64_t n, int32_t k, const float *d_a_norm,
                             const float *d_b_norm, int32_t metric_l2, float count_thr, float emit_thr,
                             int64_t row_offset, int64_t col_offset, float *d_score, int32_t *d_row, int32_t *d_col,
                             uint64_t capacity, unsigned long long *d_counters, const vsc_gemm_format *fmt,
                             vsc_stream_t stream) {
    if (metric_l2 && (!d_a_norm || !d_b_norm)) { vsc::set_error("vsc_gemm_emit: L2 metric needs squared norms"); return VSC_ERR_INVALID; }
    GemmArgs g = {};
    apply_format(g, fmt);
    g.M = m; g.N = n; g.K = k;
    g.a_norm = d_a_norm; g.b_norm = d_b_norm; g.metric_l2 = metric_l2;
    g.count_thr = count_thr; g.emit_thr = emit_thr; g.row_offset = row_offset; g.col_offset = col_offset;
    g.out_score = d_score; g.out_row = d_row; g.out_col = d_col; g.capacity = capacity; g.counters = d_counters;
    return launch<EPI_EMIT, 256>(d_a, d_b, g, static_cast<cudaStream_t>(stream));
}
