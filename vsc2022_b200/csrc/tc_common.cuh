// tcgen05 / TMEM / TMA building blocks shared by the tensor-core kernels (gemm_tc.cu, pair_gemm.cu), sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace vsc {
namespace tc {

constexpr int BM = 128, BK = 64;   // M tile and K block of every tensor-core kernel here
constexpr int UMMA_K = 16;

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int x, int y, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
// TMA im2col load: 128 consecutive output pixels (W fastest, then H, then image; padding is zero-filled by the
// hardware) x 64 channels of filter tap (off_w, off_h), starting at base pixel (w, h, n) of the bounding box.
__device__ __forceinline__ void tma_load_im2col(void *dst, const CUtensorMap *map, int c, int w, int h, int n,
                                                uint16_t off_w, uint16_t off_h, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%2, %3, %4, %5}], [%6], {%7, %8};"
        ::"r"(smem_u32(dst)), "l"(map), "r"(c), "r"(w), "r"(h), "r"(n), "r"(smem_u32(bar)), "h"(off_w), "h"(off_h)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint64_t *bar) {  // arrives on `bar` when all prior MMAs retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp receives row (lane base + t), columns c..c+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory matrix descriptor, K-major operand, SWIZZLE_128B: rows of 128 B (64 bf16), groups of
// 8 rows form one 1024-byte swizzle atom (stride byte offset 1024); leading byte offset unused.
__device__ __forceinline__ uint64_t umma_desc_k_major_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // [0,14)  start address >> 4
    d |= (uint64_t)1 << 16;                                // [16,30) leading byte offset >> 4 (ignored)
    d |= (uint64_t)(1024 >> 4) << 32;                      // [32,46) stride byte offset >> 4
    d |= (uint64_t)1 << 46;                                // [46,48) descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                                // [61,64) layout: SWIZZLE_128B
    return d;
}
// Instruction descriptor, kind::f16: D fp32, A/B bf16 (format 1) or fp16 (format 0), both K-major, shape M x N.
__host__ __device__ constexpr uint32_t umma_idesc_f16(int m, int n, bool bf16) {
    return (1u << 4)                                         // D format: fp32
           | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10)   // A, B format
           | ((uint32_t)(n >> 3) << 17)                      // N / 8
           | ((uint32_t)(m >> 4) << 24);                     // M / 16
}
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) { return umma_idesc_f16(m, n, true); }

// ---------------------------------------------------------------- host side: tensor maps
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2D tensor of 16-bit elements (bf16 or fp16: the copy does not care) [rows][k], K contiguous, row stride `ld`
// elements (0: = k); box = box_rows x 64 elements, 128-byte swizzle.
inline int make_map(CUtensorMap *map, const void *ptr, int64_t rows, int k, int box_rows, int64_t ld = 0) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { vsc::set_error("cuTensorMapEncodeTiled entry point not available"); return VSC_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)(ld > 0 ? ld : k) * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { vsc::set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return VSC_ERR_CUDA; }
    return VSC_OK;
}

}  // namespace tc
}  // namespace vsc
