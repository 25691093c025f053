// Operand preparation for the descriptor GEMM: fp32 descriptors -> K-major bf16 panels.
//
// Tensor cores multiply bf16; the reference multiplies fp32 (FAISS sgemm).  Descriptors whose values
// are bf16-representable lose nothing (mode HI).  For arbitrary fp32 descriptors the split
// x = hi + lo (hi = bf16(x), lo = bf16(x - hi)) recovers fp32-class products with three partial
// GEMMs  hi.hi + hi.lo + lo.hi,  laid out as ONE GEMM with K' = 3K:
//     A' = [hi | hi | lo]      B' = [hi | lo | hi]
// (the dropped lo.lo term is < 2^-16 relative per product).  HBM-bound elementwise kernel.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

enum { MODE_HI = 0, MODE_SPLIT_A = 1, MODE_SPLIT_B = 2 };

// One thread per four consecutive k of a row (16-byte loads when the rows allow it, 8-byte stores); the panel rows are
// contiguous, so a warp writes 256 contiguous bytes per panel part.
template <bool VEC>
__global__ void __launch_bounds__(256) prepare_kernel(const float *__restrict__ x, int64_t n, int d, int64_t ld,
                                                      int kpad, int mode, __nv_bfloat16 *__restrict__ out,
                                                      int *__restrict__ lo_flag) {
    const int quads = kpad >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = idx / quads;
    const int k = (int)(idx - row * quads) * 4;
    bool any_lo = false;
    if (row < n) {
        float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const float *src = x + row * ld + k;
        if (VEC && k + 3 < d) {
            const float4 f = *reinterpret_cast<const float4 *>(src);
            v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (k + j < d) v[j] = src[j];
        }
        __nv_bfloat16 hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            hi[j] = __float2bfloat16_rn(v[j]);
            lo[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
            any_lo = any_lo || __bfloat162float(lo[j]) != 0.0f;
        }
        const uint2 hi2 = *reinterpret_cast<const uint2 *>(hi), lo2 = *reinterpret_cast<const uint2 *>(lo);
        if (mode == MODE_HI) {
            *reinterpret_cast<uint2 *>(out + row * kpad + k) = hi2;
        } else {
            __nv_bfloat16 *o = out + row * (int64_t)(3 * kpad) + k;
            *reinterpret_cast<uint2 *>(o) = hi2;
            *reinterpret_cast<uint2 *>(o + kpad) = mode == MODE_SPLIT_A ? hi2 : lo2;
            *reinterpret_cast<uint2 *>(o + 2 * kpad) = mode == MODE_SPLIT_A ? lo2 : hi2;
        }
    }
    if (lo_flag && __any_sync(vsc::kFullMask, any_lo) && (threadIdx.x & 31) == 0) atomicOr(lo_flag, 1);
}

// squared L2 norm per row (float32, sequential-in-k per warp lane then warp reduce): for the L2 metric
__global__ void __launch_bounds__(256) sqnorm_kernel(const float *__restrict__ x, int64_t n, int d, int64_t ld,
                                                     float *__restrict__ out) {
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    float acc = 0.0f;
    for (int k = lane; k < d; k += 32) { const float v = x[row * ld + k]; acc = fmaf(v, v, acc); }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(vsc::kFullMask, acc, s);
    if (lane == 0) out[row] = acc;
}

}  // namespace

extern "C" int vsc_prepare_operand(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t kpad, int32_t mode,
                                   void *d_out_bf16, int32_t *d_lo_flag, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0) return VSC_OK;
    if (kpad < d || kpad % 64 != 0 || mode < 0 || mode > 2) {
        vsc::set_error("vsc_prepare_operand: kpad=%d must be a multiple of 64 and >= d=%d; mode in 0..2", kpad, d);
        return VSC_ERR_INVALID;
    }
    const int64_t total = n * (kpad / 4);
    const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15u) == 0;
    if (vec)
        prepare_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            d_x, n, d, ld, kpad, mode, static_cast<__nv_bfloat16 *>(d_out_bf16), d_lo_flag);
    else
        prepare_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            d_x, n, d, ld, kpad, mode, static_cast<__nv_bfloat16 *>(d_out_bf16), d_lo_flag);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

extern "C" int vsc_row_sqnorm(const float *d_x, int64_t n, int32_t d, int64_t ld, float *d_out, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0) return VSC_OK;
    sqnorm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(d_x, n, d, ld, d_out);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
