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

// one thread per (row, k) element of the padded panel; rows are contiguous so writes coalesce
__global__ void __launch_bounds__(256) prepare_kernel(const float *__restrict__ x, int64_t n, int d, int64_t ld,
                                                      int kpad, int mode, __nv_bfloat16 *__restrict__ out,
                                                      int *__restrict__ lo_flag, float *__restrict__ sq_norm) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = idx / kpad;
    const int k = (int)(idx - row * kpad);
    if (row >= n) return;
    const float v = k < d ? x[row * ld + k] : 0.0f;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    if (mode == MODE_HI) {
        out[row * kpad + k] = hi;
    } else {
        __nv_bfloat16 *o = out + row * (int64_t)(3 * kpad);
        o[k] = hi;
        o[kpad + k] = mode == MODE_SPLIT_A ? hi : lo;
        o[2 * kpad + k] = mode == MODE_SPLIT_A ? lo : hi;
    }
    if (lo_flag && __bfloat162float(lo) != 0.0f) atomicOr(lo_flag, 1);
    (void)sq_norm;
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
    const int64_t total = n * kpad;
    prepare_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        d_x, n, d, ld, kpad, mode, static_cast<__nv_bfloat16 *>(d_out_bf16), d_lo_flag, nullptr);
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
