// Data-movement and pooling kernels around the tensor-core GEMM for the SSCD ResNet-50 forward
// (vsc/baseline/inference_impl.py:210-239 runs the TorchScript model; its shape contract is documented in
// vsc/baseline/adapt_sscd_model.py:56-70: ResNet-50 trunk -> GeM pooling -> Linear(2048 -> 512), no L2 norm).
//
// Activations are NHWC bf16, so a 1x1 convolution is a GEMM on the activation tensor itself; 3x3 / 7x7
// convolutions go through an explicit im2col panel (first version; the implicit-GEMM TMA-im2col load is the
// planned replacement).  All kernels here are HBM-bound copies with 16-byte accesses.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

// ---- stem: 7x7 stride 2 pad 3 on 3 channels.  Panel row = output pixel, K index = (ky*7 + kx)*3 + c, padded 147 -> 192.
// mode 0: uint8 NHWC pixels, normalised here ((x/255 - mean)/std, inference_impl.py:39-69); padding is zero in
//         NORMALISED space, exactly like torchvision's Normalize followed by the conv's zero padding.
// mode 1: float32 NCHW tensor that is already normalised (what the reference model receives).
__global__ void __launch_bounds__(256) im2col_stem_kernel(const void *__restrict__ in, int mode, int n, int h, int w,
                                                          int ho, int wo, __nv_bfloat16 *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (pixel, ky, kx-group)
    const long long total = (long long)n * ho * wo * 64;                     // 64 slots of 3 values = 192
    if (idx >= total) return;
    const int slot = (int)(idx & 63);
    const long long pix = idx >> 6;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), img = (int)(pix / ((long long)wo * ho));
    __nv_bfloat16 v[3] = {__float2bfloat16(0.f), __float2bfloat16(0.f), __float2bfloat16(0.f)};
    if (slot < 49) {
        const int ky = slot / 7, kx = slot - ky * 7;
        const int iy = oy * 2 - 3 + ky, ix = ox * 2 - 3 + kx;
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) {
            const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float x;
                if (mode == 0) {
                    const uint8_t *p = static_cast<const uint8_t *>(in);
                    x = ((float)p[(((long long)img * h + iy) * w + ix) * 3 + c] / 255.0f - mean[c]) / stdv[c];
                } else {
                    const float *p = static_cast<const float *>(in);
                    x = p[(((long long)img * 3 + c) * h + iy) * w + ix];
                }
                v[c] = __float2bfloat16_rn(x);
            }
        }
    }
    __nv_bfloat16 *o = out + pix * 192 + slot * 3;
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
}

// ---- 3x3 pad 1, stride s: panel [n*ho*wo][9*c], K index = (ky*3 + kx)*c + ch.  One thread moves 8 channels (16 B).
__global__ void __launch_bounds__(256) im2col3x3_kernel(const uint4 *__restrict__ in, int n, int h, int w, int c8,
                                                        int stride, int ho, int wo, uint4 *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * ho * wo * 9 * c8;
    if (idx >= total) return;
    const int ch = (int)(idx % c8);
    const int tap = (int)((idx / c8) % 9);
    const long long pix = idx / ((long long)c8 * 9);
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), img = (int)(pix / ((long long)wo * ho));
    const int ky = tap / 3, kx = tap - ky * 3;
    const int iy = oy * stride - 1 + ky, ix = ox * stride - 1 + kx;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = in[(((long long)img * h + iy) * w + ix) * c8 + ch];
    out[idx] = v;
}

// ---- every second pixel (input of the stride-2 1x1 downsample convolution)
__global__ void __launch_bounds__(256) subsample2_kernel(const uint4 *__restrict__ in, int n, int h, int w, int c8,
                                                         int ho, int wo, uint4 *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * ho * wo * c8;
    if (idx >= total) return;
    const int ch = (int)(idx % c8);
    const long long pix = idx / c8;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), img = (int)(pix / ((long long)wo * ho));
    out[idx] = in[(((long long)img * h + oy * 2) * w + ox * 2) * c8 + ch];
}

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
}

// ---- max pool 3x3 stride 2 pad 1 (padding never wins: the window always holds a real pixel)
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const uint4 *__restrict__ in, int n, int h, int w, int c8,
                                                           int ho, int wo, uint4 *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * ho * wo * c8;
    if (idx >= total) return;
    const int ch = (int)(idx % c8);
    const long long pix = idx / c8;
    const int ox = (int)(pix % wo), oy = (int)((pix / wo) % ho), img = (int)(pix / ((long long)wo * ho));
    bool any = false;
    uint4 best = make_uint4(0, 0, 0, 0);
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= h) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            if (ix < 0 || ix >= w) continue;
            const uint4 v = in[(((long long)img * h + iy) * w + ix) * c8 + ch];
            if (!any) { best = v; any = true; }
            else best = make_uint4(bf16x2_max(best.x, v.x), bf16x2_max(best.y, v.y), bf16x2_max(best.z, v.z), bf16x2_max(best.w, v.w));
        }
    }
    out[idx] = best;
}

// ---- GeM pooling: (mean_hw clamp(x, eps)^p)^(1/p) per (image, channel), fp32 math, bf16 out (GEMM operand).
// One warp per (image, 64-channel group): lanes stride over pixels, 2 channels per lane.
__global__ void __launch_bounds__(256) gem_pool_kernel(const __nv_bfloat16 *__restrict__ in, int n, int hw, int c, float p,
                                                       float eps, __nv_bfloat16 *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (image, channel)
    if (idx >= (long long)n * c) return;
    const int ch = (int)(idx % c);
    const long long img = idx / c;
    float acc = 0.0f;
    for (int i = 0; i < hw; ++i) {   // consecutive threads read consecutive channels: coalesced
        const float x = fmaxf(__bfloat162float(in[(img * hw + i) * c + ch]), eps);
        acc += p == 3.0f ? x * x * x : powf(x, p);
    }
    out[idx] = __float2bfloat16_rn(powf(acc / (float)hw, 1.0f / p));
}

inline unsigned blocks(long long total) { return (unsigned)((total + 255) / 256); }

}  // namespace

extern "C" int vsc_im2col_stem(const void *d_in, int32_t mode, int32_t n, int32_t h, int32_t w, void *d_out,
                               vsc_stream_t stream) {
    const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1;
    if (n <= 0) return VSC_OK;
    im2col_stem_kernel<<<blocks((long long)n * ho * wo * 64), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        d_in, mode, n, h, w, ho, wo, static_cast<__nv_bfloat16 *>(d_out));
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
extern "C" int vsc_im2col3x3(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t stride, void *d_out,
                             vsc_stream_t stream) {
    if (c % 8 != 0 || (stride != 1 && stride != 2)) { vsc::set_error("vsc_im2col3x3: c %% 8 == 0 and stride in {1,2}"); return VSC_ERR_INVALID; }
    const int ho = (h + 2 - 3) / stride + 1, wo = (w + 2 - 3) / stride + 1;
    if (n <= 0) return VSC_OK;
    im2col3x3_kernel<<<blocks((long long)n * ho * wo * 9 * (c / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(d_in), n, h, w, c / 8, stride, ho, wo, static_cast<uint4 *>(d_out));
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
extern "C" int vsc_subsample2(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, void *d_out, vsc_stream_t stream) {
    if (c % 8 != 0) { vsc::set_error("vsc_subsample2: c %% 8 == 0"); return VSC_ERR_INVALID; }
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    if (n <= 0) return VSC_OK;
    subsample2_kernel<<<blocks((long long)n * ho * wo * (c / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(d_in), n, h, w, c / 8, ho, wo, static_cast<uint4 *>(d_out));
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
extern "C" int vsc_maxpool3x3s2(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, void *d_out, vsc_stream_t stream) {
    if (c % 8 != 0) { vsc::set_error("vsc_maxpool3x3s2: c %% 8 == 0"); return VSC_ERR_INVALID; }
    const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    if (n <= 0) return VSC_OK;
    maxpool3x3s2_kernel<<<blocks((long long)n * ho * wo * (c / 8)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(d_in), n, h, w, c / 8, ho, wo, static_cast<uint4 *>(d_out));
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
extern "C" int vsc_gem_pool(const void *d_in, int32_t n, int32_t hw, int32_t c, float p, float eps, void *d_out,
                            vsc_stream_t stream) {
    if (n <= 0) return VSC_OK;
    gem_pool_kernel<<<blocks((long long)n * c), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16 *>(d_in), n, hw, c, p, eps, static_cast<__nv_bfloat16 *>(d_out));
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
