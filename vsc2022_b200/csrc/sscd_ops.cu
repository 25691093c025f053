// Data-movement and pooling kernels around the tensor-core GEMM for the SSCD ResNet-50 forward
// (vsc/baseline/inference_impl.py:210-239 runs the TorchScript model; its shape contract is documented in
// vsc/baseline/adapt_sscd_model.py:56-70: ResNet-50 trunk -> GeM pooling -> Linear(2048 -> 512), no L2 norm).
//
// Activations are NHWC bf16, so a 1x1 convolution is a GEMM on the activation tensor itself; 3x3 convolutions are
// implicit GEMMs (TMA im2col loads inside gemm_tc.cu) and the 7x7 stem reads the space-to-depth image written here.
// What lives in this file: that space-to-depth / normalisation kernel, the pooling kernels, and two explicit
// data-movement kernels (im2col3x3, subsample2) that only the tests use, as the independent statement the implicit
// paths are compared with.  All kernels here are HBM-bound copies with 16-byte accesses.
#include <cuda_bf16.h>

#include "common.cuh"

namespace {

// ---- stem: 7x7 stride 2 pad 3 on 3 channels, as a 4x4 stride-1 convolution over a 2x2 space-to-depth image.
// stem_s2d_kernel: S[n][Y][X][(dy*2+dx)*3 + c] = normalised pixel (2Y+dy-3, 2X+dx-3, c), zero outside the frame
//         (zero in NORMALISED space, exactly like torchvision's Normalize followed by the conv's padding) and in the
//         4 padding channels; Y < ho+3, X < wo+3, 16 bf16 = 32 bytes per cell.
//         mode 0: uint8 NHWC pixels, normalised here ((x/255 - mean)/std, inference_impl.py:39-69);
//         mode 1: float32 NCHW tensor that is already normalised (what the reference model receives).
// The 4 x 16 channels of filter row ky2 at output pixel (oy, ox) are the 64 CONTIGUOUS elements that start at cell
// (oy+ky2, ox): with output rows numbered m = (n*(ho+3) + oy)*(wo+3) + ox (the padded grid), k-block ky2 of the GEMM
// is the 64-element window starting at cell m + ky2*(wo+3) -- a plain 2D TMA load from an overlapping-row view of S
// (vsc_gemm_stem, gemm_tc.cu).  No patch matrix exists; the ~4 % of rows with oy >= ho or ox >= wo are computed
// from in-bounds data and ignored by the max pool.  K index = ky2*64 + kx2*16 + (dy*2+dx)*3 + c <-> filter tap
// (2*ky2+dy, 2*kx2+dx); taps with row / column 7 and the padding channels carry zero weights.  K = 256.
// (First version: an explicit [n*ho*wo][192] patch panel, 1.3 ms per 128 frames of 288x288 -- 8x its HBM time.)
template <int MODE>
__global__ void __launch_bounds__(256) stem_s2d_kernel(const void *__restrict__ in, int n, int h, int w, int yd, int xd,
                                                       uint32_t total, uint4 *__restrict__ out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;   // one (image, Y, X) cell
    if (idx >= total) return;
    const uint32_t t = idx / (uint32_t)xd;
    const int X = (int)(idx - t * (uint32_t)xd);
    const int img = (int)(t / (uint32_t)yd), Y = (int)(t - (uint32_t)img * (uint32_t)yd);
    const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
    uint16_t v[16];
#pragma unroll
    for (int i = 12; i < 16; ++i) v[i] = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
            const int iy = 2 * Y + dy - 3, ix = 2 * X + dx - 3;
            const bool inside = img < n && iy >= 0 && iy < h && ix >= 0 && ix < w;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float x = 0.0f;
                if (inside) {
                    if (MODE == 0)
                        x = ((float)static_cast<const uint8_t *>(in)[(((size_t)img * h + iy) * w + ix) * 3 + c] / 255.0f - mean[c]) / stdv[c];
                    else
                        x = static_cast<const float *>(in)[(((size_t)img * 3 + c) * h + iy) * w + ix];
                }
                const __nv_bfloat16 b = __float2bfloat16_rn(x);
                v[(dy * 2 + dx) * 3 + c] = inside ? *reinterpret_cast<const uint16_t *>(&b) : (uint16_t)0;
            }
        }
    }
    uint4 *o = out + (size_t)idx * 2;
#pragma unroll
    for (int q = 0; q < 2; ++q)
        o[q] = make_uint4(v[q * 8] | ((uint32_t)v[q * 8 + 1] << 16), v[q * 8 + 2] | ((uint32_t)v[q * 8 + 3] << 16),
                          v[q * 8 + 4] | ((uint32_t)v[q * 8 + 5] << 16), v[q * 8 + 6] | ((uint32_t)v[q * 8 + 7] << 16));
}

// ---- 3x3 pad 1, stride s: panel [n*ho*wo][9*c], K index = (ky*3 + kx)*c + ch.  One thread moves 8 channels (16 B)
// of all nine taps of one output pixel: the pixel is decoded once (32-bit arithmetic), the nine loads are issued
// before the nine stores, and consecutive threads cover consecutive channels so every access is a full 16-byte
// piece of a contiguous c*2-byte run.
__global__ void __launch_bounds__(256) im2col3x3_kernel(const uint4 *__restrict__ in, int n, int h, int w, int c8,
                                                        int stride, int ho, int wo, uint32_t total,
                                                        uint4 *__restrict__ out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;   // (pixel, channel group)
    if (idx >= total) return;
    const uint32_t pix = idx / (uint32_t)c8, ch = idx - pix * (uint32_t)c8;
    const uint32_t t = pix / (uint32_t)wo;
    const int ox = (int)(pix - t * (uint32_t)wo);
    const int img = (int)(t / (uint32_t)ho), oy = (int)(t - (uint32_t)img * (uint32_t)ho);
    const uint4 *src = in + ((size_t)img * h * w) * c8 + ch;
    uint4 v[9];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * stride - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * stride - 1 + kx;
            v[ky * 3 + kx] = (iy >= 0 && iy < h && ix >= 0 && ix < w) ? src[((size_t)iy * w + ix) * c8] : make_uint4(0, 0, 0, 0);
        }
    }
    uint4 *dst = out + (size_t)pix * 9 * c8 + ch;
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) dst[(size_t)tap * c8] = v[tap];
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

// ---- max pool 3x3 stride 2 pad 1 (padding never wins: the window always holds a real pixel).  The input may sit in
// a larger grid (row_pitch pixels per row, img_rows rows per image): the stem writes its padded grid.
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(const uint4 *__restrict__ in, int n, int h, int w, int c8,
                                                           int row_pitch, int img_rows, int ho, int wo, uint32_t total,
                                                           uint4 *__restrict__ out) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const uint32_t pix = idx / (uint32_t)c8, ch = idx - pix * (uint32_t)c8;
    const uint32_t t = pix / (uint32_t)wo;
    const int ox = (int)(pix - t * (uint32_t)wo);
    const int img = (int)(t / (uint32_t)ho), oy = (int)(t - (uint32_t)img * (uint32_t)ho);
    bool any = false;
    uint4 best = make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = oy * 2 - 1 + ky;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = ox * 2 - 1 + kx;
            if (ix < 0 || ix >= w) continue;
            const uint4 v = in[(((size_t)img * img_rows + iy) * row_pitch + ix) * c8 + ch];
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

extern "C" int vsc_gemm_stem(const void *d_s2d, int64_t pixels, int64_t row_shift, const void *d_w, const float *d_bias,
                             void *d_out_bf16, vsc_stream_t stream);

// Stem convolution + folded BN + ReLU.  Output: bf16 [n][ho+3][wo+3][64] (the padded grid described above;
// rows oy < ho and columns ox < wo are the convolution result).
extern "C" int vsc_conv_stem(const void *d_in, int32_t mode, int32_t n, int32_t h, int32_t w, const void *d_w,
                             const float *d_bias, void *d_out_bf16, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int ho = (h + 6 - 7) / 2 + 1, wo = (w + 6 - 7) / 2 + 1, yd = ho + 3, xd = wo + 3;
    if (n <= 0) return VSC_OK;
    if (mode != 0 && mode != 1) { vsc::set_error("vsc_conv_stem: mode must be 0 (uint8 NHWC) or 1 (float32 NCHW)"); return VSC_ERR_INVALID; }
    const long long cells = (long long)n * yd * xd, padded = cells + 4;   // the last windows read 3 cells past the end
    if (padded >= (1ll << 31)) { vsc::set_error("vsc_conv_stem: batch of %d %dx%d frames too large", n, h, w); return VSC_ERR_CAPACITY; }
    vsc::keep_pool_cached();
    uint4 *s2d = nullptr;
    VSC_CUDA_CHECK(cudaMallocAsync(&s2d, (size_t)padded * 32, stream));
    if (mode == 0) stem_s2d_kernel<0><<<blocks(padded), 256, 0, stream>>>(d_in, n, h, w, yd, xd, (uint32_t)padded, s2d);
    else stem_s2d_kernel<1><<<blocks(padded), 256, 0, stream>>>(d_in, n, h, w, yd, xd, (uint32_t)padded, s2d);
    cudaError_t e = cudaGetLastError();
    vsc::count_launch();
    int rc = VSC_OK;
    if (e == cudaSuccess) rc = vsc_gemm_stem(s2d, cells, xd, d_w, d_bias, d_out_bf16, stream_);
    cudaFreeAsync(s2d, stream);
    VSC_CUDA_CHECK(e);
    return rc;
}
extern "C" int vsc_im2col3x3(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t stride, void *d_out,
                             vsc_stream_t stream) {
    if (c % 8 != 0 || (stride != 1 && stride != 2)) { vsc::set_error("vsc_im2col3x3: c %% 8 == 0 and stride in {1,2}"); return VSC_ERR_INVALID; }
    const int ho = (h + 2 - 3) / stride + 1, wo = (w + 2 - 3) / stride + 1;
    if (n <= 0) return VSC_OK;
    const long long total = (long long)n * ho * wo * (c / 8);
    if (total >= (1ll << 31)) { vsc::set_error("vsc_im2col3x3: %lld work items exceed 2^31; use a smaller batch", total); return VSC_ERR_CAPACITY; }
    im2col3x3_kernel<<<blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(d_in), n, h, w, c / 8, stride, ho, wo, (uint32_t)total, static_cast<uint4 *>(d_out));
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
extern "C" int vsc_maxpool3x3s2(const void *d_in, int32_t n, int32_t h, int32_t w, int32_t c, int32_t row_pitch,
                                int32_t img_rows, void *d_out, vsc_stream_t stream) {
    if (c % 8 != 0 || row_pitch < w || img_rows < h) { vsc::set_error("vsc_maxpool3x3s2: c %% 8 == 0, row_pitch >= w, img_rows >= h"); return VSC_ERR_INVALID; }
    const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    if (n <= 0) return VSC_OK;
    const long long total = (long long)n * ho * wo * (c / 8);
    if (total >= (1ll << 31)) { vsc::set_error("vsc_maxpool3x3s2: %lld work items exceed 2^31; use a smaller batch", total); return VSC_ERR_CAPACITY; }
    maxpool3x3s2_kernel<<<blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const uint4 *>(d_in), n, h, w, c / 8, row_pitch, img_rows, ho, wo, (uint32_t)total, static_cast<uint4 *>(d_out));
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
