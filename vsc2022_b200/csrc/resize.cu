// Frame resize ahead of the SSCD stem: the PIL bilinear resize that
// torchvision.transforms.Resize applies in the reference's transforms (vsc/baseline/inference_impl.py:39-69), on
// decoded uint8 RGB frames in device memory, with the optional centre crop of RESIZE_320_CENTER folded in.
//
// Pillow (third party; the version torchvision brings, unpinned by the reference) resamples in two separable integer
// passes, libImaging/Resample.c: per output coordinate a window [xmin, xmin + xmax) of input pixels with triangle
// weights (support = max(1, in/out): antialiased when shrinking), normalised in double precision and quantised to
// 22-bit fixed point; accumulate in int32 from 2^21, shift by 22, clip to uint8 -- horizontally into a temporary uint8
// image, then vertically.  The coefficient tables are computed on the host exactly as Pillow does
// (vsc2022_b200/preprocess.py: pil_coefficients); these kernels do the integer arithmetic, so the result equals
// Pillow's bit for bit (tests/test_preprocess_gpu.py compares with PIL itself).
//
// HBM-bound byte work: every input byte is read once per pass from DRAM (neighbouring threads share taps through
// L1/L2), one thread per output pixel (3 channels), coalesced 3-byte stores.
#include "common.cuh"

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;   // Resample.c PRECISION_BITS

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= kPrecisionBits;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: tmp[n][y][x - x0][c], x in [x0, x0 + ow): only the columns the crop keeps
__global__ void __launch_bounds__(256) resize_rows_kernel(const uint8_t *__restrict__ in, int n, int h, int w, int x0, int ow,
                                                          const int32_t *__restrict__ bounds, const int32_t *__restrict__ kk,
                                                          int ksize, uint8_t *__restrict__ tmp) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * h * ow;
    if (idx >= total) return;
    const int xo = (int)(idx % ow);
    const long long row = idx / ow;          // image * h + y
    const int xx = x0 + xo;
    const int xmin = bounds[2 * xx], xmax = bounds[2 * xx + 1];
    const int32_t *k = kk + (size_t)xx * ksize;
    const uint8_t *src = in + ((size_t)row * w + xmin) * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int x = 0; x < xmax; ++x) {
        const int c = k[x];
        s0 += src[3 * x] * c; s1 += src[3 * x + 1] * c; s2 += src[3 * x + 2] * c;
    }
    uint8_t *dst = tmp + (size_t)idx * 3;
    dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

// vertical pass: out[n][y - y0][x][c], y in [y0, y0 + oh)
__global__ void __launch_bounds__(256) resize_cols_kernel(const uint8_t *__restrict__ tmp, int n, int h, int ow, int y0, int oh,
                                                          const int32_t *__restrict__ bounds, const int32_t *__restrict__ kk,
                                                          int ksize, uint8_t *__restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * oh * ow;
    if (idx >= total) return;
    const int xo = (int)(idx % ow);
    const long long t = idx / ow;
    const int yo = (int)(t % oh), img = (int)(t / oh);
    const int yy = y0 + yo;
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const int32_t *k = kk + (size_t)yy * ksize;
    const uint8_t *src = tmp + (((size_t)img * h + ymin) * ow + xo) * 3;
    const size_t pitch = (size_t)ow * 3;
    int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
    for (int y = 0; y < ymax; ++y) {
        const int c = k[y];
        s0 += src[y * pitch] * c; s1 += src[y * pitch + 1] * c; s2 += src[y * pitch + 2] * c;
    }
    uint8_t *dst = out + (size_t)idx * 3;
    dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
}

}  // namespace

extern "C" int vsc_resize_u8(const uint8_t *d_in, int32_t n, int32_t h, int32_t w, int32_t rh, int32_t rw,
                             int32_t top, int32_t left, int32_t oh, int32_t ow,
                             const int32_t *d_xbounds, const int32_t *d_xk, int32_t xksize,
                             const int32_t *d_ybounds, const int32_t *d_yk, int32_t yksize,
                             uint8_t *d_tmp, uint8_t *d_out, vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n <= 0) return VSC_OK;
    if (!d_in || !d_out || !d_tmp || !d_xbounds || !d_xk || !d_ybounds || !d_yk || h <= 0 || w <= 0 || xksize <= 0 ||
        yksize <= 0 || top < 0 || left < 0 || oh <= 0 || ow <= 0 || top + oh > rh || left + ow > rw) {
        vsc::set_error("vsc_resize_u8: bad arguments (crop %d+%d x %d+%d inside %d x %d?)", top, oh, left, ow, rh, rw);
        return VSC_ERR_INVALID;
    }
    const long long t1 = (long long)n * h * ow, t2 = (long long)n * oh * ow;
    resize_rows_kernel<<<(unsigned)((t1 + 255) / 256), 256, 0, stream>>>(d_in, n, h, w, left, ow, d_xbounds, d_xk, xksize, d_tmp);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    resize_cols_kernel<<<(unsigned)((t2 + 255) / 256), 256, 0, stream>>>(d_tmp, n, h, ow, top, oh, d_ybounds, d_yk, yksize, d_out);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
