// Operand preparation for the descriptor GEMMs: fp32 descriptors -> K-major 16-bit panels.
//
// Tensor cores multiply 16-bit values; the reference multiplies fp32 (FAISS sgemm, np.matmul).  The product path is
// prepare_f16_kernel (vsc_prepare_operand_f16 / _more): x * 2^e is split into hi = fp16(x 2^e) and
// lo = fp16((x 2^e - hi) 2^11), 22 significant bits per value, laid out so that ONE GEMM over K' = 3K adds the three
// partial products hi.lo + lo.hi + hi.hi (include/vsc_b200.h, vsc_gemm_format); every product is exact in the fp32
// accumulator, the dropped lo.lo term is < 2^-22 relative.  Values that fit the hi part leave the lo flag clear and the
// last third of the panel alone is exact (grid / test data).
// prepare_kernel (vsc_prepare_operand) is the older bf16 form ([hi | hi | lo] / [hi | lo | hi], 16 significant bits); it
// stays for bf16 consumers (tests of the GEMM core).  HBM-bound elementwise kernels.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

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


// ---- fp16 split panels (see include/vsc_b200.h, vsc_gemm_format).  Pass 1: max |x| (float bits are monotone for
// non-negative values); pass 2 derives the power-of-two scale from it on the device and writes the panels.
__global__ void __launch_bounds__(256) absmax_kernel(const float *__restrict__ x, int64_t n, int d, int64_t ld,
                                                     unsigned int *__restrict__ out, const int32_t *__restrict__ rows) {
    unsigned int m = 0;
    const int64_t total = n * d;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t vrow = i / d;
        const int64_t row = rows ? rows[vrow] : vrow;     // optional row list: only these rows of x take part
        const float v = fabsf(x[row * ld + (i - vrow * d)]);
        if (v < INFINITY) m = max(m, __float_as_uint(v));     // NaN / inf do not take part
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) m = max(m, __shfl_xor_sync(vsc::kFullMask, m, s));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// scale exponent e: max|x| * 2^e in [2^11, 2^12)
__device__ __forceinline__ int scale_exponent(unsigned int absmax_bits) {
    if (absmax_bits == 0) return 0;
    const int ex = (int)((absmax_bits >> 23) & 0xFF) - 126;   // max|x| = m * 2^ex, m in [0.5, 1) (denormal inputs: ex = -126)
    int e = 12 - ex;
    return e > 120 ? 120 : (e < -120 ? -120 : e);
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((unsigned int)(e + 127) << 23); }

template <bool VEC>
__global__ void __launch_bounds__(256) prepare_f16_kernel(const float *__restrict__ x, int64_t n, int d, int64_t ld,
                                                          int kpad, int side, __half *__restrict__ out,
                                                          const unsigned int *__restrict__ absmax,
                                                          float *__restrict__ inv_scale, int *__restrict__ lo_flag,
                                                          const int32_t *__restrict__ rows) {
    const int e = scale_exponent(*absmax);
    const float s = pow2f(e);
    if (blockIdx.x == 0 && threadIdx.x == 0) *inv_scale = pow2f(-e);
    const int quads = kpad >> 2;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t vrow = idx / quads;
    const int k = (int)(idx - vrow * quads) * 4;
    bool any_lo = false, too_big = false;
    const int64_t row = (rows && vrow < n) ? rows[vrow] : vrow;   // optional row list (rows of x and of the panel alike)
    if (vrow < n) {
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
        __half hi[4], lo[4], hs[4], ls[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float xs = v[j] * s;                                   // exact: power of two
            too_big = too_big || (fabsf(xs) > 65504.0f && fabsf(v[j]) < INFINITY);   // only with a scale from earlier rows
            hi[j] = __float2half_rn(xs);
            const float r = xs - __half2float(hi[j]);                    // exact in fp32
            lo[j] = __float2half_rn(r * 2048.0f);
            hs[j] = __float2half_rn(__half2float(hi[j]) * (1.0f / 2048.0f));
            ls[j] = __float2half_rn(__half2float(lo[j]) * (1.0f / 2048.0f));
            any_lo = any_lo || __half2float(lo[j]) != 0.0f;
        }
        const uint2 hi2 = *reinterpret_cast<const uint2 *>(hi), lo2 = *reinterpret_cast<const uint2 *>(lo);
        const uint2 hs2 = *reinterpret_cast<const uint2 *>(hs), ls2 = *reinterpret_cast<const uint2 *>(ls);
        // the two small cross products come first, hi.hi last: the accumulator of the tensor core truncates, and
        // the error of an addition scales with the partial sum it is added to
        __half *o = out + row * (int64_t)(3 * kpad) + k;
        *reinterpret_cast<uint2 *>(o) = side == 0 ? hs2 : lo2;
        *reinterpret_cast<uint2 *>(o + kpad) = side == 0 ? ls2 : hi2;
        *reinterpret_cast<uint2 *>(o + 2 * kpad) = hi2;
    }
    if (lo_flag && __any_sync(vsc::kFullMask, any_lo) && (threadIdx.x & 31) == 0) atomicOr(lo_flag, 1);
    if (lo_flag && __any_sync(vsc::kFullMask, too_big) && (threadIdx.x & 31) == 0) atomicOr(lo_flag, 2);
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

static int prepare_f16(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t kpad, int32_t side,
                       void *d_out_f16, float *d_inv_scale, int32_t *d_lo_flag, uint32_t *d_scratch, bool keep_scale,
                       cudaStream_t stream, const int32_t *d_rows = nullptr) {
    if (kpad < d || kpad % 64 != 0 || side < 0 || side > 1 || !d_inv_scale || !d_scratch) {
        vsc::set_error("vsc_prepare_operand_f16: kpad=%d must be a multiple of 64 and >= d=%d; side in 0..1", kpad, d);
        return VSC_ERR_INVALID;
    }
    if (!keep_scale) VSC_CUDA_CHECK(cudaMemsetAsync(d_scratch, 0, sizeof(uint32_t), stream));
    if (n <= 0) {   // the scale of an empty operand is 1
        if (keep_scale) return VSC_OK;
        const float one = 1.0f;
        VSC_CUDA_CHECK(cudaMemcpyAsync(d_inv_scale, &one, sizeof(float), cudaMemcpyHostToDevice, stream));
        return VSC_OK;
    }
    if (!keep_scale) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t elems = n * d;
        const int64_t want = (elems + 255) / 256;
        absmax_kernel<<<(unsigned)(want < sms * 16 ? want : sms * 16), 256, 0, stream>>>(d_x, n, d, ld, d_scratch, d_rows);
        vsc::count_launch();
    }
    const int64_t total = n * (kpad / 4);
    const bool vec = (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(d_x) & 15u) == 0;
    if (vec)
        prepare_f16_kernel<true><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            d_x, n, d, ld, kpad, side, static_cast<__half *>(d_out_f16), d_scratch, d_inv_scale, d_lo_flag, d_rows);
    else
        prepare_f16_kernel<false><<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
            d_x, n, d, ld, kpad, side, static_cast<__half *>(d_out_f16), d_scratch, d_inv_scale, d_lo_flag, d_rows);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}

extern "C" int vsc_prepare_operand_f16(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t kpad, int32_t side,
                                       void *d_out_f16, float *d_inv_scale, int32_t *d_lo_flag, uint32_t *d_scratch,
                                       vsc_stream_t stream_) {
    return prepare_f16(d_x, n, d, ld, kpad, side, d_out_f16, d_inv_scale, d_lo_flag, d_scratch, false,
                       static_cast<cudaStream_t>(stream_));
}

// More rows of an operand whose scale is already fixed (*d_scratch as an earlier vsc_prepare_operand_f16 call on the
// same operand left it): no max|x| pass.  Bit 1 of *d_lo_flag is set when a value does not fit the fp16 range under
// that scale (the caller then prepares the operand again as a whole).
extern "C" int vsc_prepare_operand_f16_more(const float *d_x, int64_t n, int32_t d, int64_t ld, int32_t kpad,
                                            int32_t side, void *d_out_f16, float *d_inv_scale, int32_t *d_lo_flag,
                                            uint32_t *d_scratch, vsc_stream_t stream_) {
    return prepare_f16(d_x, n, d, ld, kpad, side, d_out_f16, d_inv_scale, d_lo_flag, d_scratch, true,
                       static_cast<cudaStream_t>(stream_));
}

// The same for a LIST of rows: d_x / d_out_f16 are the whole matrix and the whole panel, d_rows[0..n) the (absolute) rows to
// convert -- one launch for the hundreds of scattered videos a chunk of candidate pairs brings in.  keep_scale != 0: the
// scale of an earlier call on this operand is kept (vsc_prepare_operand_f16_more), else it is chosen from these rows.
extern "C" int vsc_prepare_operand_f16_rows(const float *d_x, const int32_t *d_rows, int64_t n, int32_t d, int64_t ld,
                                            int32_t kpad, int32_t side, void *d_out_f16, float *d_inv_scale,
                                            int32_t *d_lo_flag, uint32_t *d_scratch, int32_t keep_scale,
                                            vsc_stream_t stream_) {
    if (n > 0 && !d_rows) { vsc::set_error("vsc_prepare_operand_f16_rows: null row list"); return VSC_ERR_INVALID; }
    return prepare_f16(d_x, n, d, ld, kpad, side, d_out_f16, d_inv_scale, d_lo_flag, d_scratch, keep_scale != 0,
                       static_cast<cudaStream_t>(stream_), d_rows);
}
