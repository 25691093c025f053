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

// horizontal pass: tmp[n][y][x - x0][c], x in [x0, x0 + ow): only the columns the crop keeps.
// One CTA per kRowsPerCta consecutive input rows.  The weights and windows of the CTA's output columns are staged in shared
// memory once; every row's source span is brought in with 16-byte loads (coalesced; head / tail bytes that do not fill an aligned
// 16-byte unit inside the tensor are loaded one by one) and the output pixels are computed from shared memory.  The first
// version read its 3 x taps source bytes with single-byte global loads: 244 instructions per output pixel, load/store-unit
// bound at 1.2 TB/s (profiles/r02_new_kernels_summary.md).
constexpr int kRowsPerCta = 8;
__global__ void __launch_bounds__(512) resize_rows_kernel(const uint8_t *__restrict__ in, long long n_rows, int w, int x0, int ow,
                                                          const int32_t *__restrict__ bounds, const int32_t *__restrict__ kk,
                                                          int ksize, int taps, int span_lo, int span_px, uint8_t *__restrict__ tmp) {
    extern __shared__ __align__(16) unsigned char rz_smem[];
    int32_t *kw = reinterpret_cast<int32_t *>(rz_smem);                  // [ow][ksize] weights
    int32_t *kb = kw + (size_t)ow * ksize;                               // [ow][2] window start (relative to span_lo), count
    // the row's source span (16-byte units) behind the tables; an OFFSET into the shared array keeps the accesses in the
    // shared address space (a pointer rebuilt from an integer made them generic loads)
    uint8_t *row_s = rz_smem + ((((size_t)ow * ksize + 2 * (size_t)ow) * 4 + 15) & ~(size_t)15);
    for (int i = threadIdx.x; i < ow * ksize; i += blockDim.x) kw[i] = kk[(size_t)x0 * ksize + i];
    for (int i = threadIdx.x; i < ow; i += blockDim.x) {
        kb[2 * i] = bounds[2 * (x0 + i)] - span_lo;
        kb[2 * i + 1] = bounds[2 * (x0 + i) + 1];
    }
    const long long row0 = (long long)blockIdx.x * kRowsPerCta;
    const size_t total_bytes = (size_t)n_rows * w * 3;
    const int span_bytes = span_px * 3;
    for (int rr = 0; rr < kRowsPerCta; ++rr) {
        const long long row = row0 + rr;
        if (row >= n_rows) break;                       // uniform over the CTA
        const size_t lo = ((size_t)row * w + span_lo) * 3;               // byte offset of the span inside the tensor
        const size_t addr = reinterpret_cast<size_t>(in) + lo;
        const int head = (int)(addr & 15);                               // the span starts `head` bytes into an aligned unit
        const int units = (head + span_bytes + 15) >> 4;
        __syncthreads();                                // the previous row's readers are done (and kw / kb are written)
        for (int u = threadIdx.x; u < units; u += blockDim.x) {
            const long long off = (long long)lo - head + (long long)u * 16;   // tensor byte offset of this aligned unit
            if (off >= 0 && (size_t)off + 16 <= total_bytes) {
                *reinterpret_cast<uint4 *>(row_s + u * 16) = *reinterpret_cast<const uint4 *>(in + off);
            } else {
                for (int j = 0; j < 16; ++j) {
                    const long long o = off + j;
                    row_s[u * 16 + j] = (o >= 0 && (size_t)o < total_bytes) ? in[o] : (uint8_t)0;
                }
            }
        }
        __syncthreads();
        const uint8_t *src_row = row_s + head;
        for (int xo = threadIdx.x; xo < ow; xo += blockDim.x) {
            const int xmin = kb[2 * xo];
            const int32_t *k = kw + (size_t)xo * ksize;
            const uint8_t *src = src_row + xmin * 3;
            int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
            // every window runs over `taps` taps, the longest window of the table: Pillow zero-fills the weights behind a
            // window's own count (Resample.c: "remaining values should stay empty"), so the trip count is uniform over the
            // warp; the bytes read behind the last window lie in the slack of the shared buffer and meet a zero weight
#pragma unroll 4
            for (int x = 0; x < taps; ++x) {
                const int c = k[x];
                s0 += src[3 * x] * c; s1 += src[3 * x + 1] * c; s2 += src[3 * x + 2] * c;
            }
            uint8_t *dst = tmp + ((size_t)row * ow + xo) * 3;
            dst[0] = clip8(s0); dst[1] = clip8(s1); dst[2] = clip8(s2);
        }
    }
}

// vertical pass: out[n][y - y0][x][c], y in [y0, y0 + oh).  The pass does not care about pixels: a row is ow*3 bytes and every
// byte is filtered down its own column, so a thread takes FOUR consecutive bytes with one 32-bit load per tap (VEC) when the
// row length allows it; the generic form takes one byte.
template <bool VEC>
__global__ void __launch_bounds__(256) resize_cols_kernel(const uint8_t *__restrict__ tmp, int n, int h, int row_bytes, int y0, int oh,
                                                          const int32_t *__restrict__ bounds, const int32_t *__restrict__ kk,
                                                          int ksize, uint8_t *__restrict__ out) {
    const int per_row = VEC ? row_bytes >> 2 : row_bytes;               // work items per output row
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)n * oh * per_row;
    if (idx >= total) return;
    const int item = (int)(idx % per_row);
    const long long t = idx / per_row;
    const int yo = (int)(t % oh), img = (int)(t / oh);
    const int yy = y0 + yo;
    const int ymin = bounds[2 * yy], ymax = bounds[2 * yy + 1];
    const int32_t *k = kk + (size_t)yy * ksize;
    const size_t col = VEC ? (size_t)item * 4 : (size_t)item;
    const uint8_t *src = tmp + ((size_t)img * h + ymin) * row_bytes + col;
    uint8_t *dst = out + ((size_t)img * oh + yo) * row_bytes + col;
    if (VEC) {
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0, s3 = s0;
        for (int y = 0; y < ymax; ++y) {
            const int c = k[y];
            const uint32_t v = *reinterpret_cast<const uint32_t *>(src + (size_t)y * row_bytes);
            s0 += (int)(v & 0xFFu) * c; s1 += (int)((v >> 8) & 0xFFu) * c;
            s2 += (int)((v >> 16) & 0xFFu) * c; s3 += (int)(v >> 24) * c;
        }
        *reinterpret_cast<uint32_t *>(dst) = (uint32_t)clip8(s0) | ((uint32_t)clip8(s1) << 8) | ((uint32_t)clip8(s2) << 16) |
                                             ((uint32_t)clip8(s3) << 24);
    } else {
        int s0 = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < ymax; ++y) s0 += src[(size_t)y * row_bytes] * k[y];
        dst[0] = clip8(s0);
    }
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
    // source span of the kept output columns: [first window start, last window end).  The tables live on the device; the span is
    // derived here from Pillow's formula (support = max(1, w / rw)):
    // window of output x = [trunc(center - support + 0.5), trunc(center + support + 0.5)) clipped to [0, w).
    const double scale = (double)w / rw, support = scale < 1.0 ? 1.0 : scale;
    auto lo_of = [&](int x) { int v = (int)(((x + 0.5) * scale) - support + 0.5); return v < 0 ? 0 : v; };
    auto hi_of = [&](int x) { int v = (int)(((x + 0.5) * scale) + support + 0.5); return v > w ? w : v; };
    int span_lo = lo_of(left), span_hi = hi_of(left);
    int taps = 1;
    for (int x = left; x < left + ow; ++x) {
        const int a = lo_of(x), b = hi_of(x);
        span_lo = a < span_lo ? a : span_lo; span_hi = b > span_hi ? b : span_hi;
        taps = b - a > taps ? b - a : taps;
    }
    taps = taps + 1 < xksize ? taps + 1 : xksize;      // (+1: the tables are authoritative, this formula only sizes the loop)
    span_lo = span_lo > 0 ? span_lo - 1 : 0;            // one pixel of slack on either side: the tables are authoritative
    span_hi = span_hi < w ? span_hi + 1 : w;
    const int span_px = span_hi - span_lo;
    const long long n_rows = (long long)n * h;
    const size_t smem = (size_t)ow * xksize * 4 + (size_t)ow * 8 + 16 + (((size_t)span_px * 3 + 15 + 15) & ~(size_t)15) + 16 +
                        (size_t)xksize * 3 + 16;     // slack behind the span for the zero-weight taps of the last windows
    if (smem > 200 * 1024) { vsc::set_error("vsc_resize_u8: row of %d pixels x %d taps does not fit shared memory", ow, xksize); return VSC_ERR_CAPACITY; }
    VSC_CUDA_CHECK(cudaFuncSetAttribute(resize_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int row_threads = ow >= 512 ? 512 : (ow + 31) / 32 * 32;      // one output pixel per thread and row when it fits
    resize_rows_kernel<<<(unsigned)((n_rows + kRowsPerCta - 1) / kRowsPerCta), row_threads, smem, stream>>>(
        d_in, n_rows, w, left, ow, d_xbounds, d_xk, xksize, taps, span_lo, span_px, d_tmp);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    const int row_bytes = ow * 3;
    const bool vec = (row_bytes & 3) == 0 && (reinterpret_cast<uintptr_t>(d_tmp) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 3) == 0;
    const long long t2 = (long long)n * oh * (vec ? row_bytes / 4 : row_bytes);
    if (vec) resize_cols_kernel<true><<<(unsigned)((t2 + 255) / 256), 256, 0, stream>>>(d_tmp, n, h, row_bytes, top, oh, d_ybounds, d_yk, yksize, d_out);
    else resize_cols_kernel<false><<<(unsigned)((t2 + 255) / 256), 256, 0, stream>>>(d_tmp, n, h, row_bytes, top, oh, d_ybounds, d_yk, yksize, d_out);
    VSC_CUDA_CHECK(cudaGetLastError());
    vsc::count_launch();
    return VSC_OK;
}
