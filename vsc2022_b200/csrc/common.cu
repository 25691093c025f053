// Error string, ABI version and launch counter of libvsc_b200.so.
#include <atomic>
#include <stdarg.h>

#include "common.cuh"

namespace vsc {
static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof g_error, fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Stream-ordered scratch comes from the device's default memory pool.  By default the pool hands
// unused memory back to the driver at every synchronisation, which turns each call into a fresh
// cudaMalloc; keep it cached instead.
void keep_pool_cached() {
    static thread_local int done_for = -1;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev == done_for) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    done_for = dev;
}
}  // namespace vsc

extern "C" const char *vsc_last_error(void) { return vsc::g_error; }
extern "C" int vsc_abi_version(void) { return 1; }
extern "C" int64_t vsc_launch_count(void) { return vsc::g_launches.load(); }

// Host -> device copies of many row ranges of one array in ONE call (localization.py's block-wise descriptor upload: a
// chunk of candidate pairs touches hundreds of scattered reference videos; a Python-level copy per range costs more
// host time than the copies take).  ranges = [first_row, n_rows] pairs; every range is one cudaMemcpyAsync of
// n_rows * row_bytes bytes at the same offset of both arrays.  Pinned host memory makes them asynchronous.
extern "C" int vsc_upload_rows(void *d_dst, const void *h_src, int64_t row_bytes, const int64_t *ranges, int32_t n_ranges,
                               vsc_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (n_ranges <= 0) return VSC_OK;
    if (!d_dst || !h_src || !ranges || row_bytes <= 0) { vsc::set_error("vsc_upload_rows: bad arguments"); return VSC_ERR_INVALID; }
    for (int32_t i = 0; i < n_ranges; ++i) {
        const int64_t first = ranges[2 * i], rows = ranges[2 * i + 1];
        if (first < 0 || rows < 0) { vsc::set_error("vsc_upload_rows: negative range"); return VSC_ERR_INVALID; }
        if (rows == 0) continue;
        VSC_CUDA_CHECK(cudaMemcpyAsync(static_cast<char *>(d_dst) + first * row_bytes,
                                       static_cast<const char *>(h_src) + first * row_bytes, (size_t)(rows * row_bytes),
                                       cudaMemcpyHostToDevice, stream));
    }
    return VSC_OK;
}
