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
