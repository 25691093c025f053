// Shared host/device helpers for libvsc_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "vsc_b200.h"

namespace vsc {

void set_error(const char *fmt, ...);
void count_launch(int n = 1);
void keep_pool_cached();

#define VSC_CUDA_CHECK(expr)                                                          \
    do {                                                                              \
        cudaError_t err__ = (expr);                                                   \
        if (err__ != cudaSuccess) {                                                   \
            ::vsc::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), \
                             __FILE__, __LINE__);                                     \
            return VSC_ERR_CUDA;                                                      \
        }                                                                             \
    } while (0)

constexpr unsigned kFullMask = 0xffffffffu;

// Order-preserving float -> uint32 map.  -0.0 is folded onto +0.0 first so that
// key equality == float equality (numpy's sort treats them as equal).  Never 0
// for a non-NaN input, so 0 can mean "empty".
__device__ __forceinline__ uint32_t float_to_key(float v) {
    v += 0.0f;
    uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t k) {
    uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

}  // namespace vsc
