// Shared device helpers for libctan_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#define CTAN_SM_COUNT 148

#define CTAN_CHECK_LAUNCH()                              \
    do {                                                 \
        cudaError_t _e = cudaGetLastError();             \
        if (_e != cudaSuccess) return (int)_e;           \
    } while (0)

// Errors of our own (negative so they never collide with cudaError_t).
#define CTAN_ERR_BAD_ARG (-1)
#define CTAN_ERR_UNSUPPORTED (-2)

static inline int ctan_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }

// grid sized as a multiple of the SM count, capped (grid-stride loops inside the kernels)
static inline int ctan_grid(long long work_items, int per_block, int max_blocks_per_sm = 8) {
    long long g = (work_items + per_block - 1) / per_block;
    long long cap = (long long)CTAN_SM_COUNT * max_blocks_per_sm;
    if (g < 1) g = 1;
    if (g > cap) g = cap;
    return (int)g;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// size known on the host (dev == nullptr) or produced on the device by an earlier kernel
__device__ __forceinline__ int dev_or(const int* dev, int host) { return dev ? *dev : host; }
