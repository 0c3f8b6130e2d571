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

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// The step is a chain of ~14 dependent 3-20 us kernels, so launch latency and CTA ramp-up between them are a
// visible share of it.  Kernels on that chain start with ctan_pdl_begin(): `launch_dependents` lets the NEXT kernel
// of the stream be scheduled as soon as every CTA of this one is resident, and `wait` blocks until the PREVIOUS
// kernel has completed and its writes are visible -- so nothing before the wait may touch global memory.  The host
// side launches them with CTAN_LAUNCH (programmatic stream serialization; captured into CUDA graphs as programmatic
// edges).  A kernel launched the plain way simply falls through both instructions.  CTAN_PDL=0 disables it.
__device__ __forceinline__ void ctan_pdl_begin() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
// host-side switch (ctan_set_pdl): the engines turn PDL off for the stretches of the step DAG where parallel
// branches compete for the SMs (early-scheduled dependents would sit on their slots while they wait)
extern thread_local int g_ctan_pdl;      // per host thread: two engines on two threads do not see each other's choice
static inline bool ctan_pdl_enabled() {
    static const bool on = [] { const char* e = getenv("CTAN_PDL"); return !(e && e[0] == '0'); }();
    return on && g_ctan_pdl != 0;
}
template <class... KArgs, class... Args>
static inline void ctan_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = ctan_pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);       // errors surface through cudaGetLastError()
}
#define CTAN_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ctan_launch_pdl(kernel, dim3(grid), dim3(block), smem, stream, __VA_ARGS__)

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
