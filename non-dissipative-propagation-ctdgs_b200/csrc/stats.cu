// compute_stats on the device: mean / population std of the inter-event times
// `t - last_seen[node]` over BOTH endpoints of every training event, both endpoints of an event seeing the table
// as it was before the event, first occurrence measured from init_time (utils.py:46-92: a Python loop over all
// events with three dicts).  Here: one stable radix sort of (node << 32 | event) keys (CUB, the toolkit's own
// header library -- a one-off statistic, not part of the per-batch path), then "previous occurrence" is the
// sorted neighbour; sums are reduced in a fixed order (per-block partials, then one block) in fp64.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {

constexpr int RED_BLOCKS = 1024, RED_THREADS = 256;

__global__ void make_keys_kernel(const int64_t* __restrict__ src, const int64_t* __restrict__ dst, long long n,
                                 unsigned long long* __restrict__ keys) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x) {
        keys[2 * k] = ((unsigned long long)src[k] << 32) | (unsigned long long)k;
        keys[2 * k + 1] = ((unsigned long long)dst[k] << 32) | (unsigned long long)k;
    }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x < 32) {
        r = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r;                                             // valid in thread 0
}

// pass = 0: diff[i] and per-block sum of diffs; pass = 1: per-block sum of (diff - mean)^2
__global__ void diff_kernel(const unsigned long long* __restrict__ keys, const int64_t* __restrict__ t, long long m,
                            long long init_time, double* __restrict__ diff, double* __restrict__ partial, int pass,
                            const double* __restrict__ stats) {
    __shared__ double red[32];
    const double mean = pass ? stats[0] : 0.0;
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < m; i += (long long)gridDim.x * blockDim.x) {
        if (pass == 0) {
            const unsigned long long key = keys[i];
            long long j = i - 1;
            if (j >= 0 && keys[j] == key) --j;            // self-loop: both endpoints see the state before the event
            long long prev = init_time;
            if (j >= 0 && (keys[j] >> 32) == (key >> 32)) prev = t[keys[j] & 0xffffffffull];
            const double d = (double)(t[key & 0xffffffffull] - prev);
            diff[i] = d;
            acc += d;
        } else {
            const double d = diff[i] - mean;
            acc += d * d;
        }
    }
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// stats[0] = mean (pass 0) ; stats[1] = std (pass 1): fixed-order reduction of the per-block partials
__global__ void finish_kernel(const double* __restrict__ partial, int n_part, long long m, double* __restrict__ stats,
                              int pass) {
    __shared__ double red[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n_part; i += blockDim.x) acc += partial[i];
    const double s = block_sum(acc, red);
    if (threadIdx.x == 0) stats[pass] = pass ? sqrt(s / (double)m) : s / (double)m;
}

size_t sort_bytes(long long m) {
    size_t b = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, b, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, m);
    return b;
}

size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace

// bytes of workspace ctan_delta_t_stats needs for n events
static int ws_bytes_of(long long n, unsigned long long* bytes);
extern "C" int ctan_delta_t_stats_ws_bytes(long long n, unsigned long long* bytes_host, cudaStream_t stream) {
    (void)stream;                                          // host-side query; the argument keeps the calling convention
    return ws_bytes_of(n, bytes_host);
}
static int ws_bytes_of(long long n, unsigned long long* bytes) {
    if (n < 0 || !bytes) return CTAN_ERR_BAD_ARG;
    const long long m = 2 * n;
    *bytes = 2 * align256((size_t)m * 8) + align256((size_t)m * 8) + align256(RED_BLOCKS * 8) + align256(sort_bytes(m)) + 256;
    return 0;
}

// stats[0] = mean_delta_t, stats[1] = std_delta_t (fp64, device).  Node ids < 2^31, n < 2^32.
extern "C" int ctan_delta_t_stats(const int64_t* src, const int64_t* dst, const int64_t* t, long long n,
                                  long long init_time, void* ws, size_t ws_bytes, double* stats, cudaStream_t stream) {
    if (n <= 0 || n >= (1LL << 32)) return CTAN_ERR_BAD_ARG;
    unsigned long long need = 0;
    ws_bytes_of(n, &need);
    if (ws_bytes < need || ((uintptr_t)ws & 255)) return CTAN_ERR_BAD_ARG;
    const long long m = 2 * n;
    uint8_t* p = (uint8_t*)ws;
    unsigned long long* keys_in = (unsigned long long*)p;   p += align256((size_t)m * 8);
    unsigned long long* keys = (unsigned long long*)p;      p += align256((size_t)m * 8);
    double* diff = (double*)p;                              p += align256((size_t)m * 8);
    double* partial = (double*)p;                           p += align256(RED_BLOCKS * 8);
    size_t sb = sort_bytes(m);
    make_keys_kernel<<<ctan_grid(n, 256), 256, 0, stream>>>(src, dst, n, keys_in);
    CTAN_CHECK_LAUNCH();
    cudaError_t e = cub::DeviceRadixSort::SortKeys(p, sb, keys_in, keys, m, 0, 64, stream);
    if (e != cudaSuccess) return (int)e;
    for (int pass = 0; pass < 2; ++pass) {
        diff_kernel<<<RED_BLOCKS, RED_THREADS, 0, stream>>>(keys, t, m, init_time, diff, partial, pass, stats);
        CTAN_CHECK_LAUNCH();
        finish_kernel<<<1, 256, 0, stream>>>(partial, RED_BLOCKS, m, stats, pass);
        CTAN_CHECK_LAUNCH();
    }
    return 0;
}
