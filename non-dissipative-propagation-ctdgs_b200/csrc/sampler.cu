// Device-side negative-destination sampler (second half).
//
// Replaces negative_sampler.py:27-40 (`NegativeSampler.sample`): for every source of the batch draw a destination
// uniformly from the set of known destination nodes and, when link-existence checking is on, redraw (up to `nattemps`
// more times) while (src, dst) is an edge seen at construction time.  The reference does this with a numpy
// generator and a python set per element of the batch; here it is one thread per element, a counter-based generator
// (no state to carry: draw = hash(seed, call, element, attempt)) and a binary search in the sorted array of seen
// (src << 32 | dst) keys.  The numpy PCG64 stream itself is not reproduced: callers that need bit-equal negatives on
// both implementations pre-draw them; what is kept is the distribution (uniform over the unseen
// destinations of that source, by rejection), the attempt cap, and eval-mode determinism (same seed -> same draws on
// every call, negative_sampler.py:28).
#include "common.cuh"

__device__ __forceinline__ uint64_t mix64(uint64_t z) {          // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ bool seen_edge(const int64_t* __restrict__ keys, int64_t n, int64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const int64_t v = keys[mid];
        if (v == key) return true;
        if (v < key) lo = mid + 1; else hi = mid;
    }
    return false;
}

__global__ void neg_sample_kernel(const int64_t* __restrict__ dst_nodes, int64_t n_dst,
                                  const int64_t* __restrict__ seen, int64_t n_seen, const int64_t* __restrict__ src,
                                  int n, uint64_t seed, uint64_t call, int max_redraws, int64_t* __restrict__ out,
                                  int* __restrict__ n_failed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t s = src[i];
    const uint64_t base = mix64(mix64(seed) ^ (call * 0xD1B54A32D192ED03ull)) ^ ((uint64_t)i * 0x8CB92BA72F3D8DD7ull);
    int64_t d = 0;
    int j = 0;
    for (;; ++j) {
        const uint64_t r = mix64(base + (uint64_t)j);
        d = dst_nodes[(int64_t)__umul64hi(r, (uint64_t)n_dst)];    // unbiased to 2^-64 * n_dst
        if (!seen || j > max_redraws || !seen_edge(seen, n_seen, (s << 32) | d)) break;
    }
    out[i] = d;
    if (seen && j > max_redraws && n_failed) atomicAdd(n_failed, 1);      // (the reference prints a warning here)
}

// dst_nodes: sorted unique destination ids [n_dst]; seen: sorted (src << 32 | dst) keys [n_seen] or NULL (no
// link-existence check); out [n].  `call` distinguishes successive calls of one sampler (0 for every eval-mode call).
extern "C" int ctan_neg_sample(const int64_t* dst_nodes, int64_t n_dst, const int64_t* seen, int64_t n_seen,
                               const int64_t* src, int n, uint64_t seed, uint64_t call, int max_redraws, int64_t* out,
                               int* n_failed, cudaStream_t stream) {
    if (n <= 0) return 0;
    if (n_dst <= 0 || !dst_nodes || !src || !out) return CTAN_ERR_BAD_ARG;
    neg_sample_kernel<<<ctan_div_up(n, 128), 128, 0, stream>>>(dst_nodes, n_dst, seen, n_seen, src, n, seed, call,
                                                              max_redraws, out, n_failed);
    CTAN_CHECK_LAUNCH();
    return 0;
}
