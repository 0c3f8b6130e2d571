// fp32 CUDA-core tiled contraction C[m,n] = sum_k a(m,k) * b(k,n) with functor operand loaders
// (gathers, concatenations, the cos time encoding and the ReLU are fused into the tile loads) and
// functor epilogues (bias / residual / atomic split-K / scatter).
//
// The per-batch contractions of the reference configs are 10^7-10^8 flop on a few hundred rows:
// they are bound by load LATENCY, not by flops or bandwidth.  Hence: a 4-stage cp.async (LDGSTS)
// shared-memory ring so four k-tiles of loads are in flight per CTA, and two tile shapes -- 64x64
// (256 threads) and 32x32 (64 threads) -- the small one chosen when the big one would leave most of
// the 148 SMs idle.  True-fp32 FMA keeps the 1e-4 parity bar of the fp32 path.
// Sizes may live on the device (produced by the lookup kernels): CTAs past the real extent exit
// before touching memory.
#pragma once
#include "common.cuh"

namespace gemm {

// select one of four pointers without dynamic indexing of a kernel-parameter array (which would
// force the array into local memory)
template <class P>
__device__ __forceinline__ P sel4(const P (&p)[4], int s) {
    return s == 0 ? p[0] : (s == 1 ? p[1] : (s == 2 ? p[2] : p[3]));
}

constexpr int BK = 16, PAD = 4, STAGES = 4;

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Every operand functor provides
//   ptr(r,c): address of the element when it is a plain memory element (-> cp.async), else nullptr
//   val(r,c): the element's value when it has to be computed (only called when ptr() == nullptr)
// ----------------------------------------------------------------------------- A operands a(m,k)
struct ARowMajor {                 // a(m,k) = p[m*ld + k]
    static constexpr bool K_CONTIG = true;
    const float* p; int ld;
    __device__ __forceinline__ const float* ptr(int m, int k) const { return p + (size_t)m * ld + k; }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct AColMajor {                 // a(m,k) = p[k*ld + m]        (A = Y^T)
    static constexpr bool K_CONTIG = false;
    const float* p; int ld;
    __device__ __forceinline__ const float* ptr(int m, int k) const { return p + (size_t)k * ld + m; }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct AGather2 {                  // a(m,k) = k<K1 ? t1[row1(m)][k] : t2[row2(m)][k-K1];  row = map ? map[idx[m]] : idx[m]
    static constexpr bool K_CONTIG = true;
    const float* t1; int ld1; const int64_t* idx1; int K1;
    const float* t2; int ld2; const int64_t* idx2;
    const int64_t* map;
    __device__ __forceinline__ const float* ptr(int m, int k) const {
        if (k < K1) {
            int64_t r = __ldg(idx1 + m); if (map) r = __ldg(map + r);
            return t1 + (size_t)r * ld1 + k;
        }
        int64_t r = __ldg(idx2 + m); if (map) r = __ldg(map + r);
        return t2 + (size_t)r * ld2 + (k - K1);
    }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct AEdgeAttr {                 // a(e,k) = k<Fe ? msg[row(e)][k] : cos(w[k-Fe]*rel[e] + b[k-Fe])
    static constexpr bool K_CONTIG = true;
    const float* msg; int ldm; const int64_t* eidx; int Fe;
    const float* rel; const float* tw; const float* tb;
    __device__ __forceinline__ const float* ptr(int e, int k) const {
        if (k >= Fe) return nullptr;
        const int64_t r = eidx ? __ldg(eidx + e) : (int64_t)e;
        return msg + (size_t)r * ldm + k;
    }
    __device__ __forceinline__ float val(int e, int k) const {
        return cosf(fmaf(__ldg(rel + e), __ldg(tw + (k - Fe)), __ldg(tb + (k - Fe))));
    }
};

// ----------------------------------------------------------------------------- B operands b(k,n)
struct BWeight {                   // b(k,n) = W[n*ld + k]          (torch Linear weight [out,in])
    static constexpr bool K_CONTIG = true;
    const float* p; int ld;
    __device__ __forceinline__ const float* ptr(int k, int n) const { return p + (size_t)n * ld + k; }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct BWeightSeg4 {               // four [D,ld] weights stacked along n
    static constexpr bool K_CONTIG = true;
    const float* p[4]; int D; int ld;
    __device__ __forceinline__ const float* ptr(int k, int n) const {
        return sel4(p, n / D) + (size_t)(n % D) * ld + k;
    }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct BWeightCat2 {               // b(k,n) = k<K1 ? W1[n][k] : W2[n][k-K1]
    static constexpr bool K_CONTIG = true;
    const float* p1; int ld1; int K1; const float* p2; int ld2;
    __device__ __forceinline__ const float* ptr(int k, int n) const {
        return k < K1 ? p1 + (size_t)n * ld1 + k : p2 + (size_t)n * ld2 + (k - K1);
    }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct BRowMajor {                 // b(k,n) = X[k*ld + n]
    static constexpr bool K_CONTIG = false;
    const float* p; int ld;
    __device__ __forceinline__ const float* ptr(int k, int n) const { return p + (size_t)k * ld + n; }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct BRowSeg4 {                  // four [D,ld] matrices stacked along k
    static constexpr bool K_CONTIG = false;
    const float* p[4]; int D; int ld;
    __device__ __forceinline__ const float* ptr(int k, int n) const {
        return sel4(p, k / D) + (size_t)(k % D) * ld + n;
    }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct BGather2 {                  // b(k,n) = n<K1 ? t1[row1(k)][n] : t2[row2(k)][n-K1]
    static constexpr bool K_CONTIG = false;
    const float* t1; int ld1; const int64_t* idx1; int K1;
    const float* t2; int ld2; const int64_t* idx2;
    const int64_t* map;
    __device__ __forceinline__ const float* ptr(int k, int n) const {
        if (n < K1) {
            int64_t r = __ldg(idx1 + k); if (map) r = __ldg(map + r);
            return t1 + (size_t)r * ld1 + n;
        }
        int64_t r = __ldg(idx2 + k); if (map) r = __ldg(map + r);
        return t2 + (size_t)r * ld2 + (n - K1);
    }
    __device__ __forceinline__ float val(int, int) const { return 0.f; }
};
struct BEdgeAttr {                 // b(e,n) = n<Fe ? msg[row(e)][n] : cos(w*rel+b)
    static constexpr bool K_CONTIG = false;
    const float* msg; int ldm; const int64_t* eidx; int Fe;
    const float* rel; const float* tw; const float* tb;
    __device__ __forceinline__ const float* ptr(int e, int n) const {
        if (n >= Fe) return nullptr;
        const int64_t r = eidx ? __ldg(eidx + e) : (int64_t)e;
        return msg + (size_t)r * ldm + n;
    }
    __device__ __forceinline__ float val(int e, int n) const {
        return cosf(fmaf(__ldg(rel + e), __ldg(tw + (n - Fe)), __ldg(tb + (n - Fe))));
    }
};
struct BRelu {                     // b(k,n) = max(H[k*ld+n], 0)
    static constexpr bool K_CONTIG = false;
    const float* p; int ld;
    __device__ __forceinline__ const float* ptr(int, int) const { return nullptr; }
    __device__ __forceinline__ float val(int k, int n) const { return fmaxf(__ldg(p + (size_t)k * ld + n), 0.f); }
};

// ----------------------------------------------------------------------------- epilogues
struct EpiStore {                  // C = acc + bias[n] + bias2[n] + res[m][n]
    static constexpr bool HAS_ROWSUM = false;
    float* C; int ldc; const float* bias; const float* bias2; const float* res; int ldr;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        if (bias) v += __ldg(bias + n);
        if (bias2) v += __ldg(bias2 + n);
        if (res) v += res[(size_t)m * ldr + n];
        C[(size_t)m * ldc + n] = v;
    }
    __device__ __forceinline__ void rowsum(int, float) const {}
};
struct EpiStoreSeg4 {              // C = acc + bias_seg[n/D][n%D]
    static constexpr bool HAS_ROWSUM = false;
    float* C; int ldc; const float* bias[4]; int D;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        const float* b = sel4(bias, n / D);
        if (b) v += __ldg(b + (n % D));
        C[(size_t)m * ldc + n] = v;
    }
    __device__ __forceinline__ void rowsum(int, float) const {}
};
struct EpiAtomic {                 // split-K accumulate: C[m][n] += acc ; rs[m] += sum_k a(m,k)
    static constexpr bool HAS_ROWSUM = true;
    float* C; int ldc; float* rs;
    __device__ __forceinline__ void operator()(int m, int n, float v) const { atomicAdd(C + (size_t)m * ldc + n, v); }
    __device__ __forceinline__ void rowsum(int m, float v) const { if (rs) atomicAdd(rs + m, v); }
};
struct EpiAtomicSeg4 {             // rows split into 4 segments of D rows, each with its own output / bias-grad
    static constexpr bool HAS_ROWSUM = true;
    float* C[4]; float* rs[4]; int D; int ldc;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        atomicAdd(sel4(C, m / D) + (size_t)(m % D) * ldc + n, v);
    }
    __device__ __forceinline__ void rowsum(int m, float v) const {
        float* r = sel4(rs, m / D);
        if (r) atomicAdd(r + (m % D), v);
    }
};
struct EpiScatterAtomic {          // C[row(m)][n] += acc      (row = map ? map[idx[m]] : idx[m])
    static constexpr bool HAS_ROWSUM = false;
    float* C; int ldc; const int64_t* idx; const int64_t* map;
    __device__ __forceinline__ void operator()(int m, int n, float v) const {
        int64_t r = idx[m]; if (map) r = map[r];
        atomicAdd(C + (size_t)r * ldc + n, v);
    }
    __device__ __forceinline__ void rowsum(int, float) const {}
};

// ----------------------------------------------------------------------------- kernel
// BT x BT output tile, (BT/4)^2 threads, 4x4 register micro-tile per thread.
template <int BT, class AL, class BL, class EP>
__global__ void __launch_bounds__((BT / 4) * (BT / 4))
gemm_kernel(const AL A, const BL Bm, const EP epi, int M_host, const int* __restrict__ M_dev, int N, int K_host,
            const int* __restrict__ K_dev, int k_chunk) {
    constexpr int TPR = BT / 4;                 // threads per tile row/col
    constexpr int THREADS = TPR * TPR;
    constexpr int EPT = BT * BK / THREADS;      // elements per thread per operand tile
    __shared__ __align__(16) float As[STAGES][BK][BT + PAD];
    __shared__ __align__(16) float Bs[STAGES][BK][BT + PAD];
    const int M = M_dev ? *M_dev : M_host;
    const int K = K_dev ? *K_dev : K_host;
    const int m0 = blockIdx.y * BT, n0 = blockIdx.x * BT;
    const int kb = blockIdx.z * k_chunk;
    const int ke = min(K, kb + k_chunk);
    if (m0 >= M || n0 >= N || kb >= ke) return;        // block-uniform, before any barrier
    const int t = threadIdx.x, ty = t / TPR, tx = t % TPR;

    auto fetch_tile = [&](int k0, int st) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            int r, k;                                   // r: row within the tile (m or n), k: within BK
            if (AL::K_CONTIG) { k = t % BK; r = t / BK + i * (THREADS / BK); }
            else              { r = t % BT; k = t / BT + i * (THREADS / BT); }
            float* dst = &As[st][k][r];
            const int m = m0 + r, kk = k0 + k;
            if (m < M && kk < ke) {
                const float* p = A.ptr(m, kk);
                if (p) cp_async4(dst, p); else *dst = A.val(m, kk);
            } else {
                *dst = 0.f;
            }
        }
#pragma unroll
        for (int i = 0; i < EPT; ++i) {
            int r, k;
            if (BL::K_CONTIG) { k = t % BK; r = t / BK + i * (THREADS / BK); }
            else              { r = t % BT; k = t / BT + i * (THREADS / BT); }
            float* dst = &Bs[st][k][r];
            const int n = n0 + r, kk = k0 + k;
            if (n < N && kk < ke) {
                const float* p = Bm.ptr(kk, n);
                if (p) cp_async4(dst, p); else *dst = Bm.val(kk, n);
            } else {
                *dst = 0.f;
            }
        }
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float rsum[4] = {0.f, 0.f, 0.f, 0.f};

    const int n_tiles = (ke - kb + BK - 1) / BK;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < n_tiles) fetch_tile(kb + s * BK, s);
        cp_async_commit();                              // one group per stage, empty or not
    }
    for (int kt = 0; kt < n_tiles; ++kt) {
        cp_async_wait<STAGES - 2>();                    // tile kt has landed (own copies) ...
        __syncthreads();                                // ... and everyone's; tile kt-1's stage is free
        const int nxt = kt + STAGES - 1;
        if (nxt < n_tiles) fetch_tile(kb + nxt * BK, nxt % STAGES);
        cp_async_commit();
        const int st = kt % STAGES;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[st][k][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4*>(&Bs[st][k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            }
            if (EP::HAS_ROWSUM) {
#pragma unroll
                for (int i = 0; i < 4; ++i) rsum[i] += a[i];
            }
        }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) epi(m, n, acc[i][j]);
        }
        if (EP::HAS_ROWSUM && blockIdx.x == 0 && tx == 0) epi.rowsum(m, rsum[i]);
    }
}

// M_max bounds the grid when the true M lives on the device; splits > 1 needs an atomic epilogue.
// m_hint: the typical true M (<= M_max) used only to pick the tile shape.
template <class AL, class BL, class EP>
static inline int launch(const AL& A, const BL& B, const EP& epi, int M_max, const int* M_dev, int N, int K_max,
                         const int* K_dev, int splits, cudaStream_t stream) {
    if (M_max <= 0 || N <= 0 || K_max <= 0) return 0;
    if (splits < 1) splits = 1;
    int k_chunk = (K_max + splits - 1) / splits;
    k_chunk = ((k_chunk + BK - 1) / BK) * BK;
    splits = (K_max + k_chunk - 1) / k_chunk;
    const long long big_tiles = (long long)((N + 63) / 64) * ((M_max + 63) / 64) * splits;
    if (big_tiles >= 2 * CTAN_SM_COUNT) {
        dim3 grid((N + 63) / 64, (M_max + 63) / 64, splits);
        gemm_kernel<64, AL, BL, EP><<<grid, 256, 0, stream>>>(A, B, epi, M_max, M_dev, N, K_max, K_dev, k_chunk);
    } else {
        dim3 grid((N + 31) / 32, (M_max + 31) / 32, splits);
        gemm_kernel<32, AL, BL, EP><<<grid, 64, 0, stream>>>(A, B, epi, M_max, M_dev, N, K_max, K_dev, k_chunk);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// number of K splits that fills the chip for a reduce-over-rows contraction (32x32 tiles)
static inline int pick_splits(int M, int N, int K) {
    const int tiles = ((M + 31) / 32) * ((N + 31) / 32);
    int s = (2 * CTAN_SM_COUNT + tiles - 1) / tiles;
    const int max_s = (K + 4 * BK - 1) / (4 * BK);     // at least 64 rows per split
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return s;
}

}  // namespace gemm
