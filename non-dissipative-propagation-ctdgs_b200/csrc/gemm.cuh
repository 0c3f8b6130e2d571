// fp32 CUDA-core tiled contraction C[m,n] = sum_k a(m,k) * b(k,n) with functor operand loaders
// (gathers, concatenations, the cos time encoding and the ReLU are fused into the tile loads) and
// functor epilogues (bias / residual / atomic split-K / scatter).
//
// The per-batch contractions of the reference configs are 10^7-10^8 flop on a few hundred rows:
// they are bound by load LATENCY and by how many SMs get a CTA, not by flops or bandwidth.  Hence:
//   * every thread moves one 128-bit vector of A and one of B per k-tile (along the operand's
//     contiguous dimension), fetched one tile ahead into registers while the current tile is
//     multiplied out of a double-buffered shared-memory tile (one barrier per k-tile);
//   * split-K over gridDim.z (atomic epilogues) so that short-and-wide problems still put a CTA on
//     every SM; 64x64x16 tiles, 256 threads, 4x4 register micro-tile.
// True-fp32 FMA keeps the 1e-4 parity bar of the fp32 path.  Sizes may live on the device (produced
// by the lookup kernels): CTAs past the real extent exit before touching memory.
#pragma once
#include "common.cuh"

namespace gemm {

// select one of four pointers without dynamic indexing of a kernel-parameter array (which would
// force the array into local memory)
template <class P>
__device__ __forceinline__ P sel4(const P (&p)[4], int s) {
    return s == 0 ? p[0] : (s == 1 ? p[1] : (s == 2 ? p[2] : p[3]));
}

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;

static inline bool aligned16(const void* p) { return (((uintptr_t)p) & 15) == 0; }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Every operand functor provides
//   vec        : host-computed flag: 16-byte vector loads along the contiguous dimension are legal
//   fast4(r,c) : elements (r, c..c+3) [K_CONTIG: c is k] or (c..c+3 rows, fixed other index) -- see below
//   elem(r,c)  : one element, no alignment assumptions
// For A operands the arguments are (m, k); for B operands (k, n).  The 4-vector always runs along the
// operand's contiguous dimension: k when K_CONTIG, otherwise m (A) / n (B).
// ----------------------------------------------------------------------------- A operands a(m,k)
struct ARowMajor {                 // a(m,k) = p[m*ld + k]
    static constexpr bool K_CONTIG = true;
    const float* p; int ld; bool vec;
    __device__ __forceinline__ float4 fast4(int m, int k) const { return ldg4(p + (size_t)m * ld + k); }
    __device__ __forceinline__ float elem(int m, int k) const { return __ldg(p + (size_t)m * ld + k); }
};
struct AColMajor {                 // a(m,k) = p[k*ld + m]        (A = Y^T), vector along m
    static constexpr bool K_CONTIG = false;
    const float* p; int ld; bool vec;
    __device__ __forceinline__ float4 fast4(int m, int k) const { return ldg4(p + (size_t)k * ld + m); }
    __device__ __forceinline__ float elem(int m, int k) const { return __ldg(p + (size_t)k * ld + m); }
};
struct AGather2 {                  // a(m,k) = k<K1 ? t1[row1(m)][k] : t2[row2(m)][k-K1];  row = map ? map[idx[m]] : idx[m]
    static constexpr bool K_CONTIG = true;
    const float* t1; int ld1; const int64_t* idx1; int K1;
    const float* t2; int ld2; const int64_t* idx2;
    const int64_t* map; bool vec;
    __device__ __forceinline__ const float* at(int m, int k) const {
        if (k < K1) {
            int64_t r = __ldg(idx1 + m); if (map) r = __ldg(map + r);
            return t1 + (size_t)r * ld1 + k;
        }
        int64_t r = __ldg(idx2 + m); if (map) r = __ldg(map + r);
        return t2 + (size_t)r * ld2 + (k - K1);
    }
    __device__ __forceinline__ float4 fast4(int m, int k) const { return ldg4(at(m, k)); }
    __device__ __forceinline__ float elem(int m, int k) const { return __ldg(at(m, k)); }
};
struct AEdgeAttr {                 // a(e,k) = k<Fe ? msg[row(e)][k] : cos(w[k-Fe]*rel[e] + b[k-Fe])
    static constexpr bool K_CONTIG = true;
    const float* msg; int ldm; const int64_t* eidx; int Fe;
    const float* rel; const float* tw; const float* tb; bool vec;
    __device__ __forceinline__ float tenc(int e, int k) const {
        return cosf(fmaf(__ldg(rel + e), __ldg(tw + (k - Fe)), __ldg(tb + (k - Fe))));
    }
    __device__ __forceinline__ float4 fast4(int e, int k) const {     // vec => Fe % 4 == 0: no straddle
        if (k < Fe) {
            const int64_t r = eidx ? __ldg(eidx + e) : (int64_t)e;
            return ldg4(msg + (size_t)r * ldm + k);
        }
        return make_float4(tenc(e, k), tenc(e, k + 1), tenc(e, k + 2), tenc(e, k + 3));
    }
    __device__ __forceinline__ float elem(int e, int k) const {
        if (k < Fe) {
            const int64_t r = eidx ? __ldg(eidx + e) : (int64_t)e;
            return __ldg(msg + (size_t)r * ldm + k);
        }
        return tenc(e, k);
    }
};

// ----------------------------------------------------------------------------- B operands b(k,n)
struct BWeight {                   // b(k,n) = W[n*ld + k]          (torch Linear weight [out,in])
    static constexpr bool K_CONTIG = true;
    const float* p; int ld; bool vec;
    __device__ __forceinline__ float4 fast4(int k, int n) const { return ldg4(p + (size_t)n * ld + k); }
    __device__ __forceinline__ float elem(int k, int n) const { return __ldg(p + (size_t)n * ld + k); }
};
struct BWeightSeg4 {               // four [D,ld] weights stacked along n
    static constexpr bool K_CONTIG = true;
    const float* p[4]; int D; int ld; bool vec;
    __device__ __forceinline__ const float* at(int k, int n) const { return sel4(p, n / D) + (size_t)(n % D) * ld + k; }
    __device__ __forceinline__ float4 fast4(int k, int n) const { return ldg4(at(k, n)); }
    __device__ __forceinline__ float elem(int k, int n) const { return __ldg(at(k, n)); }
};
struct BWeightCat2 {               // b(k,n) = k<K1 ? W1[n][k] : W2[n][k-K1]
    static constexpr bool K_CONTIG = true;
    const float* p1; int ld1; int K1; const float* p2; int ld2; bool vec;
    __device__ __forceinline__ const float* at(int k, int n) const {
        return k < K1 ? p1 + (size_t)n * ld1 + k : p2 + (size_t)n * ld2 + (k - K1);
    }
    __device__ __forceinline__ float4 fast4(int k, int n) const { return ldg4(at(k, n)); }
    __device__ __forceinline__ float elem(int k, int n) const { return __ldg(at(k, n)); }
};
struct BRowMajor {                 // b(k,n) = X[k*ld + n], vector along n
    static constexpr bool K_CONTIG = false;
    const float* p; int ld; bool vec;
    __device__ __forceinline__ float4 fast4(int k, int n) const { return ldg4(p + (size_t)k * ld + n); }
    __device__ __forceinline__ float elem(int k, int n) const { return __ldg(p + (size_t)k * ld + n); }
};
struct BRowSeg4 {                  // four [D,ld] matrices stacked along k
    static constexpr bool K_CONTIG = false;
    const float* p[4]; int D; int ld; bool vec;
    __device__ __forceinline__ const float* at(int k, int n) const { return sel4(p, k / D) + (size_t)(k % D) * ld + n; }
    __device__ __forceinline__ float4 fast4(int k, int n) const { return ldg4(at(k, n)); }
    __device__ __forceinline__ float elem(int k, int n) const { return __ldg(at(k, n)); }
};
struct BGather2 {                  // b(k,n) = n<K1 ? t1[row1(k)][n] : t2[row2(k)][n-K1]
    static constexpr bool K_CONTIG = false;
    const float* t1; int ld1; const int64_t* idx1; int K1;
    const float* t2; int ld2; const int64_t* idx2;
    const int64_t* map; bool vec;
    __device__ __forceinline__ const float* at(int k, int n) const {
        if (n < K1) {
            int64_t r = __ldg(idx1 + k); if (map) r = __ldg(map + r);
            return t1 + (size_t)r * ld1 + n;
        }
        int64_t r = __ldg(idx2 + k); if (map) r = __ldg(map + r);
        return t2 + (size_t)r * ld2 + (n - K1);
    }
    __device__ __forceinline__ float4 fast4(int k, int n) const { return ldg4(at(k, n)); }
    __device__ __forceinline__ float elem(int k, int n) const { return __ldg(at(k, n)); }
};
struct BEdgeAttr {                 // b(e,n) = n<Fe ? msg[row(e)][n] : cos(w*rel+b)
    static constexpr bool K_CONTIG = false;
    const float* msg; int ldm; const int64_t* eidx; int Fe;
    const float* rel; const float* tw; const float* tb; bool vec;
    __device__ __forceinline__ float tenc(int e, int n) const {
        return cosf(fmaf(__ldg(rel + e), __ldg(tw + (n - Fe)), __ldg(tb + (n - Fe))));
    }
    __device__ __forceinline__ float4 fast4(int e, int n) const {
        if (n < Fe) {
            const int64_t r = eidx ? __ldg(eidx + e) : (int64_t)e;
            return ldg4(msg + (size_t)r * ldm + n);
        }
        return make_float4(tenc(e, n), tenc(e, n + 1), tenc(e, n + 2), tenc(e, n + 3));
    }
    __device__ __forceinline__ float elem(int e, int n) const {
        if (n < Fe) {
            const int64_t r = eidx ? __ldg(eidx + e) : (int64_t)e;
            return __ldg(msg + (size_t)r * ldm + n);
        }
        return tenc(e, n);
    }
};
struct BRelu {                     // b(k,n) = max(H[k*ld+n], 0)
    static constexpr bool K_CONTIG = false;
    const float* p; int ld; bool vec;
    __device__ __forceinline__ float4 fast4(int k, int n) const {
        const float4 v = ldg4(p + (size_t)k * ld + n);
        return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    }
    __device__ __forceinline__ float elem(int k, int n) const { return fmaxf(__ldg(p + (size_t)k * ld + n), 0.f); }
};

// ----------------------------------------------------------------------------- epilogues
// split(): true when the epilogue accumulates atomically (legal with gridDim.z > 1).
// first : this CTA owns the k-range starting at 0 (adds bias / residual exactly once).
struct EpiStore {                  // C = acc + bias[n] + bias2[n] + res[m][n]
    static constexpr bool HAS_ROWSUM = false;
    float* C; int ldc; const float* bias; const float* bias2; const float* res; int ldr;
    __device__ __forceinline__ void operator()(int m, int n, float v, bool) const {
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
    __device__ __forceinline__ void operator()(int m, int n, float v, bool) const {
        const float* b = sel4(bias, n / D);
        if (b) v += __ldg(b + (n % D));
        C[(size_t)m * ldc + n] = v;
    }
    __device__ __forceinline__ void rowsum(int, float) const {}
};
struct EpiAtomic {                 // split-K accumulate: C[m][n] += acc (+ res[m][n] once); rs[m] += sum_k a(m,k)
    static constexpr bool HAS_ROWSUM = true;
    float* C; int ldc; float* rs; const float* res; int ldr;
    __device__ __forceinline__ void operator()(int m, int n, float v, bool first) const {
        if (res && first) v += res[(size_t)m * ldr + n];
        atomicAdd(C + (size_t)m * ldc + n, v);
    }
    __device__ __forceinline__ void rowsum(int m, float v) const { if (rs) atomicAdd(rs + m, v); }
};
struct EpiAtomicSeg4 {             // rows split into 4 segments of D rows, each with its own output / bias-grad
    static constexpr bool HAS_ROWSUM = true;
    float* C[4]; float* rs[4]; int D; int ldc;
    int antisym = 0;               // 1: segment 3 is the gradient of A = W - W^T - gamma I, accumulated straight into dW
    __device__ __forceinline__ void operator()(int m, int n, float v, bool) const {
        const int seg = m / D, r = m - seg * D;
        if (antisym && seg == 3) {                            // dW += dA - dA^T   (ldc == D)
            atomicAdd(C[3] + (size_t)r * ldc + n, v);
            atomicAdd(C[3] + (size_t)n * ldc + r, -v);
        } else {
            atomicAdd(sel4(C, seg) + (size_t)r * ldc + n, v);
        }
    }
    __device__ __forceinline__ void rowsum(int m, float v) const {
        float* r = sel4(rs, m / D);
        if (r) atomicAdd(r + (m % D), v);
    }
};
struct EpiScatterAtomic {          // C[row(m)][n] += acc      (row = map ? map[idx[m]] : idx[m])
    static constexpr bool HAS_ROWSUM = false;
    float* C; int ldc; const int64_t* idx; const int64_t* map;
    __device__ __forceinline__ void operator()(int m, int n, float v, bool) const {
        int64_t r = idx[m]; if (map) r = map[r];
        atomicAdd(C + (size_t)r * ldc + n, v);
    }
    __device__ __forceinline__ void rowsum(int, float) const {}
};

// ----------------------------------------------------------------------------- kernel
// Tile shapes: BT x BT output tile with (BT/4)^2 threads (4x4 register micro-tile each):
//   BT = 64 -> 256 threads (1 operand vector per thread per k-tile), BT = 32 -> 64 threads (2 vectors).
// The small tile is used when the big one would leave most SMs without a CTA (latency-bound shapes).
// Operand tile fetch: float4 index f -> K_CONTIG: (row = f/4, k = 4*(f%4)); otherwise (k = f/(BT/4),
// row4 = 4*(f%(BT/4))).  `lim_r` bounds the row index (M or N), [k0,ke) the k range.
template <class L, bool IS_A>
__device__ __forceinline__ float4 fetch4(const L& op, int r, int k, int lim_r, int ke) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (L::K_CONTIG) {
        if (r < lim_r && k < ke) {
            if (op.vec && k + 3 < ke) {
                v = IS_A ? op.fast4(r, k) : op.fast4(k, r);
            } else {
                v.x = IS_A ? op.elem(r, k) : op.elem(k, r);
                if (k + 1 < ke) v.y = IS_A ? op.elem(r, k + 1) : op.elem(k + 1, r);
                if (k + 2 < ke) v.z = IS_A ? op.elem(r, k + 2) : op.elem(k + 2, r);
                if (k + 3 < ke) v.w = IS_A ? op.elem(r, k + 3) : op.elem(k + 3, r);
            }
        }
    } else {
        if (k < ke && r < lim_r) {
            if (op.vec && r + 3 < lim_r) {
                v = IS_A ? op.fast4(r, k) : op.fast4(k, r);
            } else {
                v.x = IS_A ? op.elem(r, k) : op.elem(k, r);
                if (r + 1 < lim_r) v.y = IS_A ? op.elem(r + 1, k) : op.elem(k, r + 1);
                if (r + 2 < lim_r) v.z = IS_A ? op.elem(r + 2, k) : op.elem(k, r + 2);
                if (r + 3 < lim_r) v.w = IS_A ? op.elem(r + 3, k) : op.elem(k, r + 3);
            }
        }
    }
    return v;
}

template <int BT, class AL, class BL, class EP>
__global__ void __launch_bounds__((BT / 4) * (BT / 4))
gemm_kernel(const AL A, const BL Bm, const EP epi, int M_host, const int* __restrict__ M_dev, int N, int K_host,
            const int* __restrict__ K_dev, int k_chunk) {
    ctan_pdl_begin();
    constexpr int TPR = BT / 4, NT = TPR * TPR;         // threads per tile row, threads per CTA
    constexpr int VPT = (BT * BK / 4) / NT;             // operand vectors per thread per k-tile (1 or 2)
    __shared__ __align__(16) float As[2][BK][BT + PAD];
    __shared__ __align__(16) float Bs[2][BK][BT + PAD];
    const int M = M_dev ? *M_dev : M_host;
    const int K = K_dev ? *K_dev : K_host;
    const int n0 = blockIdx.x * BT;
    const int kb = blockIdx.z * k_chunk;
    const int ke = min(K, kb + k_chunk);
    if (n0 >= N || kb >= ke) return;                   // block-uniform, before any barrier
    const int t = threadIdx.x, ty = t / TPR, tx = t % TPR;
    // grid.y is bounded by the host (the true M may be far below its upper bound): a CTA walks row tiles
    // blockIdx.y, blockIdx.y + gridDim.y, ... so no more than a few waves of CTAs are ever launched
  for (int m0 = blockIdx.y * BT; m0 < M; m0 += gridDim.y * BT) {

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float rsum[4] = {0.f, 0.f, 0.f, 0.f};
    float4 ra[VPT], rb[VPT];

    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int f = t + i * NT;
            if (AL::K_CONTIG) ra[i] = fetch4<AL, true>(A, m0 + (f >> 2), k0 + ((f & 3) << 2), M, ke);
            else              ra[i] = fetch4<AL, true>(A, m0 + ((f % TPR) << 2), k0 + f / TPR, M, ke);
            if (BL::K_CONTIG) rb[i] = fetch4<BL, false>(Bm, n0 + (f >> 2), k0 + ((f & 3) << 2), N, ke);
            else              rb[i] = fetch4<BL, false>(Bm, n0 + ((f % TPR) << 2), k0 + f / TPR, N, ke);
        }
    };
    auto stash = [&](int st) {
#pragma unroll
        for (int i = 0; i < VPT; ++i) {
            const int f = t + i * NT;
            if (AL::K_CONTIG) {
                const int r = f >> 2, k = (f & 3) << 2;
                As[st][k][r] = ra[i].x; As[st][k + 1][r] = ra[i].y; As[st][k + 2][r] = ra[i].z; As[st][k + 3][r] = ra[i].w;
            } else {
                *reinterpret_cast<float4*>(&As[st][f / TPR][(f % TPR) << 2]) = ra[i];
            }
            if (BL::K_CONTIG) {
                const int r = f >> 2, k = (f & 3) << 2;
                Bs[st][k][r] = rb[i].x; Bs[st][k + 1][r] = rb[i].y; Bs[st][k + 2][r] = rb[i].z; Bs[st][k + 3][r] = rb[i].w;
            } else {
                *reinterpret_cast<float4*>(&Bs[st][f / TPR][(f % TPR) << 2]) = rb[i];
            }
        }
    };

    const int n_tiles = (ke - kb + BK - 1) / BK;
    fetch(kb);
    stash(0);
    if (n_tiles > 1) fetch(kb + BK);
    __syncthreads();
    for (int kt = 0; kt < n_tiles; ++kt) {
        const int st = kt & 1;
        if (kt + 1 < n_tiles) stash(st ^ 1);            // tile kt+1 (registers) -> the buffer read in iteration kt-1
        if (kt + 2 < n_tiles) fetch(kb + (kt + 2) * BK);  // tile kt+2 -> registers, in flight during the FMAs
        // shared-memory operands are read four k-steps at a time (8 LDS.128 in flight, then 64 FMAs): ncu
        // showed short-scoreboard (LDS latency) as the top stall when every k-step waited on its own pair
#pragma unroll
        for (int k0 = 0; k0 < BK; k0 += 4) {
            float4 a4[4], b4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a4[u] = *reinterpret_cast<const float4*>(&As[st][k0 + u][ty * 4]);
                b4[u] = *reinterpret_cast<const float4*>(&Bs[st][k0 + u][tx * 4]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float a[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
                const float b[4] = {b4[u].x, b4[u].y, b4[u].z, b4[u].w};
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
        __syncthreads();
    }
    const bool first = blockIdx.z == 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n < N) epi(m, n, acc[i][j], first);
        }
        if (EP::HAS_ROWSUM && blockIdx.x == 0 && tx == 0) epi.rowsum(m, rsum[i]);
    }
  }   // row-tile loop (the k-loop's trailing barrier already protects the smem tiles for the next pass)
}

// M_max bounds the grid when the true M lives on the device (the typical M is then taken as M_max/4 when
// choosing the tile shape); splits > 1 needs an atomic epilogue.
template <class AL, class BL, class EP>
static inline int launch(const AL& A, const BL& B, const EP& epi, int M_max, const int* M_dev, int N, int K_max,
                         const int* K_dev, int splits, cudaStream_t stream, int k_typ = 0) {
    if (M_max <= 0 || N <= 0 || K_max <= 0) return 0;
    if (splits < 1) splits = 1;
    // k_typ: the typical true K when K_max is only an upper bound (device-sized K): the chunk is cut from the
    // typical extent so that the ACTIVE splits cover it; splits past the true K exit immediately
    const int k_ref = (K_dev && k_typ > 0 && k_typ < K_max) ? k_typ : K_max;
    int k_chunk = (k_ref + splits - 1) / splits;
    k_chunk = ((k_chunk + BK - 1) / BK) * BK;
    splits = (K_max + k_chunk - 1) / k_chunk;
    const int m_typ = M_dev ? (M_max / 4 > 64 ? M_max / 4 : 64) : M_max;
    const long long big = (long long)((N + 63) / 64) * ((m_typ + 63) / 64) * splits;
    static const int force_tile = [] { const char* e = getenv("CTAN_GEMM_TILE"); return e ? atoi(e) : 0; }();
    // measured on B200 (scripts/micro_kernels.py): the 32x32 tile only wins on thin, long-K problems
    const bool use_big = force_tile == 64 ? true : (force_tile == 32 ? false : (big >= 24 || K_max < 512));
    const int bt = use_big ? 64 : 32;
    const int gx = (N + bt - 1) / bt;
    int gy = (M_max + bt - 1) / bt;
    if (M_dev) {        // device-sized M: launch at most ~6 waves of CTAs, each walks several row tiles
        const int cap = (6 * CTAN_SM_COUNT) / (gx * splits);
        if (gy > cap) gy = cap > 1 ? cap : 1;
    }
    dim3 grid(gx, gy, splits);
    if (use_big) CTAN_LAUNCH((gemm_kernel<64, AL, BL, EP>), grid, 256, 0, stream, A, B, epi, M_max, M_dev, N, K_max, K_dev, k_chunk);
    else         CTAN_LAUNCH((gemm_kernel<32, AL, BL, EP>), grid, 64, 0, stream, A, B, epi, M_max, M_dev, N, K_max, K_dev, k_chunk);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// number of K splits that puts ~2 CTAs on every SM; m_typ = the typical true M when M_max is only a bound
static inline int pick_splits(int m_typ, int N, int K, int min_k = 2 * BK) {
    const int tiles = ((m_typ + 31) / 32) * ((N + 31) / 32);      // small tiles are used when the grid is thin
    int s = (3 * CTAN_SM_COUNT + tiles - 1) / tiles;
    const int max_s = (K + min_k - 1) / min_k;
    if (s > max_s) s = max_s;
    if (s < 1) s = 1;
    return s;
}

}  // namespace gemm
