// Dense contractions of the CTAN step (forward and backward), all instances of gemm.cuh with the
// gathers / concatenations / time encoding fused into the operand loaders.
//
// Replaces:
//   a6/a7  memory gather + enc_x Linear          models/memory_layers.py:155-156, models/gnn_layers.py:96
//   a7/a8  edge_attr=[msg | cos(w*rel+b)] @ lin_edge   models/gnn_layers.py:97, TGB/modules/time_enc.py:23-24
//   a9/a10 q,k,v projections and x @ A^T          PyG TransformerConv / AntiSymmetricConv (App. A.5)
//   a12    Link/TGBLinkPredictor                  models/predictors.py:16-19,31-34
//   a15    their gradients (autograd in the reference, train_link.py:279)
#include "gemm.cuh"

using namespace gemm;

static inline bool v4(const void* p, int ld) { return p && aligned16(p) && (ld % 4 == 0); }

// X0[N,D] = [mem[n_id] | xfeat[n_id]] Wenc^T + benc       (gather fused; n_id int64 global ids)
// If z0 != nullptr the operand is the already-materialised dense [N, D+Fn] matrix instead.
extern "C" int ctan_enc_fwd(const float* mem, int D, const float* xfeat, int Fn, const int64_t* n_id,
                            const float* z0, int N, const int* N_dev, const float* Wenc, const float* benc,
                            float* X0, cudaStream_t stream) {
    const int K = D + Fn;
    BWeight B{Wenc, K, v4(Wenc, K)};
    EpiStore epi{X0, D, benc, nullptr, nullptr, 0};
    if (z0) {
        ARowMajor A{z0, K, v4(z0, K)};
        return launch(A, B, epi, N, N_dev, D, K, nullptr, 1, stream);
    }
    // the tail segment (node features) is rarely 4-aligned: vector loads only when both segments allow it
    // vectors never straddle the segment boundary when D % 4 == 0; a tail shorter than 4 is always scalar
    const bool vec = v4(mem, D) && (Fn < 4 || v4(xfeat, Fn));
    AGather2 A{mem, D, n_id, D, xfeat, Fn, n_id, nullptr, vec};
    return launch(A, B, epi, N, N_dev, D, K, nullptr, 1, stream);
}

// EP[E,D] = [msg[row(e)] | cos(tw*rel[e]+tb)] We^T        (row(e) = eidx ? eidx[e] : e)
extern "C" int ctan_eproj_fwd(const float* msg, int Fe, const int64_t* eidx, const float* rel, const float* tw,
                              const float* tb, int T, const float* We, int E, const int* E_dev, int D, float* EP,
                              cudaStream_t stream) {
    AEdgeAttr A{msg, Fe, eidx, Fe, rel, tw, tb, v4(msg, Fe) || Fe == 0};
    BWeight B{We, Fe + T, v4(We, Fe + T)};
    EpiStore epi{EP, D, nullptr, nullptr, nullptr, 0};
    return launch(A, B, epi, E, E_dev, D, Fe + T, nullptr, 1, stream);
}

// ---- the same projection for a SMALL edge set (a training batch: ~1e3 edges, K = Fe+T ~ 190): latency, not throughput.
// The tiled GEMM above walks K in six dependent gather rounds (19 us at E = 1.1 k); here every global load of a CTA is
// in flight at once: the WHOLE weight [D, K] (cp.async, L2-resident, shared by all CTAs) and the 16 gathered attribute
// rows land in shared memory behind one wait, then 4 warps x (4 rows x D/32 columns per lane) finish in ~2 us.
// Shared layout: row stride S floats with S/4 odd, so the 128-bit reads of 8 consecutive lanes (rows l*S) hit all 32 banks.
namespace eproj_rows {
constexpr int TE = 16, THREADS = 128;
__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_wait_all() { asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory"); }

template <int NC>   // D = 32 * NC
__global__ void __launch_bounds__(THREADS)
kernel(const float* __restrict__ msg, int Fe, const int64_t* __restrict__ eidx, const float* __restrict__ rel,
       const float* __restrict__ tw, const float* __restrict__ tb, int T, const float* __restrict__ We, int E_host,
       const int* __restrict__ E_dev, float* __restrict__ EP, int S, int vec_msg) {
    ctan_pdl_begin();
    extern __shared__ __align__(16) float sm[];
    constexpr int D = 32 * NC;
    const int E = E_dev ? *E_dev : E_host;
    if ((long long)blockIdx.x * TE >= E) return;              // (the grid is sized by a bound)
    const int K = Fe + T, K4 = K >> 2;
    float* Ws = sm;                                           // [D][S]
    float* At = sm + (size_t)D * S;                           // [TE][S]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < D * K4; i += THREADS) {             // the weight, once per CTA
        const int d = i / K4, k4 = i - d * K4;
        cp16(Ws + (size_t)d * S + 4 * k4, We + (size_t)d * K + 4 * k4);
    }
    for (int g = blockIdx.x; (long long)g * TE < E; g += gridDim.x) {
        const int e0 = g * TE;
        if (g != (int)blockIdx.x) __syncthreads();            // the previous group's rows are still being read
        if (vec_msg) {
            const int F4 = Fe >> 2;
            for (int i = tid; i < TE * F4; i += THREADS) {
                const int r = i / F4, c4 = i - r * F4, e = e0 + r;
                if (e < E) {
                    const int64_t row = eidx ? __ldg(eidx + e) : (int64_t)e;
                    cp16(At + (size_t)r * S + 4 * c4, msg + (size_t)row * Fe + 4 * c4);
                }
            }
        } else {
            for (int i = tid; i < TE * Fe; i += THREADS) {
                const int r = i / Fe, c = i - r * Fe, e = e0 + r;
                if (e < E) {
                    const int64_t row = eidx ? __ldg(eidx + e) : (int64_t)e;
                    At[(size_t)r * S + c] = __ldg(msg + (size_t)row * Fe + c);
                }
            }
        }
        for (int i = tid; i < TE * T; i += THREADS) {         // cos(w * rel + b)   (TGB/modules/time_enc.py:23-24)
            const int r = i / T, j = i - r * T, e = e0 + r;
            if (e < E) At[(size_t)r * S + Fe + j] = cosf(fmaf(__ldg(rel + e), __ldg(tw + j), __ldg(tb + j)));
        }
        cp_wait_all();
        __syncthreads();
        float acc[4][NC];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int j = 0; j < NC; ++j) acc[r][j] = 0.f;
        const float* a0 = At + (size_t)(4 * warp) * S;
        const float* w0 = Ws + (size_t)lane * S;
#pragma unroll 2
        for (int k4 = 0; k4 < K4; ++k4) {
            float4 a[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) a[r] = *reinterpret_cast<const float4*>(a0 + (size_t)r * S + 4 * k4);
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const float4 b = *reinterpret_cast<const float4*>(w0 + (size_t)(32 * j) * S + 4 * k4);
#pragma unroll
                for (int r = 0; r < 4; ++r) {                  // k ascending, like the tiled kernel's accumulation order
                    acc[r][j] = fmaf(a[r].x, b.x, acc[r][j]);
                    acc[r][j] = fmaf(a[r].y, b.y, acc[r][j]);
                    acc[r][j] = fmaf(a[r].z, b.z, acc[r][j]);
                    acc[r][j] = fmaf(a[r].w, b.w, acc[r][j]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int e = e0 + 4 * warp + r;
            if (e < E) {
#pragma unroll
                for (int j = 0; j < NC; ++j) EP[(size_t)e * D + lane + 32 * j] = acc[r][j];
            }
        }
    }
}
}  // namespace eproj_rows

// Same result as ctan_eproj_fwd (same k-ascending fp32 accumulation), for edge sets of a few thousand rows at most.
// Requirements: D in {32,64,128,256}, (Fe+T) % 4 == 0, We 16-byte aligned, D*(K+4)*4 + 16*(K+4)*4 bytes of shared memory
// available; otherwise CTAN_ERR_UNSUPPORTED (the caller keeps ctan_eproj_fwd).
extern "C" int ctan_eproj_fwd_rows(const float* msg, int Fe, const int64_t* eidx, const float* rel, const float* tw,
                                   const float* tb, int T, const float* We, int E, const int* E_dev, int D, float* EP,
                                   cudaStream_t stream) {
    if (E <= 0) return 0;
    const int K = Fe + T;
    if (K <= 0 || (K & 3) || !aligned16(We) || (D != 32 && D != 64 && D != 128 && D != 256)) return CTAN_ERR_UNSUPPORTED;
    const int S = ((K >> 2) & 1) ? K : K + 4;                 // S/4 odd
    const size_t smem = ((size_t)D + eproj_rows::TE) * S * sizeof(float);
    if (smem > 220 * 1024) return CTAN_ERR_UNSUPPORTED;
    const int vec_msg = (Fe > 0 && (Fe & 3) == 0 && aligned16(msg)) ? 1 : 0;
    long long groups = ((long long)E + eproj_rows::TE - 1) / eproj_rows::TE;
    const int grid = (int)(groups < 2LL * CTAN_SM_COUNT ? groups : 2LL * CTAN_SM_COUNT);
#define CTAN_EPROJ_ROWS(NC)                                                                                               \
    {                                                                                                                     \
        cudaError_t e = cudaFuncSetAttribute(eproj_rows::kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return (int)e;                                                                              \
        CTAN_LAUNCH((eproj_rows::kernel<NC>), grid, eproj_rows::THREADS, smem, stream, msg, Fe, eidx, rel, tw, tb, T, We, \
                    E, E_dev, EP, S, vec_msg);                                                                            \
    }
    if (D == 32) CTAN_EPROJ_ROWS(1) else if (D == 64) CTAN_EPROJ_ROWS(2) else if (D == 128) CTAN_EPROJ_ROWS(4) else CTAN_EPROJ_ROWS(8)
#undef CTAN_EPROJ_ROWS
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

// QKVA[N,4D] = X [Wq;Wk;Wv;A]^T + [bq;bk;bv;b]
extern "C" int ctan_qkva_fwd(const float* X, int N, const int* N_dev, int D, const float* Wq, const float* Wk,
                             const float* Wv, const float* A_, const float* bq, const float* bk, const float* bv,
                             const float* b, float* QKVA, cudaStream_t stream) {
    ARowMajor A{X, D, v4(X, D)};
    BWeightSeg4 B{{Wq, Wk, Wv, A_}, D, D, v4(Wq, D) && v4(Wk, D) && v4(Wv, D) && v4(A_, D)};
    EpiStoreSeg4 epi{QKVA, 4 * D, {bq, bk, bv, b}, D};
    return launch(A, B, epi, N, N_dev, 4 * D, D, nullptr, 1, stream);
}

// ---------------------------------------------------------------------------------------- readout
// H[P,Hd] = z[row(src)] Wsrc^T + z[row(dst)] Wdst^T + bsrc + bdst   (pre-activation, saved for backward)
// OUT[P,O] = relu(H) Wfin^T + bfin                                   row(g) = map ? map[g] : g
__global__ void head_fwd_kernel(const float* __restrict__ H, int P_host, const int* __restrict__ P_dev, int Hd, int O,
                                const float* __restrict__ Wfin, const float* __restrict__ bfin,
                                float* __restrict__ OUT) {
    ctan_pdl_begin();
    const int P = P_dev ? *P_dev : P_host;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < P; p += nw) {
        for (int o = 0; o < O; ++o) {
            float s = 0.f;
            for (int h = lane; h < Hd; h += 32) s = fmaf(fmaxf(H[(size_t)p * Hd + h], 0.f), __ldg(Wfin + (size_t)o * Hd + h), s);
            s = warp_sum(s);
            if (lane == 0) OUT[(size_t)p * O + o] = s + __ldg(bfin + o);
        }
    }
}

struct EpiAtomicBias2 {            // split-K: C += acc (+ bias + bias2 once)
    static constexpr bool HAS_ROWSUM = false;
    float* C; int ldc; const float* bias; const float* bias2;
    __device__ __forceinline__ void operator()(int m, int n, float v, bool first) const {
        if (first) {
            if (bias) v += __ldg(bias + n);
            if (bias2) v += __ldg(bias2 + n);
        }
        atomicAdd(C + (size_t)m * ldc + n, v);
    }
    __device__ __forceinline__ void rowsum(int, float) const {}
};

// h_zeroed != 0: the caller guarantees H is all zeros, which lets the hidden contraction split K over
// more CTAs (P is a few hundred rows: 7 CTAs otherwise).
extern "C" int ctan_readout_fwd(const float* Z, int D, const int64_t* src, const int64_t* dst, const int64_t* map,
                                int P, const int* P_dev, int Hd, int O, const float* Wsrc, const float* bsrc,
                                const float* Wdst, const float* bdst, const float* Wfin, const float* bfin, float* H,
                                float* OUT, int h_zeroed, cudaStream_t stream) {
    if (P <= 0) return 0;
    const int K1 = Wsrc ? D : 0;          // Wsrc == nullptr: SequencePredictor (destination row only)
    AGather2 A{Z, D, src, K1, Z, D, dst, map, v4(Z, D)};
    BWeightCat2 B{Wsrc, D, K1, Wdst, D, v4(Wdst, D) && (!Wsrc || v4(Wsrc, D))};
    int rc;
    if (h_zeroed) {
        EpiAtomicBias2 epi{H, Hd, bsrc, bdst};
        rc = launch(A, B, epi, P, P_dev, Hd, K1 + D, nullptr, pick_splits(P, Hd, K1 + D), stream);
    } else {
        EpiStore epi{H, Hd, bsrc, bdst, nullptr, 0};
        rc = launch(A, B, epi, P, P_dev, Hd, K1 + D, nullptr, 1, stream);
    }
    if (rc) return rc;
    if (!OUT) return 0;                 // hidden layer only (the caller fuses the head with the loss)
    CTAN_LAUNCH((head_fwd_kernel), ctan_grid((long long)P * 32, 256), 256, 0, stream, H, P, P_dev, Hd, O, Wfin, bfin, OUT);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// Fused binary head + loss + their gradients (O = 1), one warp per scored pair:
//   OUT[p]  = relu(H[p]) . Wfin + bfin
//   loss   += softplus(-OUT[p])/n_pos (p < n_pos)  |  softplus(OUT[p])/n_neg (else)      (train_link.py:269-270)
//   dOUT[p] = dloss/dOUT[p];  dHid[p,h] = (H[p,h] > 0) dOUT[p] Wfin[h]
// loss is accumulated (atomicAdd) into loss_hist[(counters[3]-1) % hist_cap], which must be zero on entry
// (ctan_stage_batch clears it).  dOUT/dHid may be NULL (inference: loss only).
__global__ void head_loss_kernel(const float* __restrict__ H, int P, int Hd, const float* __restrict__ Wfin,
                                 const float* __restrict__ bfin, int n_pos, float* __restrict__ OUT,
                                 float* __restrict__ dOUT, float* __restrict__ dHid, float* __restrict__ loss_hist,
                                 int hist_cap, const int64_t* __restrict__ counters) {
    ctan_pdl_begin();
    __shared__ float s_loss[8];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int p = blockIdx.x * (blockDim.x >> 5) + wib;
    float my = 0.f;
    if (p < P) {
        float s = 0.f;
        for (int h = lane; h < Hd; h += 32) s = fmaf(fmaxf(H[(size_t)p * Hd + h], 0.f), __ldg(Wfin + h), s);
        const float x = warp_sum(s) + __ldg(bfin);
        const int n_neg = P - n_pos;
        float dx;
        if (p < n_pos) {
            my = (fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)))) / (float)n_pos;
            dx = -1.f / (1.f + expf(x)) / (float)n_pos;
        } else {
            my = (fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)))) / (float)n_neg;
            dx = 1.f / (1.f + expf(-x)) / (float)n_neg;
        }
        if (lane == 0) { OUT[p] = x; if (dOUT) dOUT[p] = dx; }
        if (dHid)
            for (int h = lane; h < Hd; h += 32)
                dHid[(size_t)p * Hd + h] = H[(size_t)p * Hd + h] > 0.f ? dx * __ldg(Wfin + h) : 0.f;
    }
    if (lane == 0) s_loss[wib] = my;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += s_loss[i];
        long long slot = counters ? (long long)counters[3] - 1 : 0;
        if (slot < 0) slot = 0;
        atomicAdd(loss_hist + (slot % hist_cap), t);
    }
}

extern "C" int ctan_head_loss(const float* H, int P, int Hd, const float* Wfin, const float* bfin, int n_pos,
                              float* OUT, float* dOUT, float* dHid, float* loss_hist, int hist_cap,
                              const int64_t* counters, cudaStream_t stream) {
    if (P <= 0 || n_pos <= 0 || n_pos >= P || hist_cap <= 0) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((head_loss_kernel), ctan_div_up(P, 8), 256, 0, stream, H, P, Hd, Wfin, bfin, n_pos, OUT, dOUT, dHid, loss_hist,
                                                            hist_cap, counters);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// ---- the whole binary link readout of a training batch in ONE launch (models/predictors.py:16-19 / 31-34 +
// train_link.py:269-270 + their gradients): hidden layer, head, BCE loss, dOUT, dHid and the scatter of dZ.
// P is a few hundred pairs and the three kernels it replaces (hidden contraction, head + loss, dHid -> dZ contraction)
// are each pure launch + gather latency on the step's critical path.  A CTA scores PT = 256/Hd pairs: both weight
// matrices ([Hd, D] each, L2-resident, shared by all CTAs) and the gathered embedding rows land in shared memory behind
// one wait (cp.async); thread (pair, h) forms H[pair][h]; the pair's threads reduce the head; thread (pair, side, d)
// forms dZ row elements from dHid (kept in shared memory) and adds them atomically (a node is scored by several pairs).
// H (pre-activation, biases included) and dHid are also written out for the weight-gradient kernels.
namespace link_fused {
constexpr int THREADS = 256;
__device__ __forceinline__ void cp16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__global__ void __launch_bounds__(THREADS)
kernel(const float* __restrict__ Z, int D, const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
       const int64_t* __restrict__ map, int P, int n_pos, int Hd, const float* __restrict__ Wsrc,
       const float* __restrict__ bsrc, const float* __restrict__ Wdst, const float* __restrict__ bdst,
       const float* __restrict__ Wfin, const float* __restrict__ bfin, float* __restrict__ H, float* __restrict__ OUT,
       float* __restrict__ dOUT, float* __restrict__ dHid, float* __restrict__ dZ, float* __restrict__ loss_hist,
       int hist_cap, const int64_t* __restrict__ counters, int S) {
    ctan_pdl_begin();
    extern __shared__ __align__(16) float lsm[];
    const int PT = THREADS / Hd;
    float* Ws = lsm;                                          // [Hd][S]
    float* Wd = Ws + (size_t)Hd * S;                          // [Hd][S]
    float* zs = Wd + (size_t)Hd * S;                          // [PT][2][D]   (src row | dst row)
    float* dh = zs + (size_t)PT * 2 * D;                      // [PT][Hd]
    float* red = dh + (size_t)PT * Hd;                        // [8] per-warp partial heads
    __shared__ int64_t rows[16];                              // [PT][2] resolved local rows
    __shared__ float s_loss;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int p0 = blockIdx.x * PT;
    const int D4 = D >> 2;
    if (tid == 0) s_loss = 0.f;
    for (int i = tid; i < Hd * D4; i += THREADS) {            // the weights: no dependence, issued first
        const int h = i / D4, c = i - h * D4;
        cp16(Ws + (size_t)h * S + 4 * c, Wsrc + (size_t)h * D + 4 * c);
        cp16(Wd + (size_t)h * S + 4 * c, Wdst + (size_t)h * D + 4 * c);
    }
    for (int i = tid; i < PT * 2 * D4; i += THREADS) {        // the two embedding rows of every pair
        const int pr = i / D4, c = i - pr * D4;               // pr = pair-in-CTA * 2 + side
        const int p = p0 + (pr >> 1);
        if (p < P) {
            int64_t r = (pr & 1) ? dst[p] : src[p];
            if (map) r = map[r];
            if (c == 0) rows[pr] = r;
            cp16(zs + (size_t)pr * D + 4 * c, Z + (size_t)r * D + 4 * c);
        }
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    // hidden layer: thread (pp, h)
    const int pp = tid / Hd, h = tid - pp * Hd;
    const int p = p0 + pp;
    const bool live = p < P;
    float acc = 0.f;
    if (live) {
        acc = __ldg(bsrc + h) + __ldg(bdst + h);
        const float4* ws = reinterpret_cast<const float4*>(Ws + (size_t)h * S);
        const float4* wd = reinterpret_cast<const float4*>(Wd + (size_t)h * S);
        const float4* a = reinterpret_cast<const float4*>(zs + (size_t)(2 * pp) * D);
        const float4* b = reinterpret_cast<const float4*>(zs + (size_t)(2 * pp + 1) * D);
#pragma unroll 4
        for (int c = 0; c < D4; ++c) {
            const float4 x = a[c], w = ws[c], y = b[c], v = wd[c];
            acc = fmaf(x.x, w.x, acc); acc = fmaf(x.y, w.y, acc); acc = fmaf(x.z, w.z, acc); acc = fmaf(x.w, w.w, acc);
            acc = fmaf(y.x, v.x, acc); acc = fmaf(y.y, v.y, acc); acc = fmaf(y.z, v.z, acc); acc = fmaf(y.w, v.w, acc);
        }
        H[(size_t)p * Hd + h] = acc;
    }
    const float wf = __ldg(Wfin + h);
    float part = live ? fmaxf(acc, 0.f) * wf : 0.f;
    part = warp_sum(part);                                    // Hd >= 32: a warp never straddles two pairs
    if (lane == 0) red[warp] = part;
    __syncthreads();
    const int wpp = Hd >> 5;                                  // warps per pair
    float x = __ldg(bfin);
    for (int i = 0; i < wpp; ++i) x += red[pp * wpp + i];
    float dx = 0.f;
    if (live) {
        const int n_neg = P - n_pos;
        float my;
        if (p < n_pos) {
            my = (fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)))) / (float)n_pos;
            dx = -1.f / (1.f + expf(x)) / (float)n_pos;
        } else {
            my = (fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)))) / (float)n_neg;
            dx = 1.f / (1.f + expf(-x)) / (float)n_neg;
        }
        if (h == 0) {
            OUT[p] = x;
            if (dOUT) dOUT[p] = dx;
            atomicAdd(&s_loss, my);
        }
        const float g = acc > 0.f ? dx * wf : 0.f;
        dh[pp * Hd + h] = g;
        if (dHid) dHid[(size_t)p * Hd + h] = g;
    }
    __syncthreads();
    if (tid == 0) {
        long long slot = counters ? (long long)counters[3] - 1 : 0;
        if (slot < 0) slot = 0;
        atomicAdd(loss_hist + (slot % hist_cap), s_loss);
    }
    if (!dZ) return;
    // dZ[row(src)] += dHid Wsrc ; dZ[row(dst)] += dHid Wdst : thread (pair, side, d)
    for (int i = tid; i < PT * 2 * D; i += THREADS) {
        const int pr = i / D, d = i - pr * D;
        const int q = pr >> 1;
        if (p0 + q >= P) continue;
        const float* W = (pr & 1) ? Wd : Ws;
        const float* g = dh + (size_t)q * Hd;
        float v = 0.f;
#pragma unroll 8
        for (int k = 0; k < Hd; ++k) v = fmaf(g[k], W[(size_t)k * S + d], v);
        atomicAdd(dZ + (size_t)rows[pr] * D + d, v);
    }
}
}  // namespace link_fused

// Requirements: O == 1 (the caller's head), Hd in {32,64,128,256}, D % 4 == 0, Z / Wsrc / Wdst 16-byte aligned rows,
// (2 Hd (D+4) + (256/Hd)(2D + Hd) + 8) floats of shared memory <= 200 KB; else CTAN_ERR_UNSUPPORTED (the caller keeps
// ctan_readout_fwd + ctan_head_loss + ctan_readout_bwd).  dOUT / dHid / dZ may be NULL (no-grad step: scores + loss).
// loss_hist slot as in ctan_head_loss.  dZ is ACCUMULATED (zero it first).
extern "C" int ctan_readout_link_fused(const float* Z, int D, const int64_t* src, const int64_t* dst, const int64_t* map,
                                       int P, int n_pos, int Hd, const float* Wsrc, const float* bsrc, const float* Wdst,
                                       const float* bdst, const float* Wfin, const float* bfin, float* H, float* OUT,
                                       float* dOUT, float* dHid, float* dZ, float* loss_hist, int hist_cap,
                                       const int64_t* counters, cudaStream_t stream) {
    if (P <= 0 || n_pos <= 0 || n_pos >= P || hist_cap <= 0) return CTAN_ERR_BAD_ARG;
    if ((Hd != 32 && Hd != 64 && Hd != 128 && Hd != 256) || (D & 3) || !v4(Z, D) || !v4(Wsrc, D) || !v4(Wdst, D))
        return CTAN_ERR_UNSUPPORTED;
    const int S = ((D >> 2) & 1) ? D : D + 4;                 // S/4 odd: conflict-free 128-bit reads of 8 consecutive rows
    const int PT = link_fused::THREADS / Hd;
    const size_t smem = ((size_t)2 * Hd * S + (size_t)PT * (2 * D + Hd) + 8) * sizeof(float);
    if (smem > 200 * 1024) return CTAN_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(link_fused::kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    CTAN_LAUNCH((link_fused::kernel), ctan_div_up(P, PT), link_fused::THREADS, smem, stream, Z, D, src, dst, map, P, n_pos, Hd,
                Wsrc, bsrc, Wdst, bdst, Wfin, bfin, H, OUT, dOUT, dHid, dZ, loss_hist, hist_cap, counters, S);
    CTAN_CHECK_LAUNCH();
    return 0;
}

struct BRowCat2 {                  // b(k,n) = n<D ? W1[k*ld+n] : W2[k*ld+n-D]   (vector along n)
    static constexpr bool K_CONTIG = false;
    const float* p1; const float* p2; int D; int ld; bool vec;
    __device__ __forceinline__ const float* at(int k, int n) const {
        return n < D ? p1 + (size_t)k * ld + n : p2 + (size_t)k * ld + (n - D);
    }
    __device__ __forceinline__ float4 fast4(int k, int n) const { return ldg4(at(k, n)); }
    __device__ __forceinline__ float elem(int k, int n) const { return __ldg(at(k, n)); }
};
struct EpiScatterAtomic2 {         // n<D: C[row(idx1[m])][n] += v ; else C[row(idx2[m])][n-D] += v
    static constexpr bool HAS_ROWSUM = false;
    float* C; int D; const int64_t* idx1; const int64_t* idx2; const int64_t* map;
    __device__ __forceinline__ void operator()(int m, int n, float v, bool) const {
        int64_t r = n < D ? idx1[m] : idx2[m];
        if (map) r = map[r];
        atomicAdd(C + (size_t)r * D + (n < D ? n : n - D), v);
    }
    __device__ __forceinline__ void rowsum(int, float) const {}
};

// dHid[p,h] = (H>0) * sum_o dOUT[p,o] Wfin[o,h]
__global__ void head_bwd_kernel(const float* __restrict__ dOUT, const float* __restrict__ H, int P_host,
                                const int* __restrict__ P_dev, int Hd, int O, const float* __restrict__ Wfin,
                                float* __restrict__ dHid) {
    ctan_pdl_begin();
    const int P = P_dev ? *P_dev : P_host;
    const long long n = (long long)P * Hd;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / Hd), h = (int)(i % Hd);
        float s = 0.f;
        if (H[i] > 0.f)
            for (int o = 0; o < O; ++o) s = fmaf(dOUT[(size_t)p * O + o], __ldg(Wfin + (size_t)o * Hd + h), s);
        dHid[i] = s;
    }
}

struct EpiAtomicColSeg2 {          // columns split at K1: [C0 | C1] (each row stride D); rowsum added to rs0 and rs1
    static constexpr bool HAS_ROWSUM = true;
    float* C0; float* C1; int K1; int D; float* rs0; float* rs1;
    __device__ __forceinline__ void operator()(int m, int n, float v, bool) const {
        if (n < K1) atomicAdd(C0 + (size_t)m * D + n, v); else atomicAdd(C1 + (size_t)m * D + (n - K1), v);
    }
    __device__ __forceinline__ void rowsum(int m, float v) const {
        if (rs0) atomicAdd(rs0 + m, v);
        if (rs1) atomicAdd(rs1 + m, v);
    }
};

// Gradients of the readout.  All outputs are ACCUMULATED (+=): zero them first.
//   dZ[N,D] += scatter(dHid Wsrc -> row(src)) + scatter(dHid Wdst -> row(dst))
//   dWsrc,dWdst [Hd,D], dbsrc,dbdst [Hd], dWfin [O,Hd], dbfin [O];  dHid [P,Hd] is scratch.
// Two independent halves (so a caller can put them on different streams):
//   dZ != NULL    -> dHid = head backward, then the two scatter contractions into dZ
//   dWdst != NULL -> the weight/bias gradients (reads dHid computed by the first half)
extern "C" int ctan_readout_bwd(const float* dOUT, const float* H, const float* Z, int D, const int64_t* src,
                                const int64_t* dst, const int64_t* map, int P, const int* P_dev, int Hd, int O,
                                const float* Wsrc, const float* Wdst, const float* Wfin, float* dHid, float* dZ,
                                float* dWsrc, float* dbsrc, float* dWdst, float* dbdst, float* dWfin, float* dbfin,
                                int dhid_ready, cudaStream_t stream) {
    if (P <= 0) return 0;
    int rc;
    if (dZ) {
        if (!dhid_ready) {              // (ctan_head_loss already produced dHid in the fused training step)
            CTAN_LAUNCH((head_bwd_kernel), ctan_grid((long long)P * Hd, 256), 256, 0, stream, dOUT, H, P, P_dev, Hd, O, Wfin,
                                                                                  dHid);
            CTAN_CHECK_LAUNCH();
        }
        // dZ[row(src)] += dHid Wsrc ; dZ[row(dst)] += dHid Wdst   -- one contraction against [Wsrc | Wdst]
        ARowMajor A{dHid, Hd, v4(dHid, Hd)};
        if (Wsrc) {
            BRowCat2 B{Wsrc, Wdst, D, D, v4(Wsrc, D) && v4(Wdst, D)};
            EpiScatterAtomic2 e{dZ, D, src, dst, map};
            rc = launch(A, B, e, P, P_dev, 2 * D, Hd, nullptr, pick_splits(P, 2 * D, Hd), stream);
        } else {
            BRowMajor B2{Wdst, D, v4(Wdst, D)};
            EpiScatterAtomic e2{dZ, D, dst, map};
            rc = launch(A, B2, e2, P, P_dev, D, Hd, nullptr, pick_splits(P, D, Hd), stream);
        }
        if (rc) return rc;
    }
    if (dWdst) {
        {   // dWfin[O,Hd] += dOUT^T relu(H);  dbfin += colsum(dOUT)
            AColMajor A{dOUT, O, v4(dOUT, O)};
            BRelu B{H, Hd, v4(H, Hd)};
            EpiAtomic epi{dWfin, Hd, dbfin, nullptr, 0};
            rc = launch(A, B, epi, O, nullptr, Hd, P, P_dev, pick_splits(O, Hd, P), stream);
            if (rc) return rc;
        }
        {   // [dWsrc | dWdst] += dHid^T [z[src] | z[dst]];  dbsrc,dbdst += colsum(dHid)
            const int K1 = Wsrc ? D : 0;
            AColMajor A{dHid, Hd, v4(dHid, Hd)};
            BGather2 B{Z, D, src, K1, Z, D, dst, map, v4(Z, D)};
            EpiAtomicColSeg2 epi{dWsrc, dWdst, K1, D, dbsrc, dbdst};
            rc = launch(A, B, epi, Hd, nullptr, K1 + D, P, P_dev, pick_splits(Hd, K1 + D, P), stream);
            if (rc) return rc;
        }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------- embed backward
// One anti-symmetric iteration, given DQKVA[N,4D] (from ctan_edge_bwd) and G = dL/dx_out:
//   dXin[N,D] += G + DQKVA [Wq;Wk;Wv;A]          (dXin must be ZERO on entry: the contraction splits K)
//   dWq,dWk,dWv,dA [D,D] += DQKVA_seg^T X ;  dbq,dbk,dbv,db [D] += colsum(DQKVA_seg)     (accumulated)
// Either half may be skipped (NULL outputs) so the two can run on different streams.
// n_typ: typical true N when N is only an upper bound (sizes the K split); 0 = N.
extern "C" int ctan_qkva_bwd(const float* DQKVA, const float* X, const float* G, int N, const int* N_dev, int D,
                             const float* Wq, const float* Wk, const float* Wv, const float* A_, float* dXin,
                             float* dWq, float* dWk, float* dWv, float* dA, float* dbq, float* dbk, float* dbv,
                             float* db, int n_typ, cudaStream_t stream) {
    if (N <= 0) return 0;
    if (n_typ <= 0 || n_typ > N) n_typ = N;
    int rc;
    if (dXin) {
        ARowMajor A{DQKVA, 4 * D, v4(DQKVA, 4 * D)};
        BRowSeg4 B{{Wq, Wk, Wv, A_}, D, D, v4(Wq, D) && v4(Wk, D) && v4(Wv, D) && v4(A_, D)};
        EpiAtomic epi{dXin, D, nullptr, G, D};
        rc = launch(A, B, epi, N, N_dev, D, 4 * D, nullptr, pick_splits(n_typ, D, 4 * D), stream);
        if (rc) return rc;
    }
    if (dWq) {
        AColMajor A{DQKVA, 4 * D, v4(DQKVA, 4 * D)};
        BRowMajor B{X, D, v4(X, D)};
        EpiAtomicSeg4 epi{{dWq, dWk, dWv, dA}, {dbq, dbk, dbv, db}, D, D};
        rc = launch(A, B, epi, 4 * D, nullptr, D, N, N_dev, pick_splits(4 * D, D, n_typ), stream, n_typ);
        if (rc) return rc;
    }
    return 0;
}

// The weight-gradient half alone, for the training engine: the fourth segment goes straight into dW (+dA - dA^T, no
// separate ctan_antisym_bwd pass) and, for batch-sized N, K is cut in 64-row chunks so that ~150 CTAs are active
// whatever the true N is (the estimate n_typ only bounds the grid).
extern "C" int ctan_qkva_wgrad(const float* DQKVA, const float* X, int N, const int* N_dev, int D, float* dWq, float* dWk,
                               float* dWv, float* dW, float* dbq, float* dbk, float* dbv, float* db, int n_typ,
                               cudaStream_t stream) {
    if (N <= 0) return 0;
    if (n_typ <= 0 || n_typ > N) n_typ = N;
    AColMajor A{DQKVA, 4 * D, v4(DQKVA, 4 * D)};
    BRowMajor B{X, D, v4(X, D)};
    EpiAtomicSeg4 epi{{dWq, dWk, dWv, dW}, {dbq, dbk, dbv, db}, D, D, 1};
    const int splits = n_typ <= 4096 ? (n_typ + 63) / 64 : pick_splits(4 * D, D, n_typ);
    return launch(A, B, epi, 4 * D, nullptr, D, N, N_dev, splits, stream, n_typ);
}

// dWenc[D, D+Fn] += dX0^T Z0 ; dbenc[D] += colsum(dX0)          (Z0 = materialised [mem|x] rows)
extern "C" int ctan_enc_bwd(const float* dX0, const float* Z0, int N, const int* N_dev, int D, int Fn, float* dWenc,
                            float* dbenc, int n_typ, cudaStream_t stream) {
    if (N <= 0) return 0;
    if (n_typ <= 0 || n_typ > N) n_typ = N;
    AColMajor A{dX0, D, v4(dX0, D)};
    BRowMajor B{Z0, D + Fn, v4(Z0, D + Fn)};
    EpiAtomic epi{dWenc, D + Fn, dbenc, nullptr, 0};
    // (finer K chunks than pick_splits' were measured slower here: 5.7 -> 8.7 us with 32-row chunks at N = 640)
    return launch(A, B, epi, D, nullptr, D + Fn, N, N_dev, pick_splits(D, D + Fn, n_typ), stream, n_typ);
}

// dWe[D, Fe+T] += dEP^T [msg | cos(.)] ;  dTenc[E,T] = dEP We[:,Fe:]   (scratch for ctan_tenc_bwd)
extern "C" int ctan_eproj_bwd(const float* dEP, const float* msg, int Fe, const int64_t* eidx, const float* rel,
                              const float* tw, const float* tb, int T, const float* We, int E, const int* E_dev, int D,
                              float* dWe, float* dTenc, int e_typ, cudaStream_t stream) {
    if (E <= 0) return 0;
    if (e_typ <= 0 || e_typ > E) e_typ = E;
    int rc;
    if (dWe) {                          // either half may be skipped (NULL output): the two are independent and a caller
                                        // can run them on different streams
        AColMajor A{dEP, D, v4(dEP, D)};
        BEdgeAttr B{msg, Fe, eidx, Fe, rel, tw, tb, v4(msg, Fe) || Fe == 0};
        EpiAtomic epi{dWe, Fe + T, nullptr, nullptr, 0};
        rc = launch(A, B, epi, D, nullptr, Fe + T, E, E_dev, pick_splits(D, Fe + T, e_typ), stream, e_typ);
        if (rc) return rc;
    }
    if (T > 0 && dTenc) {
        ARowMajor A{dEP, D, v4(dEP, D)};
        BRowMajor B{We + Fe, Fe + T, v4(We + Fe, Fe + T)};
        EpiStore epi{dTenc, T, nullptr, nullptr, nullptr, 0};
        rc = launch(A, B, epi, E, E_dev, T, D, nullptr, 1, stream);
        if (rc) return rc;
    }
    return 0;
}
