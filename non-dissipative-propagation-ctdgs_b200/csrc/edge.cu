// Edge-message path and anti-symmetric update (forward + backward), plus the small element-wise
// kernels around them.
//
// Replaces:
//   CTANEmbedding.forward delta-t normalisation            models/gnn_layers.py:94-95
//   PyG TransformerConv.message + utils.softmax + scatter   (call site models/gnn_layers.py:82)
//   PyG AntiSymmetricConv.forward update                    (call site models/gnn_layers.py:86,98)
//   final_act                                               models/ctdg_models.py:457-458
//
// The edge list arrives CSR-sorted by destination (the loader emits it that way), so one warp owns
// one destination row: attention logits, the segmented softmax and the weighted aggregation are
// warp-shuffle reductions over the row's <= S edges -- no atomics, no [E,D] intermediates in HBM --
// and the tanh update x <- x + eps*tanh(x A^T + phi + b) is applied in the same pass.
#include "common.cuh"
#include <cstdlib>

// rel[e] = ((|last_update[n_id[src_local[e]]] - t_e| - mean) / std) as fp32   (gnn_layers.py:94-95)
// n_id == nullptr: last_update is already the gathered per-local-node array.
__global__ void edge_rel_kernel(const int64_t* __restrict__ last_update, const int64_t* __restrict__ n_id,
                                const int64_t* __restrict__ edge_src, const int64_t* __restrict__ t_tab,
                                const int64_t* __restrict__ eidx, int E_host, const int* __restrict__ E_dev,
                                float mean, float stdv, float* __restrict__ rel) {
    ctan_pdl_begin();
    const int E = E_dev ? *E_dev : E_host;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < E; e += gridDim.x * blockDim.x) {
        const int64_t lu = last_update[n_id ? n_id[edge_src[e]] : edge_src[e]];
        const int64_t te = t_tab[eidx ? eidx[e] : (int64_t)e];
        long long d = lu - te;
        if (d < 0) d = -d;
        rel[e] = __fdiv_rn(__ll2float_rn(d) - mean, stdv);
    }
}

extern "C" int ctan_edge_rel(const int64_t* last_update, const int64_t* n_id, const int64_t* edge_src,
                             const int64_t* t_tab, const int64_t* eidx, int E, const int* E_dev, float mean,
                             float stdv, float* rel, cudaStream_t stream) {
    if (E <= 0) return 0;
    CTAN_LAUNCH((edge_rel_kernel), ctan_grid(E, 256), 256, 0, stream, last_update, n_id, edge_src, t_tab, eidx, E, E_dev, mean,
                                                          stdv, rel);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// z0[r] = [ mem[n_id[r]] | xfeat[n_id[r]] ]   (memory_layers.py:155-156, ctdg_models.py:444-453)
__global__ void gather_rows2_kernel(const float* __restrict__ t1, int K1, const float* __restrict__ t2, int K2,
                                    const int64_t* __restrict__ idx, int N_host, const int* __restrict__ N_dev,
                                    float* __restrict__ out) {
    ctan_pdl_begin();
    const int N = N_dev ? *N_dev : N_host;
    const int ld = K1 + K2;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < N; r += nw) {
        const int64_t g = idx[r];
        for (int d = lane; d < K1; d += 32) out[(size_t)r * ld + d] = __ldg(t1 + (size_t)g * K1 + d);
        for (int d = lane; d < K2; d += 32) out[(size_t)r * ld + K1 + d] = __ldg(t2 + (size_t)g * K2 + d);
    }
}

extern "C" int ctan_gather_rows2(const float* t1, int K1, const float* t2, int K2, const int64_t* idx, int N,
                                 const int* N_dev, float* out, cudaStream_t stream) {
    if (N <= 0) return 0;
    CTAN_LAUNCH((gather_rows2_kernel), ctan_grid((long long)N * 32, 256), 256, 0, stream, t1, K1, t2, K2, idx, N, N_dev, out);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// A = W - W^T - gamma*I   (AntiSymmetricConv.forward, first line)
__global__ void antisym_prep_kernel(const float* __restrict__ W, float gamma, int D, float* __restrict__ A) {
    ctan_pdl_begin();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D * D; i += gridDim.x * blockDim.x) {
        const int a = i / D, b = i % D;
        A[i] = W[i] - W[b * D + a] - (a == b ? gamma : 0.f);
    }
}
extern "C" int ctan_antisym_prep(const float* W, float gamma, int D, float* A, cudaStream_t stream) {
    CTAN_LAUNCH((antisym_prep_kernel), ctan_grid(D * D, 256), 256, 0, stream, W, gamma, D, A);
    CTAN_CHECK_LAUNCH();
    return 0;
}
// dW = dA - dA^T
__global__ void antisym_bwd_kernel(const float* __restrict__ dA, int D, float* __restrict__ dW) {
    ctan_pdl_begin();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D * D; i += gridDim.x * blockDim.x) {
        const int a = i / D, b = i % D;
        dW[i] = dA[i] - dA[b * D + a];
    }
}
extern "C" int ctan_antisym_bwd(const float* dA, int D, float* dW, cudaStream_t stream) {
    CTAN_LAUNCH((antisym_bwd_kernel), ctan_grid(D * D, 256), 256, 0, stream, dA, D, dW);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// K4+K5 forward: one warp per destination node.
//   QKVA [N,4D] = [q | k | v | x A^T + b]   (produced by the projection contraction)
//   EP   [E,D]  = lin_edge(edge_attr)        (iteration-invariant, computed once)
//   s_e = q_i.(k_j + EP_e)/sqrt(D);  alpha = softmax_i(s) with +1e-16;  phi_i = sum alpha (v_j + EP_e)
//   x_out = x_in + eps * tanh(QKVA[:,3D:] + phi)
// Saves ALPHA [E] (normalised attention) and TH [N,D] (tanh values) for the backward when non-null.
// ------------------------------------------------------------------------------------------
// One pass over edges e = e_begin, e_begin + stride, ... < e_end of destination row i (every k_j / v_j / EP row is
// read from HBM exactly once): online softmax -- running max mx, running denominator den and running weighted
// sum phi are rescaled by exp(mx_old - mx_new) whenever the max grows.  Software-pipelined: the rows of the next
// edge are requested before the current one is reduced.  Raw scores are parked in ALPHA (lane (e-r0)%32).
template <int MAXV>
__device__ __forceinline__ void edge_partial(const float* __restrict__ QKVA, const float* __restrict__ EP,
                                             const int64_t* __restrict__ esrc, int D, int ld, int lane,
                                             const float (&q)[MAXV], int e_begin, int e_end, int stride, int r0,
                                             float* __restrict__ ALPHA, float inv_sqrt_d, float& mx, float& den,
                                             float (&phi)[MAXV]) {
    if (e_begin >= e_end) return;
    float kj[MAXV], vj[MAXV], ep[MAXV];
    {
        const int64_t j = esrc[e_begin];
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int d = lane + 32 * v;
            const bool in = d < D;
            kj[v] = in ? QKVA[(size_t)j * ld + D + d] : 0.f;
            vj[v] = in ? QKVA[(size_t)j * ld + 2 * D + d] : 0.f;
            ep[v] = in ? EP[(size_t)e_begin * D + d] : 0.f;
        }
    }
    for (int e = e_begin; e < e_end; e += stride) {
        float kn[MAXV], vn[MAXV], en[MAXV];
        const int e1 = (e + stride < e_end) ? e + stride : e;          // last iteration re-reads its own rows (L1 hit)
        const int64_t jn = esrc[e1];
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int d = lane + 32 * v;
            const bool in = d < D;
            kn[v] = in ? QKVA[(size_t)jn * ld + D + d] : 0.f;
            vn[v] = in ? QKVA[(size_t)jn * ld + 2 * D + d] : 0.f;
            en[v] = in ? EP[(size_t)e1 * D + d] : 0.f;
        }
        float part = 0.f;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) part = fmaf(q[v], kj[v] + ep[v], part);
        const float s = warp_sum(part) * inv_sqrt_d;
        if (ALPHA && lane == ((e - r0) & 31)) ALPHA[e] = s;
        const float nmx = fmaxf(mx, s);
        const float scale = (mx == -INFINITY) ? 0.f : expf(mx - nmx);
        const float p = expf(s - nmx);
        den = fmaf(den, scale, p);
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            phi[v] = fmaf(phi[v], scale, p * (vj[v] + ep[v]));
            kj[v] = kn[v]; vj[v] = vn[v]; ep[v] = en[v];
        }
        mx = nmx;
    }
}

// x_out = x_in + eps tanh(x A^T + b + phi) for one (row, column)
__device__ __forceinline__ void edge_finish(const float* __restrict__ QKVA, const float* __restrict__ Xin, float eps,
                                            float* __restrict__ Xout, float* __restrict__ TH,
                                            float* __restrict__ Zout, int D, int i, int d, float phi_d) {
    const float h = QKVA[(size_t)i * 4 * D + 3 * D + d] + phi_d;
    const float th = tanhf(h);
    if (TH) TH[(size_t)i * D + d] = th;
    const float xn = fmaf(eps, th, Xin[(size_t)i * D + d]);
    Xout[(size_t)i * D + d] = xn;
    if (Zout) Zout[(size_t)i * D + d] = tanhf(xn);
}

// Rows of degree > HUB_DEG at small batches are a chain of dependent gathers in ONE warp (a degree-32 row of the
// TGB evaluation took ~20 us: the whole launch waited for it).  When the caller provides a hub workspace and the
// launch is in the latency regime (N <= HUB_MAX_N), such rows are appended to a list instead and processed by
// edge_fwd_hub_kernel: one CTA per hub row, each of its 8 warps taking every 8th edge with its own online-softmax
// state, the 8 states merged through shared memory in a fixed order.  Bandwidth-bound launches (large N) and
// callers without a workspace keep the plain warp-per-row path for every row.
// hub workspace (int32): [0] = number of deferred rows, [1] = CTA ticket of the hub kernel, [2..] = row ids.
constexpr int HUB_DEG = 8, HUB_MAX_N = 16384, EF_WARPS = 8;

template <int MAXV>
__global__ void __launch_bounds__(256, MAXV <= 8 ? 3 : 1)               // 3 CTAs/SM: the saturating shape needs the warps
edge_fwd_kernel(const float* __restrict__ QKVA, const float* __restrict__ EP,
                                const int32_t* __restrict__ rowptr, const int64_t* __restrict__ esrc, int N_host,
                                const int* __restrict__ N_dev, int D, const float* __restrict__ Xin, float eps,
                                float* __restrict__ Xout, float* __restrict__ TH, float* __restrict__ ALPHA,
                                float* __restrict__ Zout, float inv_sqrt_d, int* __restrict__ hub, int hub_cap, int hub_mode) {
    ctan_pdl_begin();
    const int N = N_dev ? *N_dev : N_host;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int ld = 4 * D;
    const bool defer = hub != nullptr && N <= HUB_MAX_N;
    for (int i = warp; i < N; i += nw) {
        const int r0 = rowptr[i], r1 = rowptr[i + 1];
        if (defer && r1 - r0 > HUB_DEG) {
            if (hub_mode == 1) continue;                                // listed by ctan_hub_list: the hub kernel's row
            int slot = 0;
            if (lane == 0) slot = atomicAdd(hub, 1);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < hub_cap) {
                if (lane == 0) hub[2 + slot] = i;
                continue;
            }                                                           // list full: plain path
        }
        float phi[MAXV];
#pragma unroll
        for (int v = 0; v < MAXV; ++v) phi[v] = 0.f;
        if (r1 > r0) {
            float q[MAXV];
#pragma unroll
            for (int v = 0; v < MAXV; ++v) {
                const int d = lane + 32 * v;
                q[v] = d < D ? QKVA[(size_t)i * ld + d] : 0.f;
            }
            float mx = -INFINITY, den = 0.f;
            edge_partial<MAXV>(QKVA, EP, esrc, D, ld, lane, q, r0, r1, 1, r0, ALPHA, inv_sqrt_d, mx, den, phi);
            const float inv = 1.f / (den + 1e-16f);
#pragma unroll
            for (int v = 0; v < MAXV; ++v) phi[v] *= inv;
            if (ALPHA) {
                for (int e = r0 + lane; e < r1; e += 32) ALPHA[e] = expf(ALPHA[e] - mx) * inv;   // own write, own read
            }
        }
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int d = lane + 32 * v;
            if (d < D) edge_finish(QKVA, Xin, eps, Xout, TH, Zout, D, i, d, phi[v]);
        }
    }
}

template <int MAXV>
__global__ void __launch_bounds__(32 * EF_WARPS)
edge_fwd_hub_kernel(const float* __restrict__ QKVA, const float* __restrict__ EP, const int32_t* __restrict__ rowptr,
                    const int64_t* __restrict__ esrc, int D, const float* __restrict__ Xin, float eps,
                    float* __restrict__ Xout, float* __restrict__ TH, float* __restrict__ ALPHA,
                    float* __restrict__ Zout, float inv_sqrt_d, int* __restrict__ hub, int hub_cap, int cleanup) {
    ctan_pdl_begin();
    __shared__ float s_mx[EF_WARPS], s_den[EF_WARPS];
    __shared__ float s_phi[EF_WARPS][MAXV * 32];
    __shared__ int s_nh;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int ld = 4 * D;
    if (threadIdx.x == 0) s_nh = min(hub[0], hub_cap);
    __syncthreads();
    const int nh = s_nh;
    for (int hidx = blockIdx.x; hidx < nh; hidx += gridDim.x) {
        const int i = hub[2 + hidx];
        const int r0 = rowptr[i], r1 = rowptr[i + 1];
        float q[MAXV], phi[MAXV];
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int d = lane + 32 * v;
            q[v] = d < D ? QKVA[(size_t)i * ld + d] : 0.f;
            phi[v] = 0.f;
        }
        float mx = -INFINITY, den = 0.f;
        edge_partial<MAXV>(QKVA, EP, esrc, D, ld, lane, q, r0 + wib, r1, EF_WARPS, r0, ALPHA, inv_sqrt_d, mx, den, phi);
#pragma unroll
        for (int v = 0; v < MAXV; ++v) s_phi[wib][lane + 32 * v] = phi[v];
        if (lane == 0) { s_mx[wib] = mx; s_den[wib] = den; }
        __syncthreads();
        float M = s_mx[0];
#pragma unroll
        for (int w = 1; w < EF_WARPS; ++w) M = fmaxf(M, s_mx[w]);
        float sc[EF_WARPS], dsum = 0.f;
#pragma unroll
        for (int w = 0; w < EF_WARPS; ++w) {                            // fixed merge order: deterministic
            sc[w] = s_mx[w] == -INFINITY ? 0.f : expf(s_mx[w] - M);
            dsum = fmaf(s_den[w], sc[w], dsum);
        }
        const float inv = 1.f / (dsum + 1e-16f);
        for (int d = threadIdx.x; d < D; d += 32 * EF_WARPS) {
            float p = 0.f;
#pragma unroll
            for (int w = 0; w < EF_WARPS; ++w) p = fmaf(s_phi[w][d], sc[w], p);
            edge_finish(QKVA, Xin, eps, Xout, TH, Zout, D, i, d, p * inv);
        }
        if (ALPHA)                                                      // raw scores were written by other warps of this CTA
            for (int e = r0 + threadIdx.x; e < r1; e += 32 * EF_WARPS) ALPHA[e] = expf(ALPHA[e] - M) * inv;
        __syncthreads();
    }
    // the last CTA to finish clears the list for the next launch (every CTA has read hub[0] before its ticket)
    if (cleanup && threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(hub + 1, 1) == (int)gridDim.x - 1) { hub[0] = 0; hub[1] = 0; }
    }
}

// The hub list straight from the CSR pointer (same rule as edge_fwd_kernel: N <= HUB_MAX_N, degree > HUB_DEG), so
// that a caller can run the plain rows (hub_mode 1) and the hub rows (hub_mode 2) of ctan_edge_fwd CONCURRENTLY on two
// streams instead of one after the other.  hub[0] must be zero on entry; row order in the list is irrelevant (one
// CTA per row, results do not depend on it).
__global__ void hub_list_kernel(const int32_t* __restrict__ rowptr, int N_host, const int* __restrict__ N_dev,
                                int* __restrict__ hub, int hub_cap) {
    ctan_pdl_begin();
    const int N = N_dev ? *N_dev : N_host;
    if (N > HUB_MAX_N) return;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        if (rowptr[i + 1] - rowptr[i] > HUB_DEG) {
            const int slot = atomicAdd(hub, 1);
            if (slot < hub_cap) hub[2 + slot] = i;
        }
    }
}
extern "C" int ctan_hub_list(const int32_t* rowptr, int N, const int* N_dev, int32_t* hub_ws, int hub_cap,
                             cudaStream_t stream) {
    if (N <= 0 || !hub_ws || hub_cap <= 0) return 0;
    // every row of degree > HUB_DEG must fit the list: ctan_edge_fwd's hub_mode 1 skips them all, so a dropped row
    // would be computed by neither kernel
    if (hub_cap < (N < HUB_MAX_N ? N : HUB_MAX_N)) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((hub_list_kernel), ctan_grid(N, 256), 256, 0, stream, rowptr, N, N_dev, hub_ws, hub_cap);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// hub_ws: optional int32 workspace of 2 + hub_cap entries; NULL = every row on the plain path.
// hub_mode 0: self-contained -- the plain-row kernel builds the list, the hub kernel consumes and clears it ([0] and [1]
//             must be zero before the first call; the call leaves them zero again);
// hub_mode 1: plain rows only, rows listed by ctan_hub_list are skipped;  hub_mode 2: listed hub rows only.
//             (1 and 2 touch disjoint rows: issue them on two streams.)
extern "C" int ctan_edge_fwd(const float* QKVA, const float* EP, const int32_t* rowptr, const int64_t* esrc, int N,
                             const int* N_dev, int D, const float* Xin, float eps, float* Xout, float* TH,
                             float* ALPHA, float* Zout, int32_t* hub_ws, int hub_cap, int hub_mode,
                             cudaStream_t stream) {
    if (N <= 0) return 0;
    if (D > 512) return CTAN_ERR_UNSUPPORTED;
    if (hub_ws && hub_cap <= 0) hub_ws = nullptr;
    if (hub_mode != 0 && (!hub_ws || hub_cap < (N < HUB_MAX_N ? N : HUB_MAX_N))) return CTAN_ERR_BAD_ARG;   // (see ctan_hub_list)
    const int grid = ctan_grid((long long)N * 32, 256);
    const float isd = 1.0f / sqrtf((float)D);
    const int hgrid = 2 * CTAN_SM_COUNT;
#define LAUNCH(V) do {                                                                                                 \
        if (hub_mode != 2)                                                                                             \
            CTAN_LAUNCH((edge_fwd_kernel<V>), grid, 256, 0, stream, QKVA, EP, rowptr, esrc, N, N_dev, D, Xin, eps, Xout, \
                        TH, ALPHA, Zout, isd, hub_ws, hub_cap, hub_mode);                                              \
        if (hub_ws && hub_mode != 1)                                                                                   \
            CTAN_LAUNCH((edge_fwd_hub_kernel<V>), hgrid, 32 * EF_WARPS, 0, stream, QKVA, EP, rowptr, esrc, D, Xin, eps,  \
                        Xout, TH, ALPHA, Zout, isd, hub_ws, hub_cap, hub_mode == 0 ? 1 : 0);                           \
    } while (0)
    if (D <= 32) LAUNCH(1); else if (D <= 64) LAUNCH(2); else if (D <= 128) LAUNCH(4);
    else if (D <= 256) LAUNCH(8); else LAUNCH(16);
#undef LAUNCH
    CTAN_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// K4+K5 backward: one warp per destination node.  G = dL/dx_out [N,D].
//   dh = eps*(1-th^2)*G          -> DQKVA[:,3D:4D]   (gradient of x A^T + b, and of phi)
//   dalpha_e = dh.(v_j+EP_e); ds_e = alpha_e (dalpha_e - sum alpha dalpha)
//   dq_i = sum ds_e (k_j+EP_e)/sqrt(D)               -> DQKVA[:,0:D]
//   dk_j += ds_e q_i/sqrt(D); dv_j += alpha_e dh      -> atomics into DQKVA[:,D:3D] (source rows are not sorted)
//   dEP_e (+)= alpha_e dh + ds_e q_i/sqrt(D)          (edge owned by this warp: no atomics)
// DQKVA must be zero on entry.  DS [E] is scratch.
// ------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void edge_bwd_kernel(const float* __restrict__ QKVA, const float* __restrict__ EP,
                                const int32_t* __restrict__ rowptr, const int64_t* __restrict__ esrc, int N_host,
                                const int* __restrict__ N_dev, int D, const float* __restrict__ G,
                                const float* __restrict__ TH, const float* __restrict__ ALPHA, float eps,
                                float* __restrict__ DQKVA, float* __restrict__ DEP, int accumulate_dep,
                                float* __restrict__ DS, float inv_sqrt_d) {
    ctan_pdl_begin();
    const int N = N_dev ? *N_dev : N_host;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const int ld = 4 * D;
    for (int i = warp; i < N; i += nw) {
        float dh[MAXV];
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int d = lane + 32 * v;
            dh[v] = 0.f;
            if (d < D) {
                const float th = TH[(size_t)i * D + d];
                dh[v] = eps * (1.f - th * th) * G[(size_t)i * D + d];
                DQKVA[(size_t)i * ld + 3 * D + d] = dh[v];
            }
        }
        const int r0 = rowptr[i], r1 = rowptr[i + 1];
        if (r1 <= r0) continue;
        float q[MAXV], dq[MAXV];
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int d = lane + 32 * v;
            q[v] = d < D ? QKVA[(size_t)i * ld + d] : 0.f;
            dq[v] = 0.f;
        }
        float tmp = 0.f;                                  // sum_e alpha_e * dalpha_e
        for (int c0 = r0; c0 < r1; c0 += 32) {
            const int cnt = min(32, r1 - c0);
            float keep = 0.f;
            for (int c = 0; c < cnt; ++c) {
                const int e = c0 + c;
                const int64_t j = esrc[e];
                float part = 0.f;
#pragma unroll
                for (int v = 0; v < MAXV; ++v) {
                    const int d = lane + 32 * v;
                    if (d < D) part = fmaf(dh[v], QKVA[(size_t)j * ld + 2 * D + d] + EP[(size_t)e * D + d], part);
                }
                const float da = warp_sum(part);
                if (lane == c) keep = da;
            }
            if (lane < cnt) {
                DS[c0 + lane] = keep;
                tmp = fmaf(ALPHA[c0 + lane], keep, tmp);
            }
        }
        tmp = warp_sum(tmp);
        for (int c0 = r0; c0 < r1; c0 += 32) {
            const int cnt = min(32, r1 - c0);
            float a_keep = 0.f, ds_keep = 0.f;
            if (lane < cnt) {
                a_keep = ALPHA[c0 + lane];
                ds_keep = a_keep * (DS[c0 + lane] - tmp);  // own write above, own read here
            }
            for (int c = 0; c < cnt; ++c) {
                const int e = c0 + c;
                const int64_t j = esrc[e];
                const float a = __shfl_sync(0xffffffffu, a_keep, c);
                const float kq = __shfl_sync(0xffffffffu, ds_keep, c) * inv_sqrt_d;
#pragma unroll
                for (int v = 0; v < MAXV; ++v) {
                    const int d = lane + 32 * v;
                    if (d < D) {
                        const float ep = EP[(size_t)e * D + d];
                        dq[v] = fmaf(kq, QKVA[(size_t)j * ld + D + d] + ep, dq[v]);
                        const float gk = kq * q[v];
                        const float gv = a * dh[v];
                        atomicAdd(DQKVA + (size_t)j * ld + D + d, gk);
                        atomicAdd(DQKVA + (size_t)j * ld + 2 * D + d, gv);
                        const float dep = gk + gv;
                        float* o = DEP + (size_t)e * D + d;
                        *o = accumulate_dep ? (*o + dep) : dep;
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
            const int d = lane + 32 * v;
            if (d < D) DQKVA[(size_t)i * ld + d] = dq[v];
        }
    }
}

extern "C" int ctan_edge_bwd(const float* QKVA, const float* EP, const int32_t* rowptr, const int64_t* esrc, int N,
                             const int* N_dev, int D, const float* G, const float* TH, const float* ALPHA, float eps,
                             float* DQKVA, float* DEP, int accumulate_dep, float* DS, cudaStream_t stream) {
    if (N <= 0) return 0;
    if (D > 512) return CTAN_ERR_UNSUPPORTED;
    const int grid = ctan_grid((long long)N * 32, 256);
    const float isd = 1.0f / sqrtf((float)D);
#define LAUNCH(V) CTAN_LAUNCH((edge_bwd_kernel<V>), grid, 256, 0, stream, QKVA, EP, rowptr, esrc, N, N_dev, D, G, TH, ALPHA, eps, \
                                                                DQKVA, DEP, accumulate_dep, DS, isd)
    if (D <= 32) LAUNCH(1); else if (D <= 64) LAUNCH(2); else if (D <= 128) LAUNCH(4);
    else if (D <= 256) LAUNCH(8); else LAUNCH(16);
#undef LAUNCH
    CTAN_CHECK_LAUNCH();
    return 0;
}

// dX = dZ * (1 - Z^2)     (final_act = tanh, ctdg_models.py:457-458)
__global__ void tanh_bwd_kernel(const float* __restrict__ dZ, const float* __restrict__ Z, long long n_host,
                                const int* __restrict__ rows_dev, int D, float* __restrict__ dX) {
    ctan_pdl_begin();
    const long long n = rows_dev ? (long long)(*rows_dev) * D : n_host;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float z = Z[i];
        dX[i] = dZ[i] * (1.f - z * z);
    }
}
extern "C" int ctan_tanh_bwd(const float* dZ, const float* Z, int N, const int* N_dev, int D, float* dX,
                             cudaStream_t stream) {
    if (N <= 0) return 0;
    CTAN_LAUNCH((tanh_bwd_kernel), ctan_grid((long long)N * D, 256), 256, 0, stream, dZ, Z, (long long)N * D, N_dev, D, dX);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// Time-encoder parameter gradients from dTenc[E,T] = dEP . We[:,Fe:]:
//   d/dw_tau = sum_e -sin(w_tau rel_e + b_tau) rel_e dTenc[e,tau];   d/db_tau = sum_e -sin(.) dTenc[e,tau]
__global__ void tenc_bwd_kernel(const float* __restrict__ dTenc, const float* __restrict__ rel,
                                const float* __restrict__ tw, const float* __restrict__ tb, int E_host,
                                const int* __restrict__ E_dev, int T, float* __restrict__ dtw, float* __restrict__ dtb,
                                int slab) {
    ctan_pdl_begin();
    // block = 256 threads = G groups x Tc columns; a block reduces a slab of `slab` edges (32 for batch-sized edge sets:
    // two dependent loads per thread instead of eight; 128 for large ones: fewer atomics per output)
    __shared__ float s_w[256], s_b[256];
    const int E = E_dev ? *E_dev : E_host;
    const int e0 = blockIdx.x * slab, e1 = min(E, e0 + slab);
    if (e0 >= E) return;
    const int Tc = min(T, 256);
    const int G = 256 / Tc;
    const int grp = threadIdx.x / Tc, col = threadIdx.x % Tc;
    for (int tau0 = 0; tau0 < T; tau0 += Tc) {
        const int tau = tau0 + col;
        float sw = 0.f, sb = 0.f;
        if (grp < G && tau < T) {
            const float w = tw[tau], b = tb[tau];
            for (int e = e0 + grp; e < e1; e += G) {
                const float r = rel[e];
                const float g = -sinf(fmaf(r, w, b)) * dTenc[(size_t)e * T + tau];
                sw = fmaf(g, r, sw);
                sb += g;
            }
        }
        s_w[threadIdx.x] = sw;
        s_b[threadIdx.x] = sb;
        __syncthreads();
        if (grp == 0 && tau < T) {
            for (int g2 = 1; g2 < G; ++g2) { sw += s_w[g2 * Tc + col]; sb += s_b[g2 * Tc + col]; }
            atomicAdd(dtw + tau, sw);
            atomicAdd(dtb + tau, sb);
        }
        __syncthreads();
    }
}
extern "C" int ctan_tenc_bwd(const float* dTenc, const float* rel, const float* tw, const float* tb, int E,
                             const int* E_dev, int T, float* dtw, float* dtb, cudaStream_t stream) {
    if (E <= 0 || T <= 0) return 0;
    const int slab = E <= 16384 ? 32 : 128;
    CTAN_LAUNCH((tenc_bwd_kernel), ctan_div_up(E, slab), 256, 0, stream, dTenc, rel, tw, tb, E, E_dev, T, dtw, dtb, slab);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// The same gradients straight from dEP (no dTenc round trip): dTenc[e,tau] = sum_d dEP[e,d] We[d, Fe+tau] is formed in
// registers, one warp per edge (the time columns of We staged once per CTA as Wt[tau][d]), and reduced like above.
// For batch-sized edge sets this replaces a [E,D]x[D,T] contraction launch plus ctan_tenc_bwd by one short kernel.
__global__ void __launch_bounds__(256)
tenc_bwd_fused_kernel(const float* __restrict__ dEP, const float* __restrict__ We, int Fe, int T, int D,
                      const float* __restrict__ rel, const float* __restrict__ tw, const float* __restrict__ tb,
                      int E_host, const int* __restrict__ E_dev, float* __restrict__ dtw, float* __restrict__ dtb) {
    ctan_pdl_begin();
    extern __shared__ float tsm[];
    float* Wt = tsm;                                          // [T][D]
    float* red = tsm + (size_t)T * D;                         // [8 warps][2][T]
    const int E = E_dev ? *E_dev : E_host;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * 8 + warp, nw = gridDim.x * 8;
    if (blockIdx.x * 8 >= E) return;
    const int K = Fe + T;
    for (int i = threadIdx.x; i < T * D; i += 256) {          // coalesced along tau within a weight row
        const int d = i / T, tau = i - d * T;
        Wt[(size_t)tau * D + d] = We[(size_t)d * K + Fe + tau];
    }
    for (int i = threadIdx.x; i < 16 * T; i += 256) red[i] = 0.f;
    __syncthreads();
    const int nq = D >> 5;                                    // D % 32 == 0, D <= 256
    for (int t0 = 0; t0 < T; t0 += 32) {                      // lane <-> time channel t0 + lane
        const int tau = t0 + lane;
        const float w = tau < T ? tw[tau] : 0.f, b = tau < T ? tb[tau] : 0.f;
        float sw = 0.f, sb = 0.f;
        for (int e = gw; e < E; e += nw) {
            float x[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = q < nq ? dEP[(size_t)e * D + lane + 32 * q] : 0.f;
            const float r = rel[e];
            float mine = 0.f;
            const int tn = min(32, T - t0);
            for (int j = 0; j < tn; ++j) {
                const float* wr = Wt + (size_t)(t0 + j) * D + lane;
                float p = 0.f;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (q < nq) p = fmaf(x[q], wr[32 * q], p);
                p = warp_sum(p);
                if (j == lane) mine = p;
            }
            if (tau < T) {
                const float g = -sinf(fmaf(r, w, b)) * mine;
                sw = fmaf(g, r, sw);
                sb += g;
            }
        }
        if (tau < T) { red[(warp * 2 + 0) * T + tau] = sw; red[(warp * 2 + 1) * T + tau] = sb; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * T; i += 256) {
        const int which = i / T, tau = i - which * T;
        float v = 0.f;
        for (int wv = 0; wv < 8; ++wv) v += red[(wv * 2 + which) * T + tau];
        atomicAdd((which ? dtb : dtw) + tau, v);
    }
}
// Requirements: D % 32 == 0, D <= 256, (T*D + 16*T) floats of shared memory (<= 96 KB); else CTAN_ERR_UNSUPPORTED
// (the caller keeps ctan_eproj_bwd(.., dTenc) + ctan_tenc_bwd).
extern "C" int ctan_tenc_bwd_fused(const float* dEP, const float* We, int Fe, int T, int D, const float* rel,
                                   const float* tw, const float* tb, int E, const int* E_dev, float* dtw, float* dtb,
                                   cudaStream_t stream) {
    if (E <= 0 || T <= 0) return 0;
    const size_t smem = ((size_t)T * D + 16 * (size_t)T) * sizeof(float);
    if (D % 32 != 0 || D > 256 || smem > 96 * 1024) return CTAN_ERR_UNSUPPORTED;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(tenc_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    int grid = ctan_div_up(E, 16);                            // ~2 edges per warp at the bound
    if (grid > 2 * CTAN_SM_COUNT) grid = 2 * CTAN_SM_COUNT;
    CTAN_LAUNCH((tenc_bwd_fused_kernel), grid, 256, smem, stream, dEP, We, Fe, T, D, rel, tw, tb, E, E_dev, dtw, dtb);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// rowptr from a destination-sorted edge list that did not come from our loader.  flag[0] |= 1 if unsorted.
__global__ void csr_from_sorted_kernel(const int64_t* __restrict__ edst, int E, int N, int32_t* __restrict__ rowptr,
                                       int* __restrict__ flag) {
    ctan_pdl_begin();
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e <= E; e += gridDim.x * blockDim.x) {
        const long long prev = e > 0 ? edst[e - 1] : -1;
        const long long cur = e < E ? edst[e] : (long long)N;
        if (cur < prev || cur > N || prev >= N) { *flag = 1; continue; }
        for (long long r = prev + 1; r <= cur; ++r) rowptr[r] = e;
    }
}
extern "C" int ctan_csr_from_sorted(const int64_t* edst, int E, int N, int32_t* rowptr, int* flag,
                                    cudaStream_t stream) {
    if (N < 0 || E < 0) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((csr_from_sorted_kernel), ctan_grid(E + 1, 256), 256, 0, stream, edst, E, N, rowptr, flag);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// Operand preparation of the tensor-core enc_x path: one warp per local node r (global id g = n_id[r])
//   zmain[r,:] = mem[g,:]                                  (dense, 16-byte aligned rows -> TMA source)
//   tail[r,n]  = sum_f xfeat[g,f] * Wenc[n, D+f]           (the node-feature columns of enc_x, a rank-Fn update
//                                                           that the contraction's epilogue adds as a residual)
__global__ void enc_prep_kernel(const float* __restrict__ mem, int D, const float* __restrict__ xfeat, int Fn,
                                const int64_t* __restrict__ n_id, int N_host, const int* __restrict__ N_dev,
                                const float* __restrict__ Wenc, float* __restrict__ zmain, float* __restrict__ tail) {
    ctan_pdl_begin();
    // the node-feature columns of enc_x.weight ([D, D+Fn], column D+f of row d) are strided by D+Fn floats: staged
    // once per CTA as Wt[f][d] so that the per-row tail is a coalesced multiply-add (was 32 sectors per warp load)
    extern __shared__ float Wt[];
    const int ldw = D + Fn;
    for (int i = threadIdx.x; i < Fn * D; i += blockDim.x) {
        const int f = i / D, d = i - f * D;
        Wt[i] = __ldg(Wenc + (size_t)d * ldw + D + f);
    }
    __syncthreads();
    const int N = N_dev ? *N_dev : N_host;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const bool vec = (D & 3) == 0 && ((((uintptr_t)mem | (uintptr_t)zmain | (uintptr_t)tail) & 15) == 0);
    for (int r = warp; r < N; r += nw) {
        const int64_t g = n_id[r];
        if (vec) {
            for (int d4 = lane; d4 < (D >> 2); d4 += 32) {
                const float4 m = __ldg(reinterpret_cast<const float4*>(mem + (size_t)g * D) + d4);
                reinterpret_cast<float4*>(zmain + (size_t)r * D)[d4] = m;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int f = 0; f < Fn; ++f) {
                    const float x = __ldg(xfeat + (size_t)g * Fn + f);
                    const float4 w4 = *reinterpret_cast<const float4*>(Wt + f * D + 4 * d4);
                    t.x = fmaf(x, w4.x, t.x); t.y = fmaf(x, w4.y, t.y); t.z = fmaf(x, w4.z, t.z); t.w = fmaf(x, w4.w, t.w);
                }
                reinterpret_cast<float4*>(tail + (size_t)r * D)[d4] = t;
            }
        } else {
            for (int d = lane; d < D; d += 32) {
                zmain[(size_t)r * D + d] = __ldg(mem + (size_t)g * D + d);
                float t = 0.f;
                for (int f = 0; f < Fn; ++f) t = fmaf(__ldg(xfeat + (size_t)g * Fn + f), Wt[f * D + d], t);
                tail[(size_t)r * D + d] = t;
            }
        }
    }
}
extern "C" int ctan_enc_prep(const float* mem, int D, const float* xfeat, int Fn, const int64_t* n_id, int N,
                             const int* N_dev, const float* Wenc, float* zmain, float* tail, cudaStream_t stream) {
    if (N <= 0) return 0;
    const size_t smem = (size_t)(Fn > 0 ? Fn : 1) * D * sizeof(float);
    if (smem > 48 * 1024) return CTAN_ERR_UNSUPPORTED;
    CTAN_LAUNCH((enc_prep_kernel), ctan_grid((long long)N * 32, 256, 4), 256, smem, stream, mem, D, xfeat, Fn, n_id, N, N_dev,
                Wenc, zmain, tail);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// Tensor-core lin_edge (evaluation shapes): edge_attr = [msg[e_id] | cos(w*rel + b)] (models/gnn_layers.py:96-97,
// TGB/modules/time_enc.py:23-24) written as a dense row-major [E, Kpad] operand (Kpad = Fe+T rounded up to 32, zero
// padded, 128-byte rows -> TMA source) for ctan_gemm_tc.  One warp per edge.
__global__ void eproj_prep_kernel(const float* __restrict__ msg, int Fe, const int64_t* __restrict__ eid,
                                  const float* __restrict__ rel, const float* __restrict__ tw,
                                  const float* __restrict__ tb, int T, int E_host, const int* __restrict__ E_dev,
                                  int Kpad, float* __restrict__ A) {
    ctan_pdl_begin();
    const int E = E_dev ? *E_dev : E_host;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int e = warp; e < E; e += nw) {
        const int64_t g = eid ? eid[e] : e;
        const float r = rel[e];
        float* out = A + (size_t)e * Kpad;
        for (int k = lane; k < Kpad; k += 32) {
            float v = 0.f;
            if (k < Fe) v = __ldg(msg + (size_t)g * Fe + k);
            else if (k < Fe + T) v = cosf(fmaf(r, __ldg(tw + (k - Fe)), __ldg(tb + (k - Fe))));
            out[k] = v;
        }
    }
}
extern "C" int ctan_eproj_prep(const float* msg, int Fe, const int64_t* eid, const float* rel, const float* tw,
                               const float* tb, int T, int E, const int* E_dev, int Kpad, float* A,
                               cudaStream_t stream) {
    if (E <= 0) return 0;
    if (Kpad < Fe + T || Kpad % 32 != 0) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((eproj_prep_kernel), ctan_grid((long long)E * 32, 256), 256, 0, stream, msg, Fe, eid, rel, tw, tb, T, E, E_dev,
                                                                           Kpad, A);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// Readout head on pre-projected node rows (evaluation): Y[N, 2*Hd] = z [Wsrc;Wdst]^T + [bsrc;bdst] is computed
// once per NODE on the tensor cores; a scored pair only gathers two rows:
//   OUT[p] = relu(Y[row(src_p), :Hd] + Y[row(dst_p), Hd:]) . Wfin + bfin         (models/predictors.py:16-19)
__global__ void head_gather_kernel(const float* __restrict__ Y, int Hd, const int64_t* __restrict__ src,
                                   const int64_t* __restrict__ dst, const int64_t* __restrict__ map, int P,
                                   const float* __restrict__ Wfin, const float* __restrict__ bfin,
                                   float* __restrict__ OUT) {
    ctan_pdl_begin();
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < P; p += nw) {
        int64_t rs = src[p], rd = dst[p];
        if (map) { rs = map[rs]; rd = map[rd]; }
        const float* ys = Y + (size_t)rs * 2 * Hd;
        const float* yd = Y + (size_t)rd * 2 * Hd + Hd;
        float s = 0.f;
        for (int h = lane; h < Hd; h += 32) s = fmaf(fmaxf(ys[h] + yd[h], 0.f), __ldg(Wfin + h), s);
        s = warp_sum(s);
        if (lane == 0) OUT[p] = s + __ldg(bfin);
    }
}
extern "C" int ctan_head_gather(const float* Y, int Hd, const int64_t* src, const int64_t* dst, const int64_t* map,
                                int P, const float* Wfin, const float* bfin, float* OUT, cudaStream_t stream) {
    if (P <= 0) return 0;
    CTAN_LAUNCH((head_gather_kernel), ctan_grid((long long)P * 32, 256), 256, 0, stream, Y, Hd, src, dst, map, P, Wfin, bfin, OUT);
    CTAN_CHECK_LAUNCH();
    return 0;
}
