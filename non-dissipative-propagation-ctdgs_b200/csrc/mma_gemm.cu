// Small-M contractions of the per-batch step on warp-level tensor-core MMAs (mma.sync m16n8k8, TF32 inputs, FP32
// accumulate), 3xTF32 error compensation for fp32 parity.
//
// Replaces: enc_x on gathered memory rows (models/gnn_layers.py:96), the
// q/k/v projections and x.A^T of TransformerConv / AntiSymmetricConv (PyG 2.3.1, App. A.5), and the gradient of those
// projections w.r.t. the node state (autograd in the reference, train_link.py:279).
//
// Why a third contraction path: at the reference batch (B = 200) these are [~700 x 128] x [128 x 512] problems --
// 0.1 GFLOP.  The tcgen05/TMA kernel (tc_gemm.cu) has ~8 us of fixed cost per launch (TMEM allocation, mbarrier /
// tensor-map set-up, first TMA round trip, TMEM drain) and the CUDA-core kernel (gemm.cuh) needs ~12 us of FMA
// issue; both are an order of magnitude above the work.  mma.sync has no set-up: a CTA loads its 64x64 tile operands
// with cp.async (double-buffered 32-float k-blocks), each of its 4 warps owns a 32x32 sub-tile, and the whole K=128
// contraction is 16 k-steps x 24 MMAs per warp.  North star: "tensor cores ... once the batch is large enough, and CUDA
// cores otherwise" -- the engines pick this kernel below ~2 k rows and the tcgen05 kernel above (measured crossover in
// DESIGN.md S4).
//
// 3xTF32: x = hi + lo with hi = tf32(x), lo = tf32(x - hi); acc += lo_a*hi_b + hi_a*lo_b + hi_a*hi_b (small terms
// first).  Both operands are split in registers, so no pre-split weight panels are needed.
#include "common.cuh"

namespace mm {

constexpr int BM = 64, BN = 64, BK = 32, THREADS = 128;
constexpr int LDA_S = BK + 4;        // smem row stride of K-contiguous tiles: fragment loads hit 32 distinct banks
constexpr int LDB_NN = BN + 8;       // smem row stride of the N-contiguous B tile (same property for its fragments)
constexpr int A_TILE = BM * LDA_S;                                  // floats
constexpr int B_TILE = (BN * LDA_S > BK * LDB_NN) ? BN * LDA_S : BK * LDB_NN;
constexpr int SMEM_BYTES = 2 * (A_TILE + B_TILE) * 4;              // two stages

struct Params {
    const float* A; int lda; const int64_t* a_idx;      // A[M,K] row-major; a_idx != NULL: row m is A[a_idx[m]]
    int M; const int* M_dev; int K;
    const float* B[4]; int ldb; int seg; int b_nn;      // b_nn = 0: B_s is [seg (rows along N), K], K contiguous ("x W^T")
                                                        // b_nn = 1: B_s is [seg (rows along K), Nout], N contiguous ("g W")
    int Nout;
    const float* bias[4]; int bseg;                     // bias of column n: bias[n / bseg][n % bseg] (NULL entries: none)
    const float* res; int ldr;                          // residual added once: C += res[m][n]
    const float* xf; int Fn; const float* xw; int ldxw; // rank-Fn tail: C[m][n] += sum_f xf[row(m)*Fn+f] * xw[n*ldxw+f]
    float* C; int ldc; int n_split; int single_pass;    // n_split > 1: atomicAdd into a zeroed C
    int b_scalar;                                       // B rows not 16-byte aligned (enc_x: ld = D + F_n): 4-byte copies
};

__device__ __forceinline__ void cp16(void* smem, const void* g, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = pred ? 16 : 0;                          // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(n));
}
__device__ __forceinline__ void cp4(void* smem, const void* g, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = pred ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(g), "r"(n));
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned tf32(float x) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ void split(float x, unsigned& hi, unsigned& lo) {
    hi = tf32(x);
    lo = tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(THREADS) gemm_mma_kernel(const Params p) {
    ctan_pdl_begin();
    extern __shared__ __align__(16) float smem[];
    const int M = p.M_dev ? min(*p.M_dev, p.M) : p.M;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    if (m0 >= M) return;
    const int nkb_all = p.K / BK;
    const int per = (nkb_all + p.n_split - 1) / p.n_split;
    const int kb0 = blockIdx.z * per, kb1 = min(nkb_all, kb0 + per);
    if (kb0 >= kb1) return;
    float* As[2] = {smem, smem + A_TILE + B_TILE};
    float* Bs[2] = {smem + A_TILE, smem + 2 * A_TILE + B_TILE};
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;
    const int g = lane >> 2, t = lane & 3;

    // ---- per-thread load assignments (fixed across k-blocks)
    // A: 64 rows x 8 float4 -> 4 per thread
    const float* a_src[4]; bool a_ok[4]; int a_dst[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int v = tid + i * THREADS, r = v >> 3, c4 = v & 7;
        const int m = m0 + r;
        a_ok[i] = m < M;
        const int64_t row = a_ok[i] ? (p.a_idx ? p.a_idx[m] : (int64_t)m) : 0;
        a_src[i] = p.A + (size_t)row * p.lda + c4 * 4;
        a_dst[i] = r * LDA_S + c4 * 4;
    }
    auto load = [&](int stage, int kb) {
        const int k0 = kb * BK;
#pragma unroll
        for (int i = 0; i < 4; ++i) cp16(As[stage] + a_dst[i], a_src[i] + k0, a_ok[i]);
        if (!p.b_nn && p.b_scalar) {                     // same tile, element-wise (rows not 16-byte aligned)
#pragma unroll 4
            for (int i = 0; i < 16; ++i) {
                const int v = tid + i * THREADS, r = v >> 5, c = v & 31;
                const int n = n0 + r;
                const bool ok = n < p.Nout;
                const int s = ok ? n / p.seg : 0, rr = ok ? n - s * p.seg : 0;
                cp4(Bs[stage] + r * LDA_S + c, p.B[s] + (size_t)rr * p.ldb + k0 + c, ok);
            }
        } else if (!p.b_nn) {                            // B tile [BN rows (n)][BK], K contiguous
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int v = tid + i * THREADS, r = v >> 3, c4 = v & 7;
                const int n = n0 + r;
                const bool ok = n < p.Nout;
                const int s = ok ? n / p.seg : 0, rr = ok ? n - s * p.seg : 0;
                cp16(Bs[stage] + r * LDA_S + c4 * 4, p.B[s] + (size_t)rr * p.ldb + k0 + c4 * 4, ok);
            }
        } else {                                         // B tile [BK rows (k)][BN], N contiguous
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int v = tid + i * THREADS, r = v >> 4, c4 = v & 15;
                const int k = k0 + r, n = n0 + c4 * 4;
                const int s = k / p.seg, rr = k - s * p.seg;
                cp16(Bs[stage] + r * LDB_NN + c4 * 4, p.B[s] + (size_t)rr * p.ldb + n, n < p.Nout);
            }
        }
        cp_commit();
    };

    float acc[2][4][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[i][j][k] = 0.f;

    load(0, kb0);
    for (int kb = kb0; kb < kb1; ++kb) {
        const int st = (kb - kb0) & 1;
        if (kb + 1 < kb1) { load(st ^ 1, kb + 1); cp_wait<1>(); } else { cp_wait<0>(); }
        __syncthreads();
        const float* as = As[st];
        const float* bs = Bs[st];
#pragma unroll
        for (int kk = 0; kk < BK; kk += 8) {
            unsigned ahi[2][4], alo[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float* ap = as + (wm + mt * 16 + g) * LDA_S + kk + t;
                split(ap[0], ahi[mt][0], alo[mt][0]);
                split(ap[8 * LDA_S], ahi[mt][1], alo[mt][1]);
                split(ap[4], ahi[mt][2], alo[mt][2]);
                split(ap[8 * LDA_S + 4], ahi[mt][3], alo[mt][3]);
            }
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                float b0f, b1f;
                if (!p.b_nn) {
                    const float* bp = bs + (wn + nt * 8 + g) * LDA_S + kk + t;
                    b0f = bp[0]; b1f = bp[4];
                } else {
                    const float* bp = bs + (kk + t) * LDB_NN + wn + nt * 8 + g;
                    b0f = bp[0]; b1f = bp[4 * LDB_NN];
                }
                unsigned bh0, bl0, bh1, bl1;
                split(b0f, bh0, bl0);
                split(b1f, bh1, bl1);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    if (!p.single_pass) {
                        mma(acc[mt][nt], alo[mt], bh0, bh1);
                        mma(acc[mt][nt], ahi[mt], bl0, bl1);
                    }
                    mma(acc[mt][nt], ahi[mt], bh0, bh1);
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue: bias / residual / node-feature tail once (split 0), then store or accumulate
    const bool first = blockIdx.z == 0;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = m0 + wm + mt * 16 + g + half * 8;
            if (m >= M) continue;
            const int64_t xrow = (p.xf && first) ? (p.a_idx ? p.a_idx[m] : (int64_t)m) : 0;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int n = n0 + wn + nt * 8 + 2 * t;
                if (n >= p.Nout) continue;
                float v0 = acc[mt][nt][half * 2], v1 = acc[mt][nt][half * 2 + 1];
                if (first) {
                    if (p.bseg > 0) {
                        const int s = n / p.bseg, r = n - s * p.bseg;
                        if (p.bias[s]) { v0 += __ldg(p.bias[s] + r); v1 += __ldg(p.bias[s] + r + 1); }
                    }
                    if (p.res) { v0 += p.res[(size_t)m * p.ldr + n]; v1 += p.res[(size_t)m * p.ldr + n + 1]; }
                    for (int f = 0; f < p.Fn; ++f) {
                        const float x = p.xf[xrow * p.Fn + f];
                        v0 = fmaf(x, __ldg(p.xw + (size_t)n * p.ldxw + f), v0);
                        v1 = fmaf(x, __ldg(p.xw + (size_t)(n + 1) * p.ldxw + f), v1);
                    }
                }
                float* c = p.C + (size_t)m * p.ldc + n;
                if (p.n_split > 1) atomicAdd(reinterpret_cast<float2*>(c), make_float2(v0, v1));   // one 8-byte red
                else *reinterpret_cast<float2*>(c) = make_float2(v0, v1);
            }
        }
}

}  // namespace mm

// C[M,Nout] (+)= A[M,K] . B + bias + res + node-feature tail   (3xTF32 on mma.sync; mode != 0: single-pass TF32, the 1e-2 tier)
//   A: row-major, leading dimension lda, optional row gather a_idx (int64 global ids);
//   B: up to four segments B0..B3 of `seg` rows each, leading dimension ldb --
//        b_nn = 0: segment s holds output columns [s*seg, (s+1)*seg), each row K-contiguous   (C = A W^T, torch Linear)
//        b_nn = 1: segment s holds contraction rows [s*seg, (s+1)*seg), each row Nout-contiguous (C = A W)
//   bias: column n gets bias[n / bseg][n % bseg] (bseg = 0: none);  res: C += res[m*ldr+n];
//   tail: C[m][n] += sum_{f<Fn} xf[row(m)*Fn+f] * xw[n*ldxw+f]  (the node-feature columns of enc_x; row(m) = a_idx[m]);
//   n_split > 1 splits K over gridDim.z and accumulates with atomicAdd: C must be zero on entry.
// Requirements: K % 32 == 0, Nout even, 16-byte aligned rows of A (and of B when b_nn = 1; W^T-form weights with
// unaligned rows, e.g. enc_x's [D, D+F_n], are copied element-wise),
// seg % 32 == 0 (b_nn = 1) -- otherwise CTAN_ERR_UNSUPPORTED and the caller uses the CUDA-core / tcgen05 path.
extern "C" int ctan_gemm_mma(const float* A, int lda, const int64_t* a_idx, int M, const int* M_dev, int K,
                             const float* B0, const float* B1, const float* B2, const float* B3, int ldb, int seg,
                             int b_nn, int Nout, const float* bias0, const float* bias1, const float* bias2,
                             const float* bias3, int bseg, const float* res, int ldr, const float* xf, int Fn,
                             const float* xw, int ldxw, float* C, int ldc, int n_split, int mode, cudaStream_t stream) {
    if (M <= 0) return 0;
    if (K <= 0 || Nout <= 0 || seg <= 0 || !A || !B0 || !C) return CTAN_ERR_BAD_ARG;
    const int nseg = b_nn ? (K + seg - 1) / seg : (Nout + seg - 1) / seg;
    const float* Bs[4] = {B0, B1, B2, B3};
    if (nseg > 4) return CTAN_ERR_BAD_ARG;
    uintptr_t bits = (uintptr_t)A | (uintptr_t)C, bbits = 0;
    for (int s = 0; s < nseg; ++s) {
        if (!Bs[s]) return CTAN_ERR_BAD_ARG;
        bbits |= (uintptr_t)Bs[s];
    }
    const bool b_scalar = (ldb & 3) || (bbits & 15);     // weight rows not 16-byte aligned: allowed for b_nn = 0 only
    if (K % mm::BK != 0 || (Nout & 1) || (lda & 3) || (ldc & 1) || (bits & 15) || (bbits & 3) || (b_scalar && b_nn) ||
        (b_nn && (seg % mm::BK != 0 || (Nout & 3))) || (res && (ldr & 1)) || (bseg & 1))
        return CTAN_ERR_UNSUPPORTED;
    if (n_split < 1) n_split = 1;
    if (n_split > K / mm::BK) n_split = K / mm::BK;
    mm::Params p;
    p.A = A; p.lda = lda; p.a_idx = a_idx; p.M = M; p.M_dev = M_dev; p.K = K;
    for (int s = 0; s < 4; ++s) p.B[s] = s < nseg ? Bs[s] : B0;
    p.ldb = ldb; p.seg = seg; p.b_nn = b_nn; p.Nout = Nout;
    p.bias[0] = bias0; p.bias[1] = bias1; p.bias[2] = bias2; p.bias[3] = bias3; p.bseg = bseg;
    p.res = res; p.ldr = ldr; p.xf = Fn > 0 ? xf : nullptr; p.Fn = xf ? Fn : 0; p.xw = xw; p.ldxw = ldxw;
    p.C = C; p.ldc = ldc; p.n_split = n_split; p.single_pass = mode != 0; p.b_scalar = b_scalar;
    static_assert(mm::SMEM_BYTES <= 48 * 1024, "fits the default dynamic shared-memory limit");
    dim3 grid(ctan_div_up(Nout, mm::BN), ctan_div_up(M, mm::BM), n_split);
    CTAN_LAUNCH((mm::gemm_mma_kernel), grid, mm::THREADS, mm::SMEM_BYTES, stream, p);
    CTAN_CHECK_LAUNCH();
    return 0;
}
