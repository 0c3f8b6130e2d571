// Tensor-core projection for the large-batch shapes: QKVA[N,4D] = X [Wq;Wk;Wv;A]^T + bias on the 5th-gen
// tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM) with operands staged by TMA
// (cp.async.bulk.tensor, 128-byte swizzle) -- the "once the batch is large enough" path of the north star.
//
// fp32 parity: 3xTF32 error compensation.  Each fp32 operand is split as x = hi + lo with hi = x truncated
// to TF32 (10-bit mantissa) and lo = x - hi; the product is accumulated as lo*hi' + hi*lo' + hi*hi' in the
// fp32 TMEM accumulator (relative error ~2^-21, inside the 1e-4 bar; plain TF32 would be ~5e-4).
// The weights are split once per step by ctan_wcat_prep (which also forms A = W - W^T - gamma I); the
// activation tile is split by the transform warps right after its TMA load.
//
// One persistent CTA per SM walks 128x128 output tiles: 320 threads, warp-specialised (TMA producer / MMA issuer /
// 4 transform warps / 4 epilogue warps).  K advances in 128-byte (32-float) blocks through a 4-stage ring of
// mbarriers; the activation operand is split in registers and handed to the MMA through TENSOR MEMORY (only the
// weight tile is read from shared memory: a tf32 MMA is shared-memory-bandwidth bound); two TMEM accumulators let the
// epilogue of tile i (tcgen05.ld -> smem staging -> +bias/+residual -> full-line stores) overlap the MMAs of tile
// i+1.  Replaces the same reference code as ctan_qkva_fwd.
#include "common.cuh"
#include <cuda.h>

// Phase timeline of the tcgen05 kernel (profiling builds only: -DCTAN_TC_TRACE; scripts/micro_tc3.py): one elected
// thread per role stores %globaltimer at the phase boundaries of the first 64 CTAs.
#ifdef CTAN_TC_TRACE
__device__ unsigned long long g_tc_trace[64 * 32];
#define TC_TRACE(slot)                                                                                         \
    do {                                                                                                       \
        const unsigned _cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);                   \
        if (_cta < 64) { unsigned long long _t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));         \
                         g_tc_trace[_cta * 32 + (slot)] = _t; }                                                \
    } while (0)
extern "C" int ctan_tc_trace_read(unsigned long long* host_out, int clear) {
    cudaError_t e = cudaMemcpyFromSymbol(host_out, g_tc_trace, sizeof(g_tc_trace));
    if (e == cudaSuccess && clear) {
        static unsigned long long z[64 * 32];
        e = cudaMemcpyToSymbol(g_tc_trace, z, sizeof(z));
    }
    return (int)e;
}
#else
#define TC_TRACE(slot) do {} while (0)
#endif

namespace tc {

constexpr int BM = 128, BN = 128, BKF = 32, STAGES = 4;
constexpr int TILE_BYTES = BM * BKF * 4;                 // 16 KB (A and B tiles have the same shape)
constexpr int STAGE_BYTES = 3 * TILE_BYTES;              // A (raw fp32), B_hi, B_lo
constexpr int EPI_LD = 36;                               // staging row pitch in floats (32 + 4: conflict-free float4)
constexpr int EPI_BYTES = 4 * 32 * EPI_LD * 4;           // one 32x32 staging block per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int TGROUPS = 1;                               // transform warp groups (k-blocks alternate between them)
constexpr int THREADS = 64 + TGROUPS * 128 + 128;        // TMA | MMA | TGROUPS x 4 transform warps | 4 epilogue warps
// TMEM columns (512 allocated): two 128-column fp32 accumulators, then per stage 32 columns A_hi + 32 columns A_lo
constexpr uint32_t TMEM_COLS = 512, TM_ACC = 0, TM_A = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%1], %0;" ::"r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a lost arrival must never hang the GPU -- trap instead (the launch then fails loudly).
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int crd_inner,
                                            int crd_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "r"(crd_inner), "r"(crd_outer)
        : "memory");
}
// Same load, delivered to the same shared-memory offset of every CTA of the cluster named in cta_mask, each of whose
// mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int crd_inner,
                                               int crd_outer, uint16_t cta_mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)), "l"((uint64_t)tmap), "r"(smem_u32(bar)), "h"(cta_mask), "r"(crd_inner), "r"(crd_outer)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// K-major operand tile, 128-byte swizzle: rows 128 B apart, 8-row groups 1024 B apart (SBO), version 1.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::tf32, fp32 accumulate, A and B K-major, M=128, N=128
__device__ __forceinline__ uint32_t instr_desc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]: the activation operand is read from tensor memory (lane = row, one tf32 per
// column), so only the weight tile goes through the shared-memory read port.
__device__ __forceinline__ void mma_tf32_ta(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// kind::f16 with BF16 operands (both from shared memory), fp32 accumulate, K-major, M=128, N=128
__device__ __forceinline__ uint32_t instr_desc_bf16() {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {        // low half = a (element k), high half = b (k+1)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// the same arrival delivered to the mbarrier at this offset in every CTA of cta_mask (a stage whose B half was multicast
// by the peer is free only when BOTH CTAs' MMAs have read it)
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t* v) {      // caller waits (tcgen05.wait::ld)
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

struct Bias4 { const float* p[4]; int D; int vec; const float* res; int ldr; int cvec;   // + optional residual res[m][n]; cvec: C rows 16-byte aligned
               float* C2; int split; int ldc2; };   // columns >= split go to C2[m][n - split] (split % BN == 0: uniform per tile)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised, persistent over row tiles (one CTA per SM; the ring and its phases run on across tiles):
//   warp 0     TMA producer (one lane): A raw fp32, B_hi, B_lo tiles of a k-block -> smem stage s
//   warp 1     TMEM owner + MMA issuer (one lane)
//   warps 2-5  operand transform: read the raw A tile from smem, split x = hi + lo (3xTF32) in registers and
//              write both halves to TENSOR MEMORY (tcgen05.st) -- the MMA then reads A from TMEM and only B from
//              smem, which halves the shared-memory traffic that bounds a tf32 MMA (8 KB of operands / 67 clk)
//   warps 6-9  epilogue of the PREVIOUS tile (second accumulator buffer) while the next tile's MMAs run:
//              tcgen05.ld -> per-warp smem staging -> +bias/+residual -> full-line coalesced stores (or atomics)
// Barriers: raw[s] TMA landed | ready[s] A of stage s is in TMEM (128 arrivals) | empty[s] MMAs of stage s retired
//           done[b] accumulator b complete | accfree[b] accumulator b drained (128 arrivals)
// mode 0: 3xTF32 (fp32 parity path); mode 1: single TF32 pass (hi only);
// mode 2: BF16 (the 1e-2 path): k-blocks of 64 elements in a 3-stage ring of [A raw k 0..31 | A raw k 32..63 | A bf16 | B bf16]
//         tiles; the transform warps convert the fp32 activation tile to a BF16 tile in the canonical K-major 128-byte-
//         swizzled layout (what TMA would have produced) and the MMA (kind::f16) reads both operands from shared memory;
//         tmBhi is then the tensor map of a BF16 weight panel, tmBlo is unused.
// PAIR (mode 0 only): the kernel runs as clusters of two CTAs that share a column tile and take adjacent row tiles.  The
// weight tiles of a k-block are the same for both, so each CTA fetches HALF of them (rank 0: B_hi, rank 1: B_lo) and
// multicasts it into both CTAs' stage: 32 KB instead of 48 KB of L2 reads per CTA per k-block -- at evaluation shapes
// (all SMs busy) the ring is bound by exactly that traffic.  A stage is released by BOTH CTAs' MMAs (multicast commit,
// empty[s] counts two arrivals); the two CTAs walk the same number of (tile, k-block) steps -- the second one takes a
// dummy tile (rows >= M: nothing stored) when the row-tile count is odd.
template <bool PAIR>
__device__ __forceinline__ void tc_body(const CUtensorMap* tmA, const CUtensorMap* tmBhi, const CUtensorMap* tmBlo,
                                        const Bias4& bias, float* __restrict__ C, int ldc, int M_host,
                                        const int* __restrict__ M_dev, int N, int K, int mode) {
    ctan_pdl_begin();
    extern __shared__ uint8_t smem_raw[];
    const int M = M_dev ? *M_dev : M_host;
    const int n0 = blockIdx.x * BN;
    uint32_t crank = 0;
    if (PAIR) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
    const int y_lead = PAIR ? (int)(blockIdx.y & ~1u) : (int)blockIdx.y;   // first row tile of the cluster
    if (y_lead * BM >= M) return;                                     // uniform over the cluster, before any allocation
    if (threadIdx.x == 0) TC_TRACE(0);
    // 1024-byte alignment as an OFFSET from the __shared__ array (not an integer round trip of the pointer): the
    // compiler keeps the shared address space, so the staging / transform accesses compile to LDS/STS, not generic LD/ST
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float* epi = (float*)(smem + STAGES * STAGE_BYTES);
    uint64_t* raw = (uint64_t*)(smem + STAGES * STAGE_BYTES + EPI_BYTES);
    uint64_t* ready = raw + STAGES;
    uint64_t* empty = ready + STAGES;
    uint64_t* done = empty + STAGES;          // [2]
    uint64_t* accfree = done + 2;             // [2]
    uint32_t* tmem_holder = (uint32_t*)(accfree + 2);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;

    if (t == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&raw[s], 1); mbar_init(&ready[s], 128); mbar_init(&empty[s], PAIR ? 2 : 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&done[b], 1); mbar_init(&accfree[b], 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_holder)),
                     "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync_all();                                     // the peer's barriers exist before anything targets them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_holder;
    if (threadIdx.x == 0) TC_TRACE(1);
    // split-K over gridDim.z (atomic epilogue): thin problems (few row tiles, long K) get more CTAs
    const bool bf = mode == 2;
    const int bke = bf ? 2 * BKF : BKF;                                 // elements of K per k-block
    const int nst = bf ? 3 : STAGES;                                    // ring depth (bf16 stages are 64 KB)
    const int stage_bytes = bf ? 4 * TILE_BYTES : STAGE_BYTES;
    const int kb_per = (K / bke + (int)gridDim.z - 1) / (int)gridDim.z;
    const int kb0 = (int)blockIdx.z * kb_per;
    const int n_kb = min(K / bke - kb0, kb_per);
    auto stage_ptr = [&](int s, int which) { return smem + s * stage_bytes + which * TILE_BYTES; };

    if (warp == 0) {
        // ------------------------------------------------ TMA producer
        if (lane == 0) {
            int it = 0;                                                 // k-blocks issued so far (all tiles)
            for (int mL = y_lead * BM; mL < M; mL += gridDim.y * BM)
            for (int kb = 0; kb < n_kb; ++kb, ++it) {
                const int m0 = mL + (int)crank * BM;
                const int s = it % nst;
                if (it >= nst) mbar_wait(&empty[s], ((it / nst) - 1) & 1);
                if (bf) {
                    mbar_expect_tx(&raw[s], 3 * TILE_BYTES);
                    tma_load_2d(stage_ptr(s, 0), tmA, &raw[s], (kb0 + kb) * bke, m0);
                    tma_load_2d(stage_ptr(s, 1), tmA, &raw[s], (kb0 + kb) * bke + BKF, m0);
                    tma_load_2d(stage_ptr(s, 3), tmBhi, &raw[s], (kb0 + kb) * bke, n0);
                    continue;
                }
                if (PAIR) {                                             // own A tile + one half of B for both CTAs
                    mbar_expect_tx(&raw[s], 3 * TILE_BYTES);
                    tma_load_2d(stage_ptr(s, 0), tmA, &raw[s], (kb0 + kb) * BKF, m0);
                    if (crank == 0) tma_load_2d_mc(stage_ptr(s, 1), tmBhi, &raw[s], (kb0 + kb) * BKF, n0, (uint16_t)3);
                    else            tma_load_2d_mc(stage_ptr(s, 2), tmBlo, &raw[s], (kb0 + kb) * BKF, n0, (uint16_t)3);
                    continue;
                }
                mbar_expect_tx(&raw[s], (mode == 0 ? 3 : 2) * TILE_BYTES);
                tma_load_2d(stage_ptr(s, 0), tmA, &raw[s], (kb0 + kb) * BKF, m0);
                tma_load_2d(stage_ptr(s, 1), tmBhi, &raw[s], (kb0 + kb) * BKF, n0);
                if (mode == 0) tma_load_2d(stage_ptr(s, 2), tmBlo, &raw[s], (kb0 + kb) * BKF, n0);
                if (it == 0) TC_TRACE(2);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = instr_desc_tf32();
            int it = 0, tile = 0;
            for (int mL = y_lead * BM; mL < M; mL += gridDim.y * BM, ++tile) {
                const int buf = tile & 1;
                if (tile >= 2) {                                        // the epilogue has drained this accumulator
                    mbar_wait(&accfree[buf], ((tile >> 1) - 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                const uint32_t acc = tmem_base + TM_ACC + (uint32_t)buf * BN;
                for (int kb = 0; kb < n_kb; ++kb, ++it) {
                    const int s = it % nst;
                    mbar_wait(&ready[s], (it / nst) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (it == 0) TC_TRACE(4);
                    if (bf) {
                        const uint32_t idb = instr_desc_bf16();
                        const uint32_t a_bf = smem_u32(stage_ptr(s, 2)), b_bf = smem_u32(stage_ptr(s, 3));
#pragma unroll
                        for (int k = 0; k < 4; ++k)                     // UMMA_K = 16 bf16 = 32 bytes inside the swizzle atom
                            mma_bf16_ss(acc, smem_desc_sw128(a_bf + k * 32), smem_desc_sw128(b_bf + k * 32), idb,
                                        (kb | k) ? 1u : 0u);
                        mma_commit(&empty[s]);
                        continue;
                    }
                    const uint32_t a_hi = tmem_base + TM_A + (uint32_t)s * 64, a_lo = a_hi + 32;
                    const uint32_t b_hi = smem_u32(stage_ptr(s, 1)), b_lo = smem_u32(stage_ptr(s, 2));
#pragma unroll
                    for (int k = 0; k < BKF / 8; ++k) {                // UMMA_K = 8 tf32 = 32 bytes inside the swizzle atom
                        const uint32_t off = k * 32;
                        if (mode == 0) {
                            mma_tf32_ta(acc, a_lo + k * 8, smem_desc_sw128(b_hi + off), idesc, (kb | k) ? 1u : 0u);
                            mma_tf32_ta(acc, a_hi + k * 8, smem_desc_sw128(b_lo + off), idesc, 1u);
                            mma_tf32_ta(acc, a_hi + k * 8, smem_desc_sw128(b_hi + off), idesc, 1u);
                        } else {
                            mma_tf32_ta(acc, a_hi + k * 8, smem_desc_sw128(b_hi + off), idesc, (kb | k) ? 1u : 0u);
                        }
                    }
                    if (PAIR) mma_commit_mc(&empty[s], (uint16_t)3);   // ... in both CTAs: the peer's next multicast lands here too
                    else mma_commit(&empty[s]);                         // arrives when these MMAs have read stage s
                }
                mma_commit(&done[buf]);                                 // this tile's accumulator is complete
                TC_TRACE(5);
            }
        }
    } else if (warp < 2 + 4 * TGROUPS) {
        // ------------------------------------------------ transform: raw A (smem, swizzled) -> hi/lo in TMEM
        // TGROUPS > 1 lets groups of four warps take alternate k-blocks (measured: no gain, the ring is bound by
        // the TMA round trip, not by this chain).
        const int grp = (warp - 2) >> 2;
        const int q = warp & 3;                                         // TMEM lane quarter this warp may touch
        const int r = q * 32 + lane;                                    // tile row = TMEM lane
        const uint32_t tlane = (uint32_t)(q * 32) << 16;
        int it = 0;
        for (int mL = y_lead * BM; mL < M; mL += gridDim.y * BM)
        for (int kb = 0; kb < n_kb; ++kb, ++it) {
            if (it % TGROUPS != grp) continue;
            const int s = it % nst;
            mbar_wait(&raw[s], (it / nst) & 1);
            if (warp == 2 && lane == 0 && it < 12) TC_TRACE(16 + it);
            if (bf) {
                uint8_t* orow = stage_ptr(s, 2) + r * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {                           // output chunk j = elements k 8j .. 8j+7 of row r
                    const uint8_t* irow = stage_ptr(s, j >> 2) + r * 128;
                    const int c0 = (2 * j) & 7;
                    const float4 v0 = *reinterpret_cast<const float4*>(irow + ((c0 ^ (r & 7)) << 4));
                    const float4 v1 = *reinterpret_cast<const float4*>(irow + (((c0 + 1) ^ (r & 7)) << 4));
                    uint4 o;
                    o.x = pack_bf16x2(v0.x, v0.y); o.y = pack_bf16x2(v0.z, v0.w);
                    o.z = pack_bf16x2(v1.x, v1.y); o.w = pack_bf16x2(v1.z, v1.w);
                    *reinterpret_cast<uint4*>(orow + ((j ^ (r & 7)) << 4)) = o;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA
                mbar_arrive(&ready[s]);
                continue;
            }
            const uint8_t* arow = stage_ptr(s, 0) + r * 128;
            const uint32_t t_hi = tmem_base + tlane + TM_A + (uint32_t)s * 64;
#pragma unroll
            for (int c = 0; c < 8; c += 2) {                            // 16-byte chunk c of row r sits at chunk c ^ (r & 7)
                const float4 v0 = *reinterpret_cast<const float4*>(arow + (((c) ^ (r & 7)) << 4));
                const float4 v1 = *reinterpret_cast<const float4*>(arow + (((c + 1) ^ (r & 7)) << 4));
                const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                uint32_t h[8], l[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    h[j] = __float_as_uint(x[j]) & 0xFFFFE000u;
                    l[j] = __float_as_uint(x[j] - __uint_as_float(h[j]));
                }
                tmem_st8(t_hi + c * 4, h);
                if (mode == 0) tmem_st8(t_hi + 32 + c * 4, l);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&ready[s]);
        }
    } else {
        // ------------------------------------------------ epilogue (overlaps the next tile's main loop)
        const int q = warp & 3;                                         // TMEM lane quarter of this warp
        float* stg = epi + (warp - (2 + 4 * TGROUPS)) * 32 * EPI_LD;
        const bool first = blockIdx.z == 0;                             // bias / residual are added by one split only
        const int sub_r = lane >> 3, sub_c = (lane & 7) * 4;
        int tile = 0;
        for (int mL = y_lead * BM; mL < M; mL += gridDim.y * BM, ++tile) {
            const int m0 = mL + (int)crank * BM;                        // (PAIR, odd tile count: may lie past M -- nothing is stored)
            const int buf = tile & 1;
            // bias of this thread's 4 columns in each 32-column chunk: requested while the MMAs of the tile still run
            float4 bvs[BN / 32];
#pragma unroll
            for (int ci = 0; ci < BN / 32; ++ci) {
                bvs[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int n = n0 + ci * 32 + sub_c;
                if (first && n < N) {
                    const int seg = n / bias.D;
                    const float* bp = seg == 0 ? bias.p[0] : (seg == 1 ? bias.p[1] : (seg == 2 ? bias.p[2] : bias.p[3]));
                    if (bp) {
                        bp += n - seg * bias.D;
                        if (bias.vec) bvs[ci] = __ldg(reinterpret_cast<const float4*>(bp));
                        else bvs[ci] = make_float4(__ldg(bp), __ldg(bp + 1), __ldg(bp + 2), __ldg(bp + 3));
                    }
                }
            }
            const bool has_res = first && bias.res != nullptr;
            const int emode = (gridDim.z > 1 ? 2 : 0) + (bias.cvec ? 0 : 1);   // 0 store v4 | 1 store scalar | 2 red v4 | 3 red scalar
            uint32_t v[2][32];
            float4 rv[2][8];
            auto load_res = [&](int ci, float4 (&r)[8]) {               // residual rows of chunk ci (clamped, unconditional)
                const int n = n0 + ci * 32 + sub_c;
                const int nc = n < N ? n : N - 4;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    int row = m0 + q * 32 + i * 4 + sub_r;
                    row = row < M ? row : M - 1;
                    const float* rp = bias.res + (size_t)row * bias.ldr + nc;
                    if (bias.vec) r[i] = __ldg(reinterpret_cast<const float4*>(rp));
                    else r[i] = make_float4(__ldg(rp), __ldg(rp + 1), __ldg(rp + 2), __ldg(rp + 3));
                }
            };
            if (has_res) load_res(0, rv[0]);                         // first chunk's residual rows: requested while the MMAs run
            mbar_wait(&done[buf], (tile >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (q == 0 && lane == 0) TC_TRACE(6);
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + TM_ACC + (uint32_t)buf * BN;
            // Software-pipelined over the four 32-column chunks: the TMEM read AND the residual rows of chunk c+1 are in flight
            // while chunk c is staged and stored (biases were requested before the accumulator was complete); residual
            // addresses are clamped and unconditional, so a chunk's loads are one batch of independent requests, not a
            // load -> add -> store chain per row (measured: that chain made the residual epilogue of dX 10 us against 4 us
            // without residual; pipelined: 4.3 -> see scripts/micro_tc3.py).
            tmem_ld32_issue(tacc, v[0]);
#pragma unroll
            for (int ci = 0; ci < BN / 32; ++ci) {
                const int c0 = ci * 32;
                const int n = n0 + c0 + sub_c;                          // 4 columns: inside one bias segment (D % 16 == 0)
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");           // chunk ci is in registers
                if (q == 0 && lane == 0 && ci < 2) TC_TRACE(9 + 3 * ci);
                if (ci + 1 < BN / 32) {                                 // next chunk: TMEM read and residual rows in flight
                    tmem_ld32_issue(tacc + c0 + 32, v[(ci + 1) & 1]);
                    if (has_res) load_res(ci + 1, rv[(ci + 1) & 1]);
                }
                const uint32_t* vc = v[ci & 1];
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * EPI_LD + j) =
                        make_float4(__uint_as_float(vc[j]), __uint_as_float(vc[j + 1]), __uint_as_float(vc[j + 2]),
                                    __uint_as_float(vc[j + 3]));
                __syncwarp();
                if (q == 0 && lane == 0 && ci < 2) TC_TRACE(10 + 3 * ci);
                const float4 bv = bvs[ci];
                // all eight staged vectors first (one batch of LDS), then the stores: no load -> store chain per row and no
                // uniform-mode branches inside the row loop
                float4 o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    o[i] = *reinterpret_cast<const float4*>(stg + (i * 4 + sub_r) * EPI_LD + sub_c);
                    o[i].x += bv.x; o[i].y += bv.y; o[i].z += bv.z; o[i].w += bv.w;
                    if (has_res) {
                        const float4 r4 = rv[ci & 1][i];
                        o[i].x += r4.x; o[i].y += r4.y; o[i].z += r4.z; o[i].w += r4.w;
                    }
                }
                float* outb = (bias.C2 && n0 >= bias.split)
                                  ? bias.C2 + (size_t)(m0 + q * 32 + sub_r) * bias.ldc2 + (n - bias.split)
                                  : C + (size_t)(m0 + q * 32 + sub_r) * ldc + n;
                const int rows_left = M - (m0 + q * 32 + sub_r);        // row i is valid iff 4*i < rows_left
                const int ldo = (bias.C2 && n0 >= bias.split) ? bias.ldc2 : ldc;
                if (n < N) {
                    if (emode == 0) {                                   // plain 16-byte stores (full 128-byte lines per row)
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (4 * i < rows_left) *reinterpret_cast<float4*>(outb + (size_t)(4 * i) * ldo) = o[i];
                    } else if (emode == 2) {                            // split-K: one 16-byte red per vector
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (4 * i < rows_left) atomicAdd(reinterpret_cast<float4*>(outb + (size_t)(4 * i) * ldo), o[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            if (4 * i < rows_left) {
                                float* out = outb + (size_t)(4 * i) * ldo;
                                if (emode == 1) { out[0] = o[i].x; out[1] = o[i].y; out[2] = o[i].z; out[3] = o[i].w; }
                                else { atomicAdd(out, o[i].x); atomicAdd(out + 1, o[i].y); atomicAdd(out + 2, o[i].z); atomicAdd(out + 3, o[i].w); }
                            }
                    }
                }
                __syncwarp();
                if (q == 0 && lane == 0 && ci < 2) TC_TRACE(11 + 3 * ci);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            if (q == 0 && lane == 0) TC_TRACE(7);
            mbar_arrive(&accfree[buf]);                                 // TMEM reads of this tile are complete
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (PAIR) cluster_sync_all();             // no CTA leaves while its peer may still signal its barriers
    if (threadIdx.x == 0) TC_TRACE(8);
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

__global__ void __launch_bounds__(THREADS, 1)
qkva_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, Bias4 bias, float* __restrict__ C, int ldc, int M_host,
               const int* __restrict__ M_dev, int N, int K, int mode) {
    tc_body<false>(&tmA, &tmBhi, &tmBlo, bias, C, ldc, M_host, M_dev, N, K, mode);
}
// clusters of two CTAs along y (launched with cudaLaunchAttributeClusterDimension {1,2,1}; gridDim.y even); mode 0 only
__global__ void __launch_bounds__(THREADS, 1)
qkva_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                    const __grid_constant__ CUtensorMap tmBlo, Bias4 bias, float* __restrict__ C, int ldc, int M_host,
                    const int* __restrict__ M_dev, int N, int K, int mode) {
    tc_body<true>(&tmA, &tmBhi, &tmBlo, bias, C, ldc, M_host, M_dev, N, K, mode);
}

// NOTE: the tcgen05 kernel is launched the plain way (no programmatic dependent launch): one of its CTAs owns a whole
// SM (216 KB of shared memory, all of TMEM); scheduled early it would sit on that SM while it waits for its
// predecessor and keep the step's parallel branches off it (measured: +5 us per training step).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// row-major fp32 [rows, K] matrix, box = 128 rows x 32 floats, 128-byte swizzle
static int make_map(CUtensorMap* m, const float* base, int rows, int K) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return CTAN_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {(cuuint32_t)BKF, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : CTAN_ERR_BAD_ARG;
}

// row-major bf16 [rows, K] weight panel, box = 128 rows x 64 elements (128 bytes), 128-byte swizzle
static int make_map_bf16(CUtensorMap* m, const void* base, int rows, int K) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return CTAN_ERR_UNSUPPORTED;
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    cuuint32_t box[2] = {(cuuint32_t)(2 * BKF), (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : CTAN_ERR_BAD_ARG;
}

// grid.y: a CTA holds ~212 KB of smem and all 512 TMEM columns, so exactly one is resident per SM; launch at most
// one wave and let each CTA walk the remaining row tiles (that is also what overlaps epilogue and main loop)
// m_typ (optional, > 0): the caller's estimate of the TRUE row count when M_max is only a workspace bound.  Every CTA of
// this kernel needs a whole SM (213 KB of shared memory, all of TMEM), idle ones included, so a grid sized by a loose
// bound makes the launch wait for dozens of SMs to drain of the step's concurrent branches.  The kernel strides over
// row tiles, so a smaller grid is always correct.
static int bounded_rows(int M_max, const int* M_dev, int gx, int m_typ = 0) {
    (void)M_dev;
    if (m_typ > 0 && m_typ < M_max) M_max = m_typ;
    int gy = (M_max + BM - 1) / BM;
    const int cap = CTAN_SM_COUNT / gx;
    if (gy > cap) gy = cap > 1 ? cap : 1;
    return gy;
}

// Launch of the contraction kernel.  pair: clusters of two CTAs along y sharing the weight tiles by TMA multicast
// (`mode` bit 6 of the entry points below, 3xTF32 only); gridDim.y is made even for it.
static inline bool want_pair(int mode) { return (mode & 0x40) && (mode & 0x3F) == 0; }
static int launch(dim3 grid, bool pair, cudaStream_t stream, const CUtensorMap& tmA, const CUtensorMap& tmBhi,
                  const CUtensorMap& tmBlo, const Bias4& epi, float* C, int ldc, int M, const int* M_dev, int N, int K,
                  int mode) {
    if (!pair) {
        qkva_tc_kernel<<<grid, THREADS, SMEM_BYTES, stream>>>(tmA, tmBhi, tmBlo, epi, C, ldc, M, M_dev, N, K, mode);
        return 0;
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(qkva_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    if (grid.y & 1)                                                    // even: up if that still fits one wave, else down
        grid.y += ((grid.x * (grid.y + 1) * grid.z <= (unsigned)CTAN_SM_COUNT || grid.y == 1) ? 1 : -1);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 2; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, qkva_tc_pair_kernel, tmA, tmBhi, tmBlo, epi, C, ldc, M, M_dev, N, K, mode);
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace tc

// Wcat_hi/Wcat_lo [4D, D]: TF32 split of the stacked weights [Wq; Wk; Wv; A], A = W - W^T - gamma I.
// A_out [D,D] (optional) receives the un-split A for the CUDA-core backward.
__global__ void wcat_prep_kernel(const float* __restrict__ Wq, const float* __restrict__ Wk,
                                 const float* __restrict__ Wv, const float* __restrict__ W, float gamma, int D,
                                 float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ A_out,
                                 float* __restrict__ hiT, float* __restrict__ loT) {
    ctan_pdl_begin();
    const int n = 4 * D * D;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int seg = i / (D * D), r = i - seg * D * D;
        float x;
        if (seg == 0) x = Wq[r];
        else if (seg == 1) x = Wk[r];
        else if (seg == 2) x = Wv[r];
        else {
            const int a = r / D, b = r % D;
            x = W[r] - W[b * D + a] - (a == b ? gamma : 0.f);
            if (A_out) A_out[r] = x;
        }
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = x - h;
        if (hiT) {      // Wcat^T [D, 4D] (K-major operand of dX = DQKVA Wcat): element (c, seg*D + a) = Wcat[seg*D + a][c]
            const int a = r / D, c = r % D;
            const size_t o = (size_t)c * 4 * D + seg * D + a;
            hiT[o] = h;
            loT[o] = x - h;
        }
    }
}

extern "C" int ctan_wcat_prep(const float* Wq, const float* Wk, const float* Wv, const float* W, float gamma, int D,
                              float* Wcat_hi, float* Wcat_lo, float* A_out, float* WcatT_hi, float* WcatT_lo,
                              cudaStream_t stream) {
    if (D <= 0) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((wcat_prep_kernel), ctan_grid(4LL * D * D, 256), 256, 0, stream, Wq, Wk, Wv, W, gamma, D, Wcat_hi, Wcat_lo, A_out,
                                                                      WcatT_hi, WcatT_lo);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// QKVA[N,4D] = X Wcat^T + [bq;bk;bv;b] on tcgen05 (3xTF32).  Requirements: D % 32 == 0, X has at least
// ceil(N/128)*128 addressable rows or N rows (TMA zero-fills out-of-bounds rows), 16-byte aligned bases.
// Returns CTAN_ERR_UNSUPPORTED when the shape does not qualify (the caller then uses ctan_qkva_fwd).
// mode: bits 0-5 = precision (0 3xTF32, 1 TF32, 2 BF16); bit 6 = run as two-CTA clusters that share the weight tiles by TMA
// multicast (3xTF32 only; same results bit for bit); bits 8+ = optional estimate of the true row count when N is only a
// workspace bound (sizes the persistent grid; 0 = unknown).
extern "C" int ctan_qkva_fwd_tc(const float* X, int N, const int* N_dev, int D, const float* Wcat_hi,
                                const float* Wcat_lo, const float* bq, const float* bk, const float* bv,
                                const float* b, float* QKVA, int mode, cudaStream_t stream) {
    if (N <= 0) return 0;
    const int m_typ = mode >> 8;              // bits 8+: expected row count (0 = unknown), sizes the persistent grid only
    const bool pair = tc::want_pair(mode & 0xFF);
    mode &= 0x3F;
    const bool bf = mode == 2;                // Wcat_hi is then a BF16 [4D, D] panel (ctan_pack_bf16), Wcat_lo unused
    if (D % (bf ? 2 * tc::BKF : tc::BKF) != 0 || ((4 * D) % tc::BN) != 0 ||
        (((uintptr_t)X | (uintptr_t)Wcat_hi | (bf ? 0 : (uintptr_t)Wcat_lo)) & 15))
        return CTAN_ERR_UNSUPPORTED;
    CUtensorMap tmA, tmBhi, tmBlo;
    int rc = tc::make_map(&tmA, X, N, D);
    if (!rc) rc = bf ? tc::make_map_bf16(&tmBhi, Wcat_hi, 4 * D, D) : tc::make_map(&tmBhi, Wcat_hi, 4 * D, D);
    if (!rc) rc = bf ? tc::make_map_bf16(&tmBlo, Wcat_hi, 4 * D, D) : tc::make_map(&tmBlo, Wcat_lo, 4 * D, D);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(tc::qkva_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             tc::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    const uintptr_t ball = (uintptr_t)bq | (uintptr_t)bk | (uintptr_t)bv | (uintptr_t)b | (uintptr_t)QKVA;
    tc::Bias4 bias{{bq, bk, bv, b}, D, (ball & 15) == 0 ? 1 : 0, nullptr, 0, ((uintptr_t)QKVA & 15) == 0 ? 1 : 0, nullptr, 0, 0};
    dim3 grid((4 * D) / tc::BN, tc::bounded_rows(N, N_dev, (4 * D) / tc::BN, m_typ));
    { const int lrc = tc::launch(grid, pair, stream, tmA, tmBhi, tmBlo, bias, QKVA, 4 * D, N, N_dev, 4 * D, D,
                                                              mode); if (lrc) return lrc; }
    CTAN_CHECK_LAUNCH();
    return 0;
}

// dXin[N,D] = G + DQKVA[N,4D] Wcat   on tcgen05 (3xTF32): the activation-gradient contraction of one
// anti-symmetric iteration (ctan_qkva_bwd's first half), as A = DQKVA (K = 4D) against the K-major operand
// Wcat^T [D,4D] prepared by ctan_wcat_prep.  Plain stores: no atomics, dXin need not be zeroed.
// Requirements: D % 128 == 0 (one or more 128-wide output tiles); else CTAN_ERR_UNSUPPORTED.
extern "C" int ctan_dx_tc(const float* DQKVA, const float* G, int N, const int* N_dev, int D, const float* WcatT_hi,
                          const float* WcatT_lo, float* dXin, int mode, int n_split, cudaStream_t stream) {
    if (N <= 0) return 0;
    const int m_typ = mode >> 8;              // (see ctan_qkva_fwd_tc)
    const bool pair = tc::want_pair(mode & 0xFF);
    mode &= 0x3F;
    if (D % tc::BN != 0 || (((uintptr_t)DQKVA | (uintptr_t)WcatT_hi | (uintptr_t)WcatT_lo | (uintptr_t)G | (uintptr_t)dXin) & 15))
        return CTAN_ERR_UNSUPPORTED;
    CUtensorMap tmA, tmBhi, tmBlo;
    int rc = tc::make_map(&tmA, DQKVA, N, 4 * D);
    if (!rc) rc = tc::make_map(&tmBhi, WcatT_hi, D, 4 * D);
    if (!rc) rc = tc::make_map(&tmBlo, WcatT_lo, D, 4 * D);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(tc::qkva_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    tc::Bias4 epi{{nullptr, nullptr, nullptr, nullptr}, D, 1, G, D, 1, nullptr, 0, 0};       // (every base was checked 16-byte aligned above)
    // K = 4D is long and there are few row tiles: split K when the caller provides a zeroed dXin (n_split > 1)
    const int n_kb = (4 * D) / tc::BKF;
    int splits = n_split < 1 ? 1 : n_split;
    while (splits > 1 && (n_kb % splits != 0)) --splits;
    dim3 grid(D / tc::BN, tc::bounded_rows(N, N_dev, (D / tc::BN) * splits, m_typ), splits);
    { const int lrc = tc::launch(grid, pair, stream, tmA, tmBhi, tmBlo, epi, dXin, D, N, N_dev, D, 4 * D, mode); if (lrc) return lrc; }
    CTAN_CHECK_LAUNCH();
    return 0;
}

// TF32 split of a row-major [rows, cols] block read with leading dimension ld: packs (and 16-byte aligns) weights
// such as enc_x.weight[:, :D] (ld = D+Fn) or the readout's lin_src/lin_dst so that TMA can stage them.
__global__ void pack_split_kernel(const float* __restrict__ src, int rows, int cols, int ld, int ld_out,
                                  float* __restrict__ hi, float* __restrict__ lo) {
    ctan_pdl_begin();
    const int n = rows * ld_out;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / ld_out, c = i - r * ld_out;
        const float x = c < cols ? src[(size_t)r * ld + c] : 0.f;          // columns [cols, ld_out) are zero padding
        const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        hi[i] = h;
        lo[i] = x - h;
    }
}
extern "C" int ctan_pack_split(const float* src, int rows, int cols, int ld, float* hi, float* lo, cudaStream_t stream) {
    if (rows <= 0 || cols <= 0) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((pack_split_kernel), ctan_grid((long long)rows * cols, 256), 256, 0, stream, src, rows, cols, ld, cols, hi, lo);
    CTAN_CHECK_LAUNCH();
    return 0;
}
// same, into panels whose rows are zero-padded to ld_out >= cols floats (K rounded up to the 32-float k-block)
extern "C" int ctan_pack_split_pad(const float* src, int rows, int cols, int ld, int ld_out, float* hi, float* lo,
                                   cudaStream_t stream) {
    if (rows <= 0 || cols <= 0 || ld_out < cols) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((pack_split_kernel), ctan_grid((long long)rows * ld_out, 256), 256, 0, stream, src, rows, cols, ld, ld_out, hi, lo);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// BF16 weight panel of a row-major fp32 [rows, cols] block (leading dimension ld), rows zero-padded to ld_out elements
// (K rounded up to the 64-element k-block of the BF16 path).
__global__ void pack_bf16_kernel(const float* __restrict__ src, int rows, int cols, int ld, int ld_out,
                                 unsigned short* __restrict__ dst) {
    ctan_pdl_begin();
    const int n = rows * ld_out;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int r = i / ld_out, c = i - r * ld_out;
        const float x = c < cols ? src[(size_t)r * ld + c] : 0.f;
        dst[i] = (unsigned short)(tc::pack_bf16x2(x, 0.f) & 0xFFFFu);
    }
}
extern "C" int ctan_pack_bf16(const float* src, int rows, int cols, int ld, int ld_out, void* dst_bf16,
                              cudaStream_t stream) {
    if (rows <= 0 || cols <= 0 || ld_out < cols) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((pack_bf16_kernel), ctan_grid((long long)rows * ld_out, 256), 256, 0, stream, src, rows, cols, ld, ld_out,
                (unsigned short*)dst_bf16);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// Generic tensor-core linear layer: C[M,Nout] = A[M,K] B^T + bias_seg[n/seg][n%seg] + res[m][n]  (3xTF32 / TF32).
// A row-major [M(max),K] dense, B = packed split weights [Nout,K].  K % 32 == 0, Nout % 128 == 0, seg % 16 == 0.
extern "C" int ctan_gemm_tc(const float* A, int M, const int* M_dev, int K, const float* Bhi, const float* Blo, int Nout,
                            const float* b0, const float* b1, const float* b2, const float* b3, int seg,
                            const float* res, int ldr, float* C, int mode, cudaStream_t stream) {
    if (M <= 0) return 0;
    const int m_typ = mode >> 8;              // (see ctan_qkva_fwd_tc)
    const bool pair = tc::want_pair(mode & 0xFF);
    mode &= 0x3F;
    const bool bf = mode == 2;                // Bhi is then a BF16 [Nout, K] panel (ctan_pack_bf16), Blo unused
    if (K % (bf ? 2 * tc::BKF : tc::BKF) != 0 || Nout % tc::BN != 0 || seg <= 0 || seg % 16 != 0 ||
        (((uintptr_t)A | (uintptr_t)Bhi | (bf ? 0 : (uintptr_t)Blo) | (uintptr_t)C) & 15))
        return CTAN_ERR_UNSUPPORTED;
    CUtensorMap tmA, tmBhi, tmBlo;
    int rc = tc::make_map(&tmA, A, M, K);
    if (!rc) rc = bf ? tc::make_map_bf16(&tmBhi, Bhi, Nout, K) : tc::make_map(&tmBhi, Bhi, Nout, K);
    if (!rc) rc = bf ? tc::make_map_bf16(&tmBlo, Bhi, Nout, K) : tc::make_map(&tmBlo, Blo, Nout, K);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(tc::qkva_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    const uintptr_t ball = (uintptr_t)b0 | (uintptr_t)b1 | (uintptr_t)b2 | (uintptr_t)b3 | (uintptr_t)res |
                           (uintptr_t)(ldr * 4);
    tc::Bias4 epi{{b0, b1, b2, b3}, seg, (ball & 15) == 0 ? 1 : 0, res, ldr, 1, nullptr, 0, 0};   // (C was checked 16-byte aligned above, Nout % 128 == 0)
    dim3 grid(Nout / tc::BN, tc::bounded_rows(M, M_dev, Nout / tc::BN, m_typ));
    { const int lrc = tc::launch(grid, pair, stream, tmA, tmBhi, tmBlo, epi, C, Nout, M, M_dev, Nout, K, mode); if (lrc) return lrc; }
    CTAN_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------ enc_x folded into the projection
// With ONE anti-symmetric iteration the first thing the step does is two chained linear maps of the gathered rows:
//   X0 = z0 Wenc^T + benc ;  [q|k|v|x.A^T] = X0 Wcat^T + bcat     (models/gnn_layers.py:96, PyG App. A.5)
// Both are linear in z0, so with frozen weights (evaluation: tgb_test is @torch.no_grad) they are ONE contraction
//   [X0 | QKVA] = z0 [Wenc ; Wcat Wenc]^T + [benc ; Wcat benc + bcat]
// against a [5D, Kx] panel formed once per run.  z0 = [memory row | node features | 0-pad] with Kx = D+Fn rounded up to
// 32 (the node features become ordinary K columns: no residual pass).  One launch and one dependent stage fewer per batch.
__global__ void fuse_enc_kernel(const float* __restrict__ Wq, const float* __restrict__ Wk, const float* __restrict__ Wv,
                                const float* __restrict__ A, const float* __restrict__ bq, const float* __restrict__ bk,
                                const float* __restrict__ bv, const float* __restrict__ b, const float* __restrict__ Wenc,
                                const float* __restrict__ benc, int D, int Fn, int Kx, float* __restrict__ Wce,
                                float* __restrict__ bce) {
    const int ldw = D + Fn;
    const long long n = 5LL * D * Kx;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n + 5LL * D; i += (long long)gridDim.x * blockDim.x) {
        const bool is_bias = i >= n;
        const int r = is_bias ? (int)(i - n) : (int)(i / Kx), c = is_bias ? 0 : (int)(i - (long long)r * Kx);
        float v;
        if (r < D) {                                          // the enc_x block itself (zero padded to Kx)
            v = is_bias ? benc[r] : (c < ldw ? Wenc[(size_t)r * ldw + c] : 0.f);
        } else {                                              // row r-D of Wcat = [Wq;Wk;Wv;A] times enc_x
            const int rr = r - D, seg = rr / D, ro = rr - seg * D;
            const float* W = seg == 0 ? Wq : (seg == 1 ? Wk : (seg == 2 ? Wv : A));
            const float* bb = seg == 0 ? bq : (seg == 1 ? bk : (seg == 2 ? bv : b));
            double acc = 0.0;                                 // fp64 accumulation: the panel is formed once per run
            if (is_bias) {
                for (int d = 0; d < D; ++d) acc += (double)W[(size_t)ro * D + d] * (double)benc[d];
                acc += (double)bb[ro];
            } else if (c < ldw) {
                for (int d = 0; d < D; ++d) acc += (double)W[(size_t)ro * D + d] * (double)Wenc[(size_t)d * ldw + c];
            }
            v = (float)acc;
        }
        if (is_bias) bce[r] = v; else Wce[i] = v;
    }
}
// Wce [5D, Kx], bce [5D]; A = W - W^T - gamma I (from ctan_wcat_prep / ctan_antisym_prep).
extern "C" int ctan_fuse_enc_prep(const float* Wq, const float* Wk, const float* Wv, const float* A, const float* bq,
                                  const float* bk, const float* bv, const float* b, const float* Wenc, const float* benc,
                                  int D, int Fn, int Kx, float* Wce, float* bce, cudaStream_t stream) {
    if (D <= 0 || Fn < 0 || Kx < D + Fn || Kx % 32 != 0) return CTAN_ERR_BAD_ARG;
    fuse_enc_kernel<<<ctan_grid(5LL * D * Kx, 256), 256, 0, stream>>>(Wq, Wk, Wv, A, bq, bk, bv, b, Wenc, benc, D, Fn, Kx, Wce, bce);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// The same panel formed EVERY training step (the weights move with each optimizer step): one kernel, tiled through
// shared memory, fp32 accumulation, A = W - W^T - gamma I taken on the fly, hi/lo split (3xTF32 operands) written directly.
//   rows [0,D)   : Wenc zero-padded to Kx                       bias benc
//   rows [D,5D)  : [Wq;Wk;Wv;A] Wenc                            bias [Wq;Wk;Wv;A] benc + [bq;bk;bv;b]
// The bias is column Kx of the extended product.  grid = (Kx/32 + 1, 5D/32), 256 threads, 4 outputs per thread;
// D % 32 == 0, D <= 256 (each thread stages D/8 elements of either tile from registers).
__global__ void __launch_bounds__(256)
fuse_enc_split_kernel(const float* __restrict__ Wq, const float* __restrict__ Wk, const float* __restrict__ Wv,
                      const float* __restrict__ W, float gamma, const float* __restrict__ bq,
                      const float* __restrict__ bk, const float* __restrict__ bv, const float* __restrict__ b,
                      const float* __restrict__ Wenc, const float* __restrict__ benc, int D, int Fn, int Kx,
                      float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ bce) {
    ctan_pdl_begin();
    extern __shared__ float fsm[];
    const int ldw = D + Fn;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;                       // output column; c == Kx is the bias
    const int r0 = blockIdx.y * 32;
    auto put = [&](int r, float v) {
        if (c < Kx) {
            const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
            hi[(size_t)r * Kx + c] = h;
            lo[(size_t)r * Kx + c] = v - h;
        } else if (c == Kx) {
            bce[r] = v;
        }
    };
    if (r0 < D) {                                             // the enc_x block itself
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = r0 + ty + 8 * j;
            put(r, c < ldw ? Wenc[(size_t)r * ldw + c] : (c == Kx ? benc[r] : 0.f));
        }
        return;
    }
    const int seg = (r0 - D) / D, ro0 = (r0 - D) - seg * D;   // (D % 32 == 0: a tile never straddles two segments)
    const float* Wl = seg == 0 ? Wq : (seg == 1 ? Wk : (seg == 2 ? Wv : W));
    const float* bb = seg == 0 ? bq : (seg == 1 ? bk : (seg == 2 ? bv : b));
    // the whole K extent of both operand tiles at once (every global load of the CTA in flight together: the kernel
    // sits on the step's critical path, and it is latency, not bandwidth, that it costs)
    const int SA = D + 4;                                     // 16-byte aligned rows: the k loop reads them as float4
    float* As = fsm;                                          // [32][D + 4]   rows of [Wq;Wk;Wv;A]   (warp-uniform reads)
    float* Bs = fsm + 32 * SA;                                // [D][33]       columns of [Wenc | benc]
    const int nq = D >> 5;                                    // <= 8
    float va[4][8], vb[4][8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < nq) {
                va[j][q] = Wl[(size_t)(ro0 + ty + 8 * j) * D + tx + 32 * q];       // row ty+8j, k = tx+32q
                const int dk = ty + 8 * (j * nq + q);                               // k = dk, column tx
                vb[j][q] = c < ldw ? Wenc[(size_t)dk * ldw + c] : (c == Kx ? benc[dk] : 0.f);
            }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (q < nq) {
                As[(ty + 8 * j) * SA + tx + 32 * q] = va[j][q];
                Bs[(ty + 8 * (j * nq + q)) * 33 + tx] = vb[j][q];
            }
    if (seg == 3) {                                           // A = W - W^T - gamma I: the transposed half, read coalesced
        __syncthreads();                                      // (row = lane, k = ty + 8u) and folded in through shared memory
        for (int u0 = 0; u0 < (D >> 3); u0 += 8) {
            float vt[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) vt[u] = (u0 + u) < (D >> 3) ? W[(size_t)(ty + 8 * (u0 + u)) * D + ro0 + tx] : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int d = ty + 8 * (u0 + u);
                if (d < D) As[tx * SA + d] = As[tx * SA + d] - vt[u] - ((ro0 + tx) == d ? gamma : 0.f);
            }
        }
    }
    __syncthreads();
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < D; k += 4) {
        const float b0 = Bs[(k + 0) * 33 + tx], b1 = Bs[(k + 1) * 33 + tx], b2 = Bs[(k + 2) * 33 + tx], b3 = Bs[(k + 3) * 33 + tx];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 a = *reinterpret_cast<const float4*>(As + (ty + 8 * j) * SA + k);
            acc[j] = fmaf(a.x, b0, acc[j]);
            acc[j] = fmaf(a.y, b1, acc[j]);
            acc[j] = fmaf(a.z, b2, acc[j]);
            acc[j] = fmaf(a.w, b3, acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ro = ro0 + ty + 8 * j;
        put(r0 + ty + 8 * j, c == Kx ? acc[j] + bb[ro] : acc[j]);
    }
}
// Wce_hi / Wce_lo [5D, Kx], bce [5D] from the raw weights; D % 32 == 0, D <= 256, Kx % 32 == 0, Kx >= D + Fn.
extern "C" int ctan_fuse_enc_prep_split(const float* Wq, const float* Wk, const float* Wv, const float* W, float gamma,
                                        const float* bq, const float* bk, const float* bv, const float* b,
                                        const float* Wenc, const float* benc, int D, int Fn, int Kx, float* Wce_hi,
                                        float* Wce_lo, float* bce, cudaStream_t stream) {
    if (D <= 0 || D % 32 != 0 || D > 256 || Fn < 0 || Kx < D + Fn || Kx % 32 != 0) return CTAN_ERR_BAD_ARG;
    const size_t smem = (size_t)(32 * (D + 4) + D * 33) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(fuse_enc_split_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    CTAN_LAUNCH((fuse_enc_split_kernel), dim3(Kx / 32 + 1, 5 * D / 32), 256, smem, stream, Wq, Wk, Wv, W, gamma, bq, bk, bv, b,
                Wenc, benc, D, Fn, Kx, Wce_hi, Wce_lo, bce);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// zext[r] = [ mem[n_id[r]] (D) | xfeat[n_id[r]] (Fn) | 0 ... ] with row pitch Kx (a dense TMA operand); one warp per row.
__global__ void enc_prep_ext_kernel(const float* __restrict__ mem, int D, const float* __restrict__ xfeat, int Fn,
                                    const int64_t* __restrict__ n_id, int N_host, const int* __restrict__ N_dev, int Kx,
                                    float* __restrict__ zext) {
    ctan_pdl_begin();
    const int N = N_dev ? *N_dev : N_host;
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const bool vec = (D & 3) == 0 && (Kx & 3) == 0 && ((((uintptr_t)mem | (uintptr_t)zext) & 15) == 0);
    for (int r = warp; r < N; r += nw) {
        const int64_t g = n_id[r];
        float* out = zext + (size_t)r * Kx;
        if (vec) {
            for (int d4 = lane; d4 < (D >> 2); d4 += 32)
                reinterpret_cast<float4*>(out)[d4] = __ldg(reinterpret_cast<const float4*>(mem + (size_t)g * D) + d4);
        } else {
            for (int d = lane; d < D; d += 32) out[d] = __ldg(mem + (size_t)g * D + d);
        }
        for (int c = D + lane; c < Kx; c += 32) out[c] = c < D + Fn ? __ldg(xfeat + (size_t)g * Fn + (c - D)) : 0.f;
    }
}
extern "C" int ctan_enc_prep_ext(const float* mem, int D, const float* xfeat, int Fn, const int64_t* n_id, int N,
                                 const int* N_dev, int Kx, float* zext, cudaStream_t stream) {
    if (N <= 0) return 0;
    if (Kx < D + Fn) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((enc_prep_ext_kernel), ctan_grid((long long)N * 32, 256, 4), 256, 0, stream, mem, D, xfeat, Fn, n_id, N, N_dev,
                Kx, zext);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// [X0 | QKVA] = zext Wce^T + bce on tcgen05: columns [0,D) -> X0 [N,D], columns [D,5D) -> QKVA [N,4D].
// Wce_hi/Wce_lo: the split panel of ctan_fuse_enc_prep's Wce (ctan_pack_split), or a BF16 panel in mode 2.
extern "C" int ctan_enc_qkva_tc(const float* zext, int N, const int* N_dev, int Kx, const float* Wce_hi, const float* Wce_lo,
                                int D, const float* bce, float* X0, float* QKVA, int mode, cudaStream_t stream) {
    if (N <= 0) return 0;
    const int m_typ = mode >> 8;
    const bool pair = tc::want_pair(mode & 0xFF);
    mode &= 0x3F;
    const bool bf = mode == 2;
    if (Kx % (bf ? 2 * tc::BKF : tc::BKF) != 0 || D % tc::BN != 0 ||
        (((uintptr_t)zext | (uintptr_t)Wce_hi | (bf ? 0 : (uintptr_t)Wce_lo) | (uintptr_t)X0 | (uintptr_t)QKVA | (uintptr_t)bce) & 15))
        return CTAN_ERR_UNSUPPORTED;
    CUtensorMap tmA, tmBhi, tmBlo;
    int rc = tc::make_map(&tmA, zext, N, Kx);
    if (!rc) rc = bf ? tc::make_map_bf16(&tmBhi, Wce_hi, 5 * D, Kx) : tc::make_map(&tmBhi, Wce_hi, 5 * D, Kx);
    if (!rc) rc = bf ? tc::make_map_bf16(&tmBlo, Wce_hi, 5 * D, Kx) : tc::make_map(&tmBlo, Wce_lo, 5 * D, Kx);
    if (rc) return rc;
    cudaError_t e = cudaFuncSetAttribute(tc::qkva_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);
    if (e != cudaSuccess) return (int)e;
    // one bias segment spanning all 5D columns; columns >= D are redirected to QKVA (row pitch 4D)
    tc::Bias4 epi{{bce, nullptr, nullptr, nullptr}, 5 * D, 1, nullptr, 0, 1, QKVA, D, 4 * D};
    dim3 grid((5 * D) / tc::BN, tc::bounded_rows(N, N_dev, (5 * D) / tc::BN, m_typ));
    { const int lrc = tc::launch(grid, pair, stream, tmA, tmBhi, tmBlo, epi, X0, D, N, N_dev, 5 * D, Kx, mode); if (lrc) return lrc; }
    CTAN_CHECK_LAUNCH();
    return 0;
}
