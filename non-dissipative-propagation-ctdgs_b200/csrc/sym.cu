// Peer-memory exchange of the sharded TGB evaluation: the ONE exchange step per 200-query batch,
// done by our own kernels over NVLink instead of a host-issued NCCL all-gather between two CUDA graphs.
//
// Replaces: the reference has no multi-GPU path; this is the collective of the restructured tgb_test loop
// (train_link.py:111-199: per-query scores + the (src, pos_dst) embeddings every replica needs for the identical
// memory write-back at :198).
//
// Every rank owns one SYMMETRIC allocation (cudaMalloc, exported with CUDA IPC and mapped by every peer):
//     [ flags: 2 parities x world x uint64 | pad to 256 B | rows: 2 parities x B x row fp32 ]
// A rank that has scored its slice of the batch stores its packed query rows straight into EVERY peer's rows buffer
// (plain st.global over NVLink / NVSwitch, 16-byte vectors), fences system-wide, and the last CTA to finish
// releases flag[parity][my_rank] = epoch on every peer.  The consumer side is a one-CTA kernel that spins
// (ld.acquire.sys) until all `world` flags of the parity carry the epoch.  Both kernels sit INSIDE the captured
// per-batch graph: one graph replay per batch, no host-side collective call, no NCCL launch latency.
// Buffers are double-buffered by batch parity: a rank can be at most one batch ahead of the slowest one (it cannot
// pass its own wait), so parity p is rewritten only after every rank has consumed the previous use of p.
#include "common.cuh"


// flag store after ONE system-scope fence (a st.release.sys per peer would emit a MEMBAR.SYS each)
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------ host-side setup (no stream)
extern "C" int ctan_sym_alloc(size_t bytes, void** out) {
    if (!out || bytes == 0) return CTAN_ERR_BAD_ARG;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(*out, 0, bytes);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaDeviceSynchronize();
}
extern "C" int ctan_sym_free(void* p) { return p ? (int)cudaFree(p) : 0; }
// handle: 64 bytes (cudaIpcMemHandle_t)
extern "C" int ctan_ipc_export(void* p, unsigned char* handle) {
    if (!p || !handle) return CTAN_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) return (int)e;
    memcpy(handle, &h, sizeof(h));
    return 0;
}
extern "C" int ctan_ipc_import(const unsigned char* handle, void** out) {
    if (!handle || !out) return CTAN_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    return (int)cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess);
}
extern "C" int ctan_ipc_close(void* p) { return p ? (int)cudaIpcCloseMemHandle(p) : 0; }

// ------------------------------------------------------------------------------------------ exchange kernels
// peers: device array of `world` base pointers (the local mapping of every rank's symmetric allocation, own rank
// included).  Layout constants are passed explicitly so that the kernels stay plain C ABI.
//   flag(r_owner, parity, r_writer) = (uint64*)peers[r_owner] + parity * world + r_writer
//   rows(r_owner, parity)           = (float*)((char*)peers[r_owner] + rows_off) + parity * B * row
__global__ void xchg_push_kernel(const float* __restrict__ send, int n_rows, int row, int q0, int B,
                                 void* const* __restrict__ peers, int world, int my_rank, int parity, size_t rows_off,
                                 const unsigned long long* __restrict__ epoch_ptr, unsigned int* __restrict__ ticket) {
    ctan_pdl_begin();
    const size_t n = (size_t)n_rows * row;                  // contiguous block of this rank's rows
    for (int p = blockIdx.y; p < world; p += gridDim.y) {
        float* dst = reinterpret_cast<float*>(reinterpret_cast<char*>(peers[p]) + rows_off)
                     + ((size_t)parity * B + q0) * row;
        const bool v4 = ((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(send)) & 15) == 0;
        if (v4) {
            const size_t n4 = n >> 2;
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
                reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(send)[i];
            for (size_t i = (n4 << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
                dst[i] = send[i];
        } else {
            for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
                dst[i] = send[i];
        }
    }
    // every CTA: make its stores visible system-wide, then take a ticket; the last one signals all peers
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int total = gridDim.x * gridDim.y;
        if (atomicAdd(ticket, 1u) == total - 1) {
            *ticket = 0;
            __threadfence_system();
            const unsigned long long e = *epoch_ptr + 1;
            for (int p = 0; p < world; ++p)
                st_relaxed_sys(reinterpret_cast<unsigned long long*>(peers[p]) + (size_t)parity * world + my_rank, e);
        }
    }
}

// send [n_rows,row]: this rank's packed rows (queries q0 .. q0+n_rows of the batch).  epoch (device uint64): number
// of exchanges completed so far on this rank (advanced by ctan_xchg_wait); ticket: device uint32, zero.
extern "C" int ctan_xchg_push(const float* send, int n_rows, int row, int q0, int B, void* const* peers, int world,
                              int my_rank, int parity, size_t rows_off, const unsigned long long* epoch,
                              unsigned int* ticket, cudaStream_t stream) {
    if (world < 1 || my_rank < 0 || my_rank >= world || (parity & ~1) || !peers || !epoch || !ticket)
        return CTAN_ERR_BAD_ARG;
    if (n_rows < 0 || q0 < 0 || q0 + n_rows > B) return CTAN_ERR_BAD_ARG;
    const size_t n = (size_t)n_rows * row;
    int gx = (int)((n / 4 + 255) / 256);
    if (gx < 1) gx = 1;
    if (gx > 16) gx = 16;
    CTAN_LAUNCH((xchg_push_kernel), dim3(gx, world), 256, 0, stream, send, n_rows, row, q0, B, peers, world, my_rank,
                parity, rows_off, epoch, ticket);
    CTAN_CHECK_LAUNCH();
    return 0;
}

__global__ void xchg_wait_kernel(const unsigned long long* __restrict__ flags, int world,
                                 unsigned long long* __restrict__ epoch_ptr, unsigned long long* __restrict__ spins) {
    ctan_pdl_begin();
    const unsigned long long want = *epoch_ptr + 1;
    unsigned long long n = 0;
    if ((int)threadIdx.x < world)
        while (ld_acquire_sys(flags + threadIdx.x) < want) { ++n; __nanosleep(20); }
    __syncthreads();
    if (threadIdx.x == 0) {
        *epoch_ptr = want;
        if (spins) *spins += n;
    }
}

// flags: THIS rank's flag row of the parity ([world] uint64 inside its own symmetric allocation).
extern "C" int ctan_xchg_wait(const unsigned long long* flags, int world, unsigned long long* epoch,
                              unsigned long long* spins, cudaStream_t stream) {
    if (world < 1 || world > 1024 || !flags || !epoch) return CTAN_ERR_BAD_ARG;
    const int threads = ((world + 31) / 32) * 32;
    CTAN_LAUNCH((xchg_wait_kernel), 1, threads, 0, stream, flags, world, epoch, spins);
    CTAN_CHECK_LAUNCH();
    return 0;
}
