// Last-neighbour ring buffer (insert / K-hop lookup / relabel) and memory write-back.
//
// Replaces, bit-exactly (CPU / stable-sort semantics):
//   LastNeighborLoader.reset_state / insert / __call__   TGB/modules/neighbor_loader.py:16-85
//   seed-set unique + helper[n_id]=arange                train_link.py:252,258-263
//   SimpleMemory.update_state + LastAggregator           models/memory_layers.py:132-144,
//                                                        TGB/modules/msg_agg.py:15-21
//
// HBM layout: nbr[num_nodes,S] i64, eid[num_nodes,S] i64 (row = descending e_id, -1 padding last,
// exactly the reference's tensors) plus deg[num_nodes] i32 = number of valid slots per row
// (auxiliary, maintained by insert; lets lookups skip reading rows just to count them).
// All kernels are warp-cooperative integer work, bound by HBM/L2 latency; no sort, no unique,
// no top-k, no host synchronisation.
#include "common.cuh"

// ------------------------------------------------------------------------------------------
// K0: reset
// ------------------------------------------------------------------------------------------
__global__ void nbr_reset_kernel(int64_t* __restrict__ nbr, int64_t* __restrict__ eid,
                                 int32_t* __restrict__ deg, long long rows, int S) {
    ctan_pdl_begin();
    const long long n = rows * S;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        eid[i] = -1;
        nbr[i] = 0;
    }
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += stride) deg[i] = 0;
}

extern "C" int ctan_nbr_reset(int64_t* nbr, int64_t* eid, int32_t* deg, int64_t num_nodes, int S,
                              cudaStream_t stream) {
    if (num_nodes <= 0 || S <= 0) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((nbr_reset_kernel), ctan_grid(num_nodes * S, 256 * 4), 256, 0, stream, nbr, eid, deg, num_nodes, S);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// K2: insert.  Entry p in [0,2B): p<B -> (node=dst[p], other=src[p]); p>=B -> (node=src[p-B],
// other=dst[p-B]); event id = cur + (p mod B).  This enumeration IS the stable-sorted order
// within each node (neighbor_loader.py:47-56).  For a node with c entries the last min(c,S)
// survive (dense slot = position % S, last write wins, :61-70); survivors are then the largest
// e_ids of the merged row (top-k, :79), so they occupy slots [0,m) in descending e_id and the
// old entries shift right by m.
//   kernel A (warp per entry): rank/count of the entry within its node; the rank-0 warp shifts
//                              the old row and updates deg.
//   kernel B (warp per entry): surviving entries compute their slot and write.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t ins_node(const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                            int B, int p) {
    return p < B ? dst[p] : src[p - B];
}

__global__ void nbr_insert_rank_shift_kernel(int64_t* __restrict__ nbr, int64_t* __restrict__ eid,
                                             int32_t* __restrict__ deg, const int64_t* __restrict__ src,
                                             const int64_t* __restrict__ dst, int B, int S,
                                             int2* __restrict__ rc) {
    ctan_pdl_begin();
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int n = 2 * B;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < n; p += gridDim.x * wpb) {
        const int64_t u = ins_node(src, dst, B, p);
        int rank = 0, cnt = 0;
        for (int q0 = 0; q0 < n; q0 += 32) {
            const int q = q0 + lane;
            const bool same = (q < n) && (ins_node(src, dst, B, q) == u);
            const unsigned m = __ballot_sync(0xffffffffu, same);
            cnt += __popc(m);
            if (q0 + 32 <= p) rank += __popc(m);
            else if (q0 <= p) rank += __popc(m & ((1u << (p - q0)) - 1u));
        }
        if (lane == 0) rc[p] = make_int2(rank, cnt);
        if (rank == 0) {  // warp-uniform: this warp owns node u's row for the shift
            const int m = min(cnt, S);
            const int dd = deg[u];
            const int keep = min(dd, S - m);
            int64_t* rn = nbr + u * S;
            int64_t* re = eid + u * S;
            if (keep > 0) {
                for (int c0 = ((keep - 1) >> 5) << 5; c0 >= 0; c0 -= 32) {  // descending chunks: in-place safe
                    const int s = c0 + lane;
                    const bool act = s < keep;
                    int64_t vn = 0, ve = 0;
                    if (act) { vn = rn[s]; ve = re[s]; }
                    __syncwarp();
                    if (act) { rn[s + m] = vn; re[s + m] = ve; }
                    __syncwarp();
                }
            }
            __syncwarp();                    // every lane has read deg[u] before lane 0 rewrites it
            if (lane == 0) deg[u] = min(S, dd + m);
        }
    }
}

__global__ void nbr_insert_write_kernel(int64_t* __restrict__ nbr, int64_t* __restrict__ eid,
                                        const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                        int B, int S, int64_t cur_host, const int64_t* __restrict__ cur_dev,
                                        const int2* __restrict__ rc) {
    ctan_pdl_begin();
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int n = 2 * B;
    const int64_t cur = cur_dev ? *cur_dev : cur_host;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < n; p += gridDim.x * wpb) {
        const int2 me = rc[p];
        const int thr = me.y - S;        // survive iff rank >= cnt - S
        if (me.x < thr) continue;        // warp-uniform
        const int64_t u = ins_node(src, dst, B, p);
        const int ip = p < B ? p : p - B;
        int pos = 0;
        for (int q0 = 0; q0 < n; q0 += 32) {
            const int q = q0 + lane;
            bool pre = false;
            if (q < n && q != p && ins_node(src, dst, B, q) == u && rc[q].x >= thr) {
                const int iq = q < B ? q : q - B;
                pre = (iq > ip) || (iq == ip && q > p);   // larger e_id first; self-loop twins: any fixed order
            }
            pos += __popc(__ballot_sync(0xffffffffu, pre));
        }
        if (lane == 0) {
            nbr[u * S + pos] = p < B ? src[p] : dst[p - B];
            eid[u * S + pos] = cur + ip;
        }
    }
}

// ws: >= 2*B*sizeof(int2) bytes.
extern "C" int ctan_nbr_insert(int64_t* nbr, int64_t* eid, int32_t* deg, const int64_t* src, const int64_t* dst,
                               int B, int64_t cur_e_id, const int64_t* cur_e_id_dev, int S, void* ws,
                               cudaStream_t stream) {
    if (B < 0 || S <= 0) return CTAN_ERR_BAD_ARG;
    if (B == 0) return 0;
    const int wpb = 8;
    const int grid = ctan_grid(2LL * B, wpb);
    CTAN_LAUNCH((nbr_insert_rank_shift_kernel), grid, wpb * 32, 0, stream, nbr, eid, deg, src, dst, B, S, (int2*)ws);
    CTAN_CHECK_LAUNCH();
    CTAN_LAUNCH((nbr_insert_write_kernel), grid, wpb * 32, 0, stream, nbr, eid, src, dst, B, S, cur_e_id, cur_e_id_dev,
                                                          (const int2*)ws);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// K1: K-hop lookup + relabel.
//   bitmap  : node is in the current set          (n_id of the reference after the hops so far)
//   cbitmap : node has been expanded (= "centre") (the n_id FED to the reference's last call)
// expand (one launch per hop): an item is claimed through cbitmap (atomicOr elects one owner),
// its valid slots are read and the neighbours are marked in bitmap; first-time-marked nodes are
// appended to the next hop's frontier.  The resulting SETS do not depend on the atomics' order.
// scan (single CTA): popcount prefix over bitmap -> sorted-unique n_id + assoc (global->local) and,
// from deg of the centre bits, the CSR rowptr of the destination-sorted edge list.
// fill: one lane group per local node writes its edges (neighbour local id, centre local id, e_id).
// ------------------------------------------------------------------------------------------
__global__ void lk_expand_kernel(const int64_t* __restrict__ list, int n_host, const int* __restrict__ n_dev,
                                 int list_cap, const int64_t* __restrict__ nbr, const int32_t* __restrict__ deg, int S, int lpi,
                                 uint32_t* __restrict__ bitmap, uint32_t* __restrict__ cbitmap,
                                 int64_t* __restrict__ next, int* __restrict__ next_cnt, int next_cap,
                                 int* __restrict__ err) {
    ctan_pdl_begin();
    const int end = min(n_dev ? *n_dev : n_host, list_cap);
    const int gpw = 32 / lpi;
    const int lane = threadIdx.x & 31, g = lane / lpi, gl = lane % lpi;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int base = warp * gpw; base < end; base += nwarps * gpw) {   // warp-uniform trip count
        const int it = base + g;
        const bool valid = it < end;
        const int64_t u = valid ? list[it] : 0;
        int own = 0;
        if (valid && gl == 0) {
            const unsigned bit = 1u << (u & 31);
            const unsigned old = atomicOr(&cbitmap[u >> 5], bit);
            own = !(old & bit);
            if (own) atomicOr(&bitmap[u >> 5], bit);
        }
        own = __shfl_sync(0xffffffffu, own, g * lpi);
        if (own) {
            const int d = deg[u];
            for (int s = gl; s < d; s += lpi) {
                const int64_t v = nbr[u * S + s];
                const unsigned bit = 1u << (v & 31);
                const unsigned old = atomicOr(&bitmap[v >> 5], bit);
                if (!(old & bit) && next_cnt) {
                    const int k = atomicAdd(next_cnt, 1);
                    if (k < next_cap) next[k] = v; else *err = 1;
                }
            }
        }
    }
}

__global__ void __launch_bounds__(1024)
lk_scan_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ cbitmap,
               const int32_t* __restrict__ deg, int n_words, int64_t* __restrict__ n_id_out,
               int64_t* __restrict__ assoc, int32_t* __restrict__ rowptr, int32_t* __restrict__ counts,
               int n_cap, int e_cap, int* __restrict__ frontier_cnt, int n_frontier_cnt) {
    ctan_pdl_begin();
    __shared__ int s_n[32], s_e[32], s_c[32];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int chunk = (n_words + 1023) / 1024;
    const int w0 = min(n_words, tid * chunk), w1 = min(n_words, w0 + chunk);
    int nn = 0, ne = 0, nc = 0;
    for (int w = w0; w < w1; ++w) {
        const uint32_t b = bitmap[w];
        if (!b) continue;
        nn += __popc(b);
        uint32_t c = cbitmap[w];
        nc += __popc(c);
        while (c) {
            const int bit = __ffs(c) - 1;
            c &= c - 1;
            ne += deg[(long long)w * 32 + bit];
        }
    }
    // block-wide exclusive scan of (nn, ne); total of nc
    int in = nn, ie = ne, ic = nc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int a = __shfl_up_sync(0xffffffffu, in, o);
        const int b = __shfl_up_sync(0xffffffffu, ie, o);
        const int c = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) { in += a; ie += b; ic += c; }
    }
    if (lane == 31) { s_n[wid] = in; s_e[wid] = ie; s_c[wid] = ic; }
    __syncthreads();
    if (wid == 0) {
        int a = s_n[lane], b = s_e[lane], c = s_c[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, a, o);
            const int y = __shfl_up_sync(0xffffffffu, b, o);
            const int z = __shfl_up_sync(0xffffffffu, c, o);
            if (lane >= o) { a += x; b += y; c += z; }
        }
        s_n[lane] = a; s_e[lane] = b; s_c[lane] = c;
    }
    __syncthreads();
    int base_n = in - nn + (wid ? s_n[wid - 1] : 0);
    int base_e = ie - ne + (wid ? s_e[wid - 1] : 0);
    const int tot_n = s_n[31], tot_e = s_e[31], tot_c = s_c[31];
    const bool fits = (tot_n <= n_cap) && (tot_e <= e_cap);
    if (fits) {
        for (int w = w0; w < w1; ++w) {
            uint32_t b = bitmap[w];
            if (!b) continue;
            const uint32_t c = cbitmap[w];
            while (b) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                const long long node = (long long)w * 32 + bit;
                n_id_out[base_n] = node;
                assoc[node] = base_n;
                rowptr[base_n] = base_e;
                if ((c >> bit) & 1u) base_e += deg[node];
                ++base_n;
            }
        }
    }
    if (tid == 0) {
        if (fits) rowptr[tot_n] = tot_e;
        counts[0] = fits ? tot_n : 0;
        counts[1] = fits ? tot_e : 0;
        counts[2] = tot_c;
        if (!fits) counts[3] = 2;       // capacity overflow (1 = frontier overflow, set by expand)
        counts[4] = tot_n;              // true sizes, for diagnostics
        counts[5] = tot_e;
        for (int i = 0; i < n_frontier_cnt; ++i) frontier_cnt[i] = 0;   // restore the entry invariant
    }
}

// When the scan reported an overflow the bitmaps still have to be wiped: done by a plain memset
// from the host wrapper of the strict API; the zero-sync engine sizes its capacities so that it
// cannot happen and checks counts[3] at the end of the run.
__global__ void lk_fill_kernel(const int64_t* __restrict__ nbr, const int64_t* __restrict__ eid, int S, int lpi,
                               const int64_t* __restrict__ n_id, const int32_t* __restrict__ rowptr,
                               const int64_t* __restrict__ assoc, int n_host, const int32_t* __restrict__ n_dev,
                               int64_t* __restrict__ edge_src, int64_t* __restrict__ edge_dst,
                               int64_t* __restrict__ e_id_out, uint32_t* __restrict__ bitmap,
                               uint32_t* __restrict__ cbitmap, const int64_t* __restrict__ last_update,
                               const int64_t* __restrict__ t_tab, float mean, float stdv,
                               float* __restrict__ rel) {
    ctan_pdl_begin();
    const int N = n_dev ? *n_dev : n_host;
    const int gpw = 32 / lpi;
    const int lane = threadIdx.x & 31, g = lane / lpi, gl = lane % lpi;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp * gpw + g; r < N; r += nwarps * gpw) {
        const int64_t node = n_id[r];
        const int off = rowptr[r], d = rowptr[r + 1] - off;
        for (int s = gl; s < d; s += lpi) {
            const int64_t nb = nbr[node * S + s], ev = eid[node * S + s];
            edge_src[off + s] = assoc[nb];
            edge_dst[off + s] = r;
            e_id_out[off + s] = ev;
            if (rel) {      // fused delta-t normalisation (gnn_layers.py:94-95): |last_update[src] - t_e|
                long long dt = last_update[nb] - t_tab[ev];
                if (dt < 0) dt = -dt;
                rel[off + s] = __fdiv_rn(__ll2float_rn(dt) - mean, stdv);
            }
        }
        if (gl == 0) { bitmap[node >> 5] = 0u; cbitmap[node >> 5] = 0u; }
    }
}

// Two-level variant of lk_scan_kernel for large id spaces: SCAN_CTAS CTAs each own a contiguous range of bitmap
// words.  Pass 1 writes per-CTA (nodes, edges, centres) totals; pass 2 turns them into each CTA's base offsets
// (a 128-entry serial prefix, redone by every CTA) and emits n_id / assoc / rowptr exactly like the 1-CTA kernel.
constexpr int SCAN_CTAS = 128;

__device__ __forceinline__ void scan_count_words(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ cbitmap,
                                                 const int32_t* __restrict__ deg, int w0, int w1, int& nn, int& ne,
                                                 int& nc) {
    nn = ne = nc = 0;
    for (int w = w0; w < w1; ++w) {
        const uint32_t b = bitmap[w];
        if (!b) continue;
        nn += __popc(b);
        uint32_t c = cbitmap[w];
        nc += __popc(c);
        while (c) {
            const int bit = __ffs(c) - 1;
            c &= c - 1;
            ne += deg[(long long)w * 32 + bit];
        }
    }
}

// exclusive block scan of (a,b) over 256 threads; returns this thread's exclusive prefixes and the block totals
__device__ __forceinline__ void block_scan2(int a, int b, int& ea, int& eb, int& ta, int& tb, int* s_a, int* s_b) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int ia = a, ib = b;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int x = __shfl_up_sync(0xffffffffu, ia, o), y = __shfl_up_sync(0xffffffffu, ib, o);
        if (lane >= o) { ia += x; ib += y; }
    }
    if (lane == 31) { s_a[wid] = ia; s_b[wid] = ib; }
    __syncthreads();
    int pa = 0, pb = 0, sa = 0, sb = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) {
        if (i < wid) { pa += s_a[i]; pb += s_b[i]; }
        sa += s_a[i]; sb += s_b[i];
    }
    ea = ia - a + pa; eb = ib - b + pb; ta = sa; tb = sb;
    __syncthreads();
}

__global__ void __launch_bounds__(256)
lk_scan_partial_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ cbitmap,
                       const int32_t* __restrict__ deg, int n_words, int* __restrict__ partials) {
    ctan_pdl_begin();
    __shared__ int s_a[8], s_b[8], s_c[8];
    const int chunk = (n_words + SCAN_CTAS - 1) / SCAN_CTAS;
    const int c0 = min(n_words, (int)blockIdx.x * chunk), c1 = min(n_words, c0 + chunk);
    const int sub = (chunk + 255) / 256;
    const int w0 = min(c1, c0 + (int)threadIdx.x * sub), w1 = min(c1, w0 + sub);
    int nn, ne, nc;
    scan_count_words(bitmap, cbitmap, deg, w0, w1, nn, ne, nc);
    nn = warp_sum_i(nn); ne = warp_sum_i(ne); nc = warp_sum_i(nc);
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = nn; s_b[threadIdx.x >> 5] = ne; s_c[threadIdx.x >> 5] = nc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        int a = 0, b = 0, c = 0;
        for (int i = 0; i < 8; ++i) { a += s_a[i]; b += s_b[i]; c += s_c[i]; }
        partials[blockIdx.x * 3 + 0] = a; partials[blockIdx.x * 3 + 1] = b; partials[blockIdx.x * 3 + 2] = c;
    }
}

__global__ void __launch_bounds__(256)
lk_scan_emit_kernel(const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ cbitmap,
                    const int32_t* __restrict__ deg, int n_words, const int* __restrict__ partials,
                    int64_t* __restrict__ n_id_out, int64_t* __restrict__ assoc, int32_t* __restrict__ rowptr,
                    int32_t* __restrict__ counts, int n_cap, int e_cap, int* __restrict__ frontier_cnt,
                    int n_frontier_cnt) {
    ctan_pdl_begin();
    __shared__ int s_a[8], s_b[8], s_base[5];
    if (threadIdx.x == 0) {
        int bn = 0, be = 0, tn = 0, te = 0, tcn = 0;
        for (int i = 0; i < SCAN_CTAS; ++i) {
            if (i < (int)blockIdx.x) { bn += partials[i * 3]; be += partials[i * 3 + 1]; }
            tn += partials[i * 3]; te += partials[i * 3 + 1]; tcn += partials[i * 3 + 2];
        }
        s_base[0] = bn; s_base[1] = be; s_base[2] = tn; s_base[3] = te; s_base[4] = tcn;
    }
    __syncthreads();
    const int tot_n = s_base[2], tot_e = s_base[3];
    const bool fits = (tot_n <= n_cap) && (tot_e <= e_cap);
    const int chunk = (n_words + SCAN_CTAS - 1) / SCAN_CTAS;
    const int c0 = min(n_words, (int)blockIdx.x * chunk), c1 = min(n_words, c0 + chunk);
    const int sub = (chunk + 255) / 256;
    const int w0 = min(c1, c0 + (int)threadIdx.x * sub), w1 = min(c1, w0 + sub);
    int nn, ne, nc;
    scan_count_words(bitmap, cbitmap, deg, w0, w1, nn, ne, nc);
    int en, ee, tn2, te2;
    block_scan2(nn, ne, en, ee, tn2, te2, s_a, s_b);
    int base_n = s_base[0] + en, base_e = s_base[1] + ee;
    if (fits) {
        for (int w = w0; w < w1; ++w) {
            uint32_t b = bitmap[w];
            if (!b) continue;
            const uint32_t c = cbitmap[w];
            while (b) {
                const int bit = __ffs(b) - 1;
                b &= b - 1;
                const long long node = (long long)w * 32 + bit;
                n_id_out[base_n] = node;
                assoc[node] = base_n;
                rowptr[base_n] = base_e;
                if ((c >> bit) & 1u) base_e += deg[node];
                ++base_n;
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (fits) rowptr[tot_n] = tot_e;
        counts[0] = fits ? tot_n : 0;
        counts[1] = fits ? tot_e : 0;
        counts[2] = s_base[4];
        if (!fits) counts[3] = 2;
        counts[4] = tot_n;
        counts[5] = tot_e;
        for (int i = 0; i < n_frontier_cnt; ++i) frontier_cnt[i] = 0;
    }
}

static inline int lanes_per_item(int S) {
    int l = 1;
    while (l < S && l < 32) l <<= 1;
    return l;
}

// Phase 1 of a lookup: mark + expand `num_hops` times, then scan.  On return (stream-ordered):
//   counts[0]=N, counts[1]=E, counts[2]=#centres, counts[3]=error flag (0 ok), counts[4..5]=true N,E
//   n_id_out[0..N) sorted unique, assoc[n_id_out[r]] = r, rowptr[0..N] (edges sorted by centre).
// bitmap/cbitmap [ceil(num_nodes/32)] and frontier_cnt[num_hops] must be zero on entry; frontier_cnt
// is zero again on exit, the bitmaps after ctan_nbr_fill.  frontier: [max(num_hops-1,1) * frontier_cap].
extern "C" int ctan_nbr_lookup(const int64_t* nbr, const int32_t* deg, int64_t num_nodes, int S,
                               const int64_t* seeds, int n_seeds, const int* n_seeds_dev, int num_hops,
                               uint32_t* bitmap, uint32_t* cbitmap, int64_t* frontier, int* frontier_cnt,
                               int frontier_cap, int64_t* n_id_out, int n_cap, int64_t* assoc, int32_t* rowptr,
                               int e_cap, int32_t* counts, cudaStream_t stream) {
    if (num_hops < 1 || S <= 0 || n_seeds < 0 || num_nodes <= 0) return CTAN_ERR_BAD_ARG;
    const int lpi = lanes_per_item(S);
    const int threads = 256;
    const int* cur_cnt_dev = n_seeds_dev;
    const int64_t* cur_list = seeds;
    int cur_n = n_seeds;
    for (int h = 0; h < num_hops; ++h) {
        const bool last = (h == num_hops - 1);
        int64_t* nxt = last ? nullptr : frontier + (size_t)h * frontier_cap;
        int* nxt_cnt = last ? nullptr : frontier_cnt + h;
        const long long items = (h == 0) ? n_seeds : frontier_cap;
        const int grid = ctan_grid(items * lpi, threads);
        CTAN_LAUNCH((lk_expand_kernel), grid, threads, 0, stream, cur_list, cur_n, cur_cnt_dev, h == 0 ? n_seeds : frontier_cap,
                                                       nbr, deg, S, lpi, bitmap, cbitmap, nxt, nxt_cnt, frontier_cap,
                                                       counts + 3);
        CTAN_CHECK_LAUNCH();
        cur_list = nxt; cur_cnt_dev = nxt_cnt; cur_n = 0;
    }
    const int n_words = (int)((num_nodes + 31) / 32);
    // the frontier lists are dead once the last hop has run: their storage doubles as the scan's partial sums
    int* scan_partials = (frontier && (size_t)frontier_cap * sizeof(int64_t) >= SCAN_CTAS * 3 * sizeof(int))
                             ? reinterpret_cast<int*>(frontier) : nullptr;
    if (n_words <= 4096 || !scan_partials) {
        CTAN_LAUNCH((lk_scan_kernel), 1, 1024, 0, stream, bitmap, cbitmap, deg, n_words, n_id_out, assoc, rowptr, counts, n_cap,
                                               e_cap, frontier_cnt, num_hops > 1 ? num_hops - 1 : 0);
    } else {
        // large id spaces (e.g. 994k nodes = 31k bitmap words): two-level scan over SCAN_CTAS CTAs
        CTAN_LAUNCH((lk_scan_partial_kernel), SCAN_CTAS, 256, 0, stream, bitmap, cbitmap, deg, n_words, scan_partials);
        CTAN_CHECK_LAUNCH();
        CTAN_LAUNCH((lk_scan_emit_kernel), SCAN_CTAS, 256, 0, stream, bitmap, cbitmap, deg, n_words, scan_partials, n_id_out, assoc,
                                                         rowptr, counts, n_cap, e_cap, frontier_cnt,
                                                         num_hops > 1 ? num_hops - 1 : 0);
    }
    CTAN_CHECK_LAUNCH();
    return 0;
}

// Phase 2: write the edge list (edge_src = neighbour local id, edge_dst = centre local id, e_id) and
// clear the bitmaps.  n = counts[0] on the host, or n_dev = &counts[0] (then n is the grid bound).
extern "C" int ctan_nbr_fill(const int64_t* nbr, const int64_t* eid, int S, const int64_t* n_id,
                             const int32_t* rowptr, const int64_t* assoc, int n, const int32_t* n_dev,
                             int64_t* edge_src, int64_t* edge_dst, int64_t* e_id_out, uint32_t* bitmap,
                             uint32_t* cbitmap, cudaStream_t stream) {
    if (n <= 0) return 0;
    const int lpi = lanes_per_item(S);
    CTAN_LAUNCH((lk_fill_kernel), ctan_grid((long long)n * lpi, 256), 256, 0, stream, nbr, eid, S, lpi, n_id, rowptr, assoc, n,
                                                                           n_dev, edge_src, edge_dst, e_id_out,
                                                                           bitmap, cbitmap, nullptr, nullptr, 0.f, 1.f,
                                                                           nullptr);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// Same as ctan_nbr_fill, and also writes rel[e] = ((|last_update[src_e] - t_tab[e_id_e]| - mean)/std) as fp32
// (the delta-t input of the time encoder) while the row is in registers: saves a dependent launch.
extern "C" int ctan_nbr_fill_rel(const int64_t* nbr, const int64_t* eid, int S, const int64_t* n_id,
                                 const int32_t* rowptr, const int64_t* assoc, int n, const int32_t* n_dev,
                                 int64_t* edge_src, int64_t* edge_dst, int64_t* e_id_out, uint32_t* bitmap,
                                 uint32_t* cbitmap, const int64_t* last_update, const int64_t* t_tab, float mean,
                                 float stdv, float* rel, cudaStream_t stream) {
    if (n <= 0) return 0;
    const int lpi = lanes_per_item(S);
    CTAN_LAUNCH((lk_fill_kernel), ctan_grid((long long)n * lpi, 256), 256, 0, stream, nbr, eid, S, lpi, n_id, rowptr, assoc, n,
                                                                           n_dev, edge_src, edge_dst, e_id_out,
                                                                           bitmap, cbitmap, last_update, t_tab, mean,
                                                                           stdv, rel);
    CTAN_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------
// K6: memory write-back.  Entry p in [0,2B): node = cat(src,dst)[p], time = cat(t,t)[p], row =
// cat(src_emb,dst_emb)[p].  Entry p wins its node iff no other entry of the node has a larger time,
// or the same time and a smaller position (first arg-max; torch-scatter CPU).  One warp per entry:
// lanes scan the 2B (node,time) pairs, the winner copies its D floats and writes last_update.
// Rows come either from explicit [B,D] arrays, or (zero-sync engine) from z[assoc[node]].
// ------------------------------------------------------------------------------------------
__global__ void mem_update_kernel(float* __restrict__ mem, int64_t* __restrict__ last_update, int D,
                                  const int64_t* __restrict__ src, const int64_t* __restrict__ dst,
                                  const int64_t* __restrict__ t, int B, const float* __restrict__ emb_src,
                                  const float* __restrict__ emb_dst, int emb_ld, const float* __restrict__ z,
                                  const int64_t* __restrict__ assoc) {
    ctan_pdl_begin();
    const int wpb = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int n = 2 * B;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < n; p += gridDim.x * wpb) {
        const int64_t u = p < B ? src[p] : dst[p - B];
        const int64_t tp = t[p < B ? p : p - B];
        bool beaten = false;
        for (int q0 = 0; q0 < n; q0 += 32) {
            const int q = q0 + lane;
            bool b = false;
            if (q < n && q != p) {
                const int64_t v = q < B ? src[q] : dst[q - B];
                if (v == u) {
                    const int64_t tq = t[q < B ? q : q - B];
                    b = (tq > tp) || (tq == tp && q < p);
                }
            }
            if (__any_sync(0xffffffffu, b)) { beaten = true; break; }
        }
        if (beaten) continue;
        const float* row;
        if (z) row = z + (size_t)assoc[u] * D;
        else row = p < B ? emb_src + (size_t)p * emb_ld : emb_dst + (size_t)(p - B) * emb_ld;
        float* out = mem + (size_t)u * D;
        for (int d = lane; d < D; d += 32) out[d] = row[d];
        if (lane == 0) last_update[u] = tp;
    }
}

extern "C" int ctan_mem_update(float* mem, int64_t* last_update, int D, const int64_t* src, const int64_t* dst,
                               const int64_t* t, int B, const float* emb_src, const float* emb_dst, int emb_ld,
                               const float* z, const int64_t* assoc, cudaStream_t stream) {
    if (B < 0 || D <= 0) return CTAN_ERR_BAD_ARG;
    if (B == 0) return 0;
    if (!z && (!emb_src || !emb_dst)) return CTAN_ERR_BAD_ARG;
    const int wpb = 4;
    CTAN_LAUNCH((mem_update_kernel), ctan_grid(2LL * B, wpb), wpb * 32, 0, stream, mem, last_update, D, src, dst, t, B, emb_src,
                                                                       emb_dst, emb_ld > 0 ? emb_ld : D, z, assoc);
    CTAN_CHECK_LAUNCH();
    return 0;
}

__global__ void mem_reset_kernel(float* __restrict__ mem, int64_t* __restrict__ last_update, long long rows, int D,
                                 int64_t init_time) {
    ctan_pdl_begin();
    const long long n = rows * D;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) mem[i] = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < rows; i += stride)
        last_update[i] = init_time;
}

// SimpleMemory.reset_state (models/memory_layers.py:146-149)
extern "C" int ctan_mem_reset(float* mem, int64_t* last_update, int64_t num_nodes, int D, int64_t init_time,
                              cudaStream_t stream) {
    if (num_nodes <= 0 || D <= 0) return CTAN_ERR_BAD_ARG;
    CTAN_LAUNCH((mem_reset_kernel), ctan_grid(num_nodes * D, 256 * 4), 256, 0, stream, mem, last_update, num_nodes, D, init_time);
    CTAN_CHECK_LAUNCH();
    return 0;
}
