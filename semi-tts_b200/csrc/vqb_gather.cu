// Gather-only / scatter-only paths: inference gather (src/embed.py:96-103, :180-185), the
// index-keyed codebook-gradient scatter-add fused with the usage histogram (autograd of the
// F.embedding at src/embed.py:134; histogram semantics of bin/train_vqvae.py:256-261), and the
// backward of the loss extensions.
#include "vqb_common.cuh"

namespace vqb {

// out[m,:] = table[clamp(txt[m]),:] -- one warp per token row, 128-bit lanes
__global__ void __launch_bounds__(256)
gather_rows_kernel(const long long* __restrict__ txt, long long n, const float* __restrict__ table,
                   int K, int D, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int D4 = D >> 2;
    for (long long m = warp; m < n; m += nwarps) {
        long long k = txt[m];
        k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        const float* src = table + (size_t)k * D;
        float* dst = out + (size_t)m * D;
        for (int c = lane; c < D4; c += 32) stg4_stream(dst + 4 * c, ldg4(src + 4 * c));
    }
}

// -------------------------------------------------------------------------------------------------
// dtable[idx[n],:] += g[n,:], hist[idx[n]] += 1.
//
// Rows are split into contiguous spans, one per warp (encoder frames arrive in time order, so equal
// indices come in runs).  A warp walks its span with lanes across the D columns (128-bit loads,
// fully coalesced), keeps the running sum of the current run in registers and only emits it when
// the index changes (segmented reduction: one flush per run, not per row).
//  * SMEM variant (K*D*4 <= budget): flush = plain read-modify-write into a per-WARP private copy of
//    the table gradient in shared memory (no atomics at all inside the loop); the copies are summed
//    and sent to global memory once per CTA with red.global.add.v4.f32.
//  * GLOBAL variant (large codebooks): flush = red.global.add.v4.f32 straight to dtable.
// -------------------------------------------------------------------------------------------------
template <bool SMEM, int VPL>   // VPL = float4 per lane (D <= 128 * VPL)
__global__ void __launch_bounds__(256)
scatter_hist_kernel(const long long* __restrict__ idx, long long n, const float* __restrict__ g,
                    int K, int D, float* __restrict__ dtable, unsigned long long* __restrict__ hist,
                    int rows_per_warp) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int D4 = D >> 2;
    float* mine = smem + (size_t)wib * K * D;
    if (SMEM) {
        for (int i = threadIdx.x; i < wpb * K * D; i += blockDim.x) smem[i] = 0.f;
        __syncthreads();
    }
    const long long warp = (long long)blockIdx.x * wpb + wib;
    const long long beg = warp * rows_per_warp;
    const long long end = min(n, beg + (long long)rows_per_warp);

    float4 run[VPL];
#pragma unroll
    for (int v = 0; v < VPL; ++v) run[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    long long cur = -1;
    unsigned long long run_len = 0;

    auto flush = [&]() {
        if (cur < 0) return;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            const int c = lane + 32 * v;
            if (c < D4) {
                if (SMEM) {
                    float4* dst = reinterpret_cast<float4*>(mine + (size_t)cur * D) + c;
                    float4 a = *dst;
                    a.x += run[v].x; a.y += run[v].y; a.z += run[v].z; a.w += run[v].w;
                    *dst = a;
                } else {
                    red_add_v4(dtable + (size_t)cur * D + 4 * c, run[v]);
                }
            }
            run[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (hist && lane == 0) atomicAdd(hist + cur, run_len);
        run_len = 0;
    };

    for (long long r = beg; r < end; ++r) {
        long long k = idx[r];                         // warp-uniform (broadcast load)
        k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        if (k != cur) { flush(); cur = k; }
        ++run_len;
        if (g) {
            const float* src = g + (size_t)r * D;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = lane + 32 * v;
                if (c < D4) {
                    const float4 x = ldg4_stream(src + 4 * c);
                    run[v].x += x.x; run[v].y += x.y; run[v].z += x.z; run[v].w += x.w;
                }
            }
        }
    }
    flush();
    if (SMEM && g) {
        __syncthreads();
        const int KD4 = (K * D) >> 2;
        for (int i = threadIdx.x; i < KD4; i += blockDim.x) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int w = 0; w < wpb; ++w) {
                const float4 b = reinterpret_cast<const float4*>(smem + (size_t)w * K * D)[i];
                a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
            }
            if (a.x != 0.f || a.y != 0.f || a.z != 0.f || a.w != 0.f) red_add_v4(dtable + 4 * (size_t)i, a);
        }
    }
}

template <bool SMEM>
static int launch_scatter_v(const long long* idx, long long n, const float* g, int K, int D, float* dtable,
                            unsigned long long* hist, int wpb, size_t smem, cudaStream_t s) {
    const int D4 = D / 4;
    const int vpl = (D4 + 31) / 32;
    // spans: enough warps to fill the machine ~4x over, but at least 32 rows per warp so runs can form
    const long long target_warps = (long long)sm_count() * 32;
    long long rpw = ceil_div(n, target_warps);
    if (rpw < 32) rpw = 32;
    const long long nwarps = ceil_div(n, rpw);
    const unsigned grid = (unsigned)ceil_div(nwarps, wpb);
#define VQB_SC(V)                                                                                  \
    {                                                                                              \
        auto kern = scatter_hist_kernel<SMEM, V>;                                                  \
        if (smem > 48 * 1024)                                                                      \
            VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, wpb * 32, smem, s>>>(idx, n, g, K, D, dtable, hist, (int)rpw);                \
    }
    if (vpl <= 1) VQB_SC(1) else if (vpl <= 2) VQB_SC(2) else if (vpl <= 4) VQB_SC(4)
    else return invalid("scatter_add: D=%d is not supported (D <= 512)", D);
#undef VQB_SC
    VQB_CHECK_LAUNCH("scatter_hist_kernel");
    return VQB_OK;
}


// -------------------------------------------------------------------------------------------------
// Large codebooks (the table does not fit per-warp shared-memory copies): sort-free segmented sum.
//   rank_kernel      rank[r] = position of row r among the rows of its code (atomic ticket per code, which also IS
//                    the usage histogram of this call);  warp-aggregated: lanes of a warp that hit the same code
//                    take one atomic between them
//   offsets_kernel   exclusive scan of the per-code counts -> first slot of each code in the permutation, and the
//                    work list: one item per (code, chunk of <= CHUNK rows)
//   permute_kernel   perm[offset[idx[r]] + rank[r]] = r
//   segsum_kernel    one warp per work item: gathers its rows (whole 16-byte-vectorised rows, several in flight),
//                    adds them in registers and writes the code's row of dtable once -- a plain store when the code
//                    has a single chunk, one red.global.add.v4 per chunk otherwise.  No atomics per input row.
// HBM traffic: g once (4D B/row) + idx twice + 8 B/row of rank/perm, against 4D + 8 algorithmic bytes.
// The order of the rows inside a code follows the atomic tickets, so the fp32 sums of this path are not
// bit-reproducible from run to run (the small-codebook path above and the fused backward are).
// -------------------------------------------------------------------------------------------------
constexpr int SEG_CHUNK = 128;

__global__ void __launch_bounds__(256)
rank_kernel(const long long* __restrict__ idx, long long n, int K, int* __restrict__ cnt, int* __restrict__ rank) {
    const int lane = threadIdx.x & 31;
    pdl_launch();
    // (the loop bound is warp-uniform: all 32 lanes take part in every match / shuffle.  Four rows per thread and step with
    //  the tickets issued back to back was measured and is NOT faster -- 29.1 vs 26.7 us at N = 2^20, K = 8192: the returning
    //  atomics are bound by L2 throughput, not by their latency)
    const long long T = (long long)gridDim.x * blockDim.x;
    for (long long r0 = (long long)blockIdx.x * blockDim.x + threadIdx.x - lane; r0 < n; r0 += T) {
        const long long r = r0 + lane;
        const bool in = r < n;
        long long k = in ? idx[r] : -1;
        if (in) k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        // warp-aggregated ticket: lanes holding the same code form a group; its leader takes `size` tickets at once
        const unsigned peers = __match_any_sync(0xffffffffu, (int)k);
        const int leader = __ffs(peers) - 1;
        const int pos = __popc(peers & ((1u << lane) - 1u));
        int base = 0;
        if (in && lane == leader) base = atomicAdd(cnt + k, __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (in) rank[r] = base + pos;
    }
}

// Small codebooks (K <= RANK_SMEM_K): the global ticket counters would be hit by thousands of rows each.  A block
// takes tickets for a span of RANK_SPAN rows in shared memory first (shared-memory atomics), then reserves, per code
// present in the span, one contiguous range of global tickets with a single atomic.
constexpr int RANK_SMEM_K = 2048;
constexpr int RANK_SPAN = 256 * 16;

__global__ void __launch_bounds__(256)
rank_smem_kernel(const long long* __restrict__ idx, long long n, int K, int* __restrict__ cnt, int* __restrict__ rank) {
    __shared__ int s_cnt[RANK_SMEM_K];
    pdl_launch();
    const long long n_spans = (n + RANK_SPAN - 1) / RANK_SPAN;
    for (long long span = blockIdx.x; span < n_spans; span += gridDim.x) {
        const long long r0 = span * RANK_SPAN;
        for (int k = threadIdx.x; k < K; k += 256) s_cnt[k] = 0;
        __syncthreads();
        int code[16], local[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const long long r = r0 + u * 256 + threadIdx.x;
            code[u] = -1;
            if (r < n) {
                long long k = idx[r];
                k = k < 0 ? 0 : (k >= K ? K - 1 : k);
                code[u] = (int)k;
            }
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) local[u] = code[u] >= 0 ? atomicAdd(&s_cnt[code[u]], 1) : 0;
        __syncthreads();
        for (int k = threadIdx.x; k < K; k += 256) {
            const int c = s_cnt[k];
            if (c) s_cnt[k] = atomicAdd(cnt + k, c);              // the span's first global ticket for code k
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const long long r = r0 + u * 256 + threadIdx.x;
            if (code[u] >= 0) rank[r] = s_cnt[code[u]] + local[u];
        }
        __syncthreads();
    }
}

// single CTA: offs[k] = sum_{j<k} cnt[j];  item_off[k] = sum_{j<k} ceil(cnt[j] / CHUNK) (first work item of code k);
// n_items[0] = number of work items
__global__ void __launch_bounds__(1024)
offsets_kernel(const int* __restrict__ cnt, int K, int* __restrict__ offs, int* __restrict__ item_off,
               int* __restrict__ n_items, unsigned long long* __restrict__ hist) {
    __shared__ int s_w[32], s_w2[32];
    __shared__ int s_carry[2];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    constexpr int E = 8;                                       // consecutive codes per thread: K <= 8192 is ONE pass (one round of
    pdl_launch();                                              // global loads, one block scan); measured neutral against 8 passes
    pdl_wait();                                                // the ticket kernel has completed
    if (threadIdx.x == 0) { s_carry[0] = 0; s_carry[1] = 0; }
    __syncthreads();
    for (int k0 = 0; k0 < K; k0 += 1024 * E) {
        const int kb = k0 + threadIdx.x * E;
        int c[E], m[E];
        if (kb + E <= K && (reinterpret_cast<uintptr_t>(cnt) & 15) == 0) {
            const int4 v0 = *reinterpret_cast<const int4*>(cnt + kb), v1 = *reinterpret_cast<const int4*>(cnt + kb + 4);
            c[0] = v0.x; c[1] = v0.y; c[2] = v0.z; c[3] = v0.w; c[4] = v1.x; c[5] = v1.y; c[6] = v1.z; c[7] = v1.w;
        } else {
#pragma unroll
            for (int j = 0; j < E; ++j) c[j] = kb + j < K ? cnt[kb + j] : 0;
        }
        int ta = 0, tb = 0;                                    // this thread's totals
#pragma unroll
        for (int j = 0; j < E; ++j) { m[j] = (c[j] + SEG_CHUNK - 1) / SEG_CHUNK; ta += c[j]; tb += m[j]; }
        int a = ta, b = tb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int x = __shfl_up_sync(0xffffffffu, a, o), y = __shfl_up_sync(0xffffffffu, b, o);
            if (lane >= o) { a += x; b += y; }
        }
        if (lane == 31) { s_w[w] = a; s_w2[w] = b; }
        __syncthreads();
        int ba = s_carry[0], bb = s_carry[1];
        for (int i = 0; i < w; ++i) { ba += s_w[i]; bb += s_w2[i]; }
        int off = ba + a - ta, it0 = bb + b - tb;              // exclusive prefix of this thread's first code
#pragma unroll
        for (int j = 0; j < E; ++j) {
            if (kb + j < K) {
                offs[kb + j] = off;
                item_off[kb + j] = it0;
                // (plain read-modify-write: every code has exactly one writer, and calls are ordered by the stream)
                if (hist && c[j]) hist[kb + j] += (unsigned long long)c[j];
            }
            off += c[j]; it0 += m[j];
        }
        __syncthreads();
        if (threadIdx.x == 1023) { s_carry[0] = ba + a; s_carry[1] = bb + b; }
        __syncthreads();
    }
    if (threadIdx.x == 0) n_items[0] = s_carry[1];
}

__global__ void __launch_bounds__(256)
permute_kernel(const long long* __restrict__ idx, long long n, int K, const int* __restrict__ offs,
               const int* __restrict__ rank, int* __restrict__ perm) {
    pdl_launch();
    pdl_wait();                                                // offsets (and, before them, the tickets) are final
    for (long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (long long)gridDim.x * blockDim.x) {
        long long k = idx[r];
        k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        perm[__ldg(offs + k) + rank[r]] = (int)r;
    }
}

// LPR lanes share one row (LPR * VPL float4 >= D/4); 32 / LPR rows are in flight per warp step, U steps unrolled
template <int LPR, int VPL>
__global__ void __launch_bounds__(256)
segsum_kernel(const float* __restrict__ g, const int* __restrict__ perm, const int* __restrict__ cnt,
              const int* __restrict__ offs, const int* __restrict__ item_off, const int* __restrict__ n_items,
              int K, int D, float* __restrict__ dtable) {
    constexpr int RPW = 32 / LPR;                  // rows per warp step
    constexpr int U = 4;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, col = lane % LPR;
    const int D4 = D >> 2;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    pdl_wait();                                                // the permutation is complete
    const int total = __ldg(n_items);
    for (int item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < total; item += warps) {
        // code of this item: the last k with item_off[k] <= item (codes without rows share their successor's offset)
        int lo = 0, hi = K - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (__ldg(item_off + mid) <= item) lo = mid; else hi = mid - 1;
        }
        const int k = lo;
        const int code_off = __ldg(offs + k), code_cnt = __ldg(cnt + k);
        const int first = code_off + (item - __ldg(item_off + k)) * SEG_CHUNK;
        const int len = min(SEG_CHUNK, code_off + code_cnt - first);
        float4 acc[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i0 = 0; i0 < len; i0 += U * RPW) {
            int row[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * RPW + sub;
                row[u] = i < len ? __ldg(perm + first + i) : -1;
            }
            float4 x[U][VPL];
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    const int c = col + LPR * v;
                    x[u][v] = (row[u] >= 0 && c < D4) ? ldg4_stream(g + (size_t)row[u] * D + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int v = 0; v < VPL; ++v) { acc[v].x += x[u][v].x; acc[v].y += x[u][v].y; acc[v].z += x[u][v].z; acc[v].w += x[u][v].w; }
        }
        // fold the row groups of the warp together
#pragma unroll
        for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                acc[v].x += __shfl_xor_sync(0xffffffffu, acc[v].x, o); acc[v].y += __shfl_xor_sync(0xffffffffu, acc[v].y, o);
                acc[v].z += __shfl_xor_sync(0xffffffffu, acc[v].z, o); acc[v].w += __shfl_xor_sync(0xffffffffu, acc[v].w, o);
            }
        if (sub == 0) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                const int c = col + LPR * v;
                if (c < D4) {
                    float* dst = dtable + (size_t)k * D + 4 * c;
                    red_add_v4(dst, acc[v]);          // dtable is an accumulation target (+=) in every route
                }
            }
        }
    }
}

static size_t sorted_ws_bytes(long long n, long long K) { return (3 * (size_t)K + 4 + 2 * (size_t)n) * sizeof(int); }

static int launch_scatter_sorted(const long long* idx, long long n, const float* g, int K, int D, float* dtable,
                                 unsigned long long* hist, void* workspace, cudaStream_t s) {
    // caller-provided scratch: cnt[K] | offs[K] | n_items[4] | rank[n] | perm[n] | item_off[K];
    // at most m = n / CHUNK + K work items
    const size_t m = (size_t)(n / SEG_CHUNK) + (size_t)K + 1;
    int* ws = reinterpret_cast<int*>(workspace);
    int* cnt = ws; int* offs = cnt + K; int* n_items = offs + K; int* rank = n_items + 4; int* perm = rank + n;
    int* item_off = perm + n;
    VQB_CUDA(cudaMemsetAsync(cnt, 0, (size_t)K * sizeof(int), s));
    const long long gcap = (long long)sm_count() * 16;
    const unsigned grid = (unsigned)(ceil_div(n, 256) < gcap ? ceil_div(n, 256) : gcap);
    // the four kernels are chained with programmatic dependent launch (each one is this library's own predecessor):
    // launch latencies overlap, every kernel waits for its predecessor's completion before it reads its results
    kernel_event_begin(s);
    if (K <= RANK_SMEM_K) {
        const long long spans = ceil_div(n, RANK_SPAN);
        rank_smem_kernel<<<(unsigned)(spans < gcap ? spans : gcap), 256, 0, s>>>(idx, n, K, cnt, rank);
    } else {
        rank_kernel<<<grid, 256, 0, s>>>(idx, n, K, cnt, rank);
    }
    VQB_CHECK_LAUNCH("rank_kernel");
    VQB_CUDA(launch_pdl(offsets_kernel, dim3(1), dim3(1024), 0, s, (const int*)cnt, K, offs, item_off, n_items, hist));
    VQB_CHECK_LAUNCH("offsets_kernel");
    VQB_CUDA(launch_pdl(permute_kernel, dim3(grid), dim3(256), 0, s, idx, n, K, (const int*)offs, (const int*)rank, perm));
    VQB_CHECK_LAUNCH("permute_kernel");
    const long long scap = (long long)sm_count() * 8;
    const unsigned sgrid = (unsigned)(ceil_div((long long)m, 8) < scap ? ceil_div((long long)m, 8) : scap);
    const int D4 = D / 4;
#define VQB_SS(L, V) VQB_CUDA(launch_pdl(segsum_kernel<L, V>, dim3(sgrid), dim3(256), 0, s, g, (const int*)perm, (const int*)cnt, \
                                         (const int*)offs, (const int*)item_off, (const int*)n_items, K, D, dtable))
    if (D4 <= 4) VQB_SS(4, 1); else if (D4 <= 8) VQB_SS(8, 1); else if (D4 <= 16) VQB_SS(16, 1);
    else if (D4 <= 32) VQB_SS(32, 1); else if (D4 <= 64) VQB_SS(32, 2); else VQB_SS(32, 4);
#undef VQB_SS
    VQB_CHECK_LAUNCH("segsum_kernel");
    kernel_event_end(s);
    return VQB_OK;
}

// per-warp private copies only pay off while 8 warps of them fit (small codebooks, e.g. K=43: 11 KB each)
static bool scatter_fits_smem(int64_t K, int64_t D) { return (size_t)K * D * 4 * 8 <= 96 * 1024; }

size_t scatter_workspace_bytes(int64_t n, int64_t K, int64_t D) {
    return (!scatter_fits_smem(K, D) && n >= 65536) ? sorted_ws_bytes(n, K) : 0;
}

int launch_scatter_add(const int64_t* idx, int64_t n, const float* g, int64_t K, int64_t D, float* dtable,
                       int64_t* hist, void* workspace, size_t workspace_bytes, cudaStream_t s) {
    if (n == 0) return VQB_OK;
    if (D % 4 != 0) return invalid("scatter_add: D must be a multiple of 4 (got %lld)", (long long)D);
    // per-warp private copies only pay off while 8 warps of them fit (small codebooks, e.g. K=43: 11 KB each);
    // larger tables go straight to 128-bit global reductions
    const size_t per_warp = (size_t)K * D * 4;
    const int wpb = 8;
    if (g && per_warp * wpb <= 96 * 1024)
        return launch_scatter_v<true>((const long long*)idx, n, g, (int)K, (int)D, dtable,
                                      (unsigned long long*)hist, wpb, per_warp * wpb, s);
    // large tables: ticket + permutation + one gather-sum per code (no atomics per row).  Short inputs stay on the
    // direct 128-bit global reductions: four extra launches would cost more than they save.
    static const bool direct = getenv("VQB_SCATTER_DIRECT") != nullptr;       // developer switch (A/B)
    const size_t need = scatter_workspace_bytes(n, K, D);
    if (g && !direct && need && workspace && workspace_bytes >= need)
        return launch_scatter_sorted((const long long*)idx, n, g, (int)K, (int)D, dtable, (unsigned long long*)hist, workspace, s);
    return launch_scatter_v<false>((const long long*)idx, n, g, (int)K, (int)D, dtable,
                                   (unsigned long long*)hist, 8, 0, s);
}

// dx[n,:] (+)= gc * 2 (x - c) / (N D);  dtable[idx[n],:] += gv * 2 (c - x) / (N D)
__global__ void __launch_bounds__(256)
loss_backward_kernel(const float* __restrict__ x, const float* __restrict__ table, const long long* __restrict__ idx,
                     long long n, int D, int K, const float* __restrict__ g_vq, const float* __restrict__ g_commit,
                     float* __restrict__ dx, int dx_acc, float* __restrict__ dtable) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const float scale = 2.f / ((float)n * (float)D);
    const float gv = g_vq ? __ldg(g_vq) * scale : 0.f;
    const float gc = g_commit ? __ldg(g_commit) * scale : 0.f;
    const int D4 = D >> 2;
    for (long long r = warp; r < n; r += nwarps) {
        long long k = idx[r];
        k = k < 0 ? 0 : (k >= K ? K - 1 : k);
        for (int c = lane; c < D4; c += 32) {
            const float4 xv = ldg4_stream(x + (size_t)r * D + 4 * c);
            const float4 cv = ldg4(table + (size_t)k * D + 4 * c);
            const float4 df = make_float4(xv.x - cv.x, xv.y - cv.y, xv.z - cv.z, xv.w - cv.w);
            if (dx) {
                float4 o = make_float4(gc * df.x, gc * df.y, gc * df.z, gc * df.w);
                float* dp = dx + (size_t)r * D + 4 * c;
                if (dx_acc) { const float4 old = *reinterpret_cast<const float4*>(dp); o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
                *reinterpret_cast<float4*>(dp) = o;
            }
            if (dtable && gv != 0.f)
                red_add_v4(dtable + (size_t)k * D + 4 * c, make_float4(-gv * df.x, -gv * df.y, -gv * df.z, -gv * df.w));
        }
    }
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_inference_gather(const int64_t* txt, int64_t n_tokens, const float* table, int64_t n_codes,
                                    int64_t dim, float* out, void* stream) {
    if (n_tokens == 0) return VQB_OK;
    if (!txt || !table || !out) return invalid("vqb_inference_gather: NULL pointer");
    if (dim % 4 != 0 || !aligned16(table) || !aligned16(out))
        return invalid("vqb_inference_gather: D must be a multiple of 4 and pointers 16-byte aligned");
    const int64_t blocks = ceil_div(n_tokens, 8);
    const unsigned grid = (unsigned)(blocks < (int64_t)sm_count() * 16 ? blocks : (int64_t)sm_count() * 16);
    gather_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const long long*)txt, n_tokens, table,
                                                              (int)n_codes, (int)dim, out);
    VQB_CHECK_LAUNCH("gather_rows_kernel");
    return VQB_OK;
}

extern "C" int vqb_scatter_workspace(int64_t n_tokens, int64_t n_codes, int64_t dim, size_t* bytes) {
    if (!bytes) return invalid("vqb_scatter_workspace: bytes is NULL");
    *bytes = (n_tokens > 0 && n_codes > 0 && dim > 0) ? scatter_workspace_bytes(n_tokens, n_codes, dim) : 0;
    return VQB_OK;
}

extern "C" int vqb_scatter_add(const int64_t* txt, int64_t n_tokens, const float* g, int64_t n_codes, int64_t dim,
                               float* dtable, int64_t* hist, void* workspace, size_t workspace_bytes, void* stream) {
    if (n_tokens == 0) return VQB_OK;
    if (!txt || (g && !dtable)) return invalid("vqb_scatter_add: NULL pointer");
    if (g && (!aligned16(g) || !aligned16(dtable))) return invalid("vqb_scatter_add: pointers must be 16-byte aligned");
    return launch_scatter_add(txt, n_tokens, g, n_codes, dim, dtable, hist, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int vqb_loss_backward(const float* x, const float* table, const int64_t* idx, int64_t n_rows, int64_t dim,
                                 int64_t n_codes, const float* g_vq, const float* g_commit, float* dx,
                                 int dx_accumulate, float* dtable, void* stream) {
    if (n_rows == 0) return VQB_OK;
    if (!x || !table || !idx) return invalid("vqb_loss_backward: NULL pointer");
    if (dim % 4 != 0) return invalid("vqb_loss_backward: D must be a multiple of 4");
    const int64_t blocks = ceil_div(n_rows, 8);
    const unsigned grid = (unsigned)(blocks < (int64_t)sm_count() * 16 ? blocks : (int64_t)sm_count() * 16);
    loss_backward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, table, (const long long*)idx, n_rows, (int)dim,
                                                                (int)n_codes, g_vq, g_commit, dx, dx_accumulate, dtable);
    VQB_CHECK_LAUNCH("loss_backward_kernel");
    return VQB_OK;
}
