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

int launch_scatter_add(const int64_t* idx, int64_t n, const float* g, int64_t K, int64_t D, float* dtable,
                       int64_t* hist, cudaStream_t s) {
    if (n == 0) return VQB_OK;
    if (D % 4 != 0) return invalid("scatter_add: D must be a multiple of 4 (got %lld)", (long long)D);
    // per-warp private copies only pay off while 8 warps of them fit (small codebooks, e.g. K=43: 11 KB each);
    // larger tables go straight to 128-bit global reductions
    const size_t per_warp = (size_t)K * D * 4;
    const int wpb = 8;
    if (g && per_warp * wpb <= 96 * 1024)
        return launch_scatter_v<true>((const long long*)idx, n, g, (int)K, (int)D, dtable,
                                      (unsigned long long*)hist, wpb, per_warp * wpb, s);
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

extern "C" int vqb_scatter_add(const int64_t* txt, int64_t n_tokens, const float* g, int64_t n_codes, int64_t dim,
                               float* dtable, int64_t* hist, void* stream) {
    if (n_tokens == 0) return VQB_OK;
    if (!txt || (g && !dtable)) return invalid("vqb_scatter_add: NULL pointer");
    if (g && (!aligned16(g) || !aligned16(dtable))) return invalid("vqb_scatter_add: pointers must be 16-byte aligned");
    return launch_scatter_add(txt, n_tokens, g, n_codes, dim, dtable, hist, (cudaStream_t)stream);
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
