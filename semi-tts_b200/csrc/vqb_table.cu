// Codebook table assembly and its backward (reference: src/embed.py:109-112, :87-94).
#include <cuda_bf16.h>
#include "vqb_common.cuh"
#include "vqb_f16x2.cuh"

namespace vqb {

// One CTA per code row: table[k,:] = cat(learnable[k,:], attr[k,:] @ W^T + b); enorm[k] = |table[k,:]|^2.
__device__ __forceinline__ float table_entry(const float* __restrict__ learnable, const float* __restrict__ attr,
                                             const float* __restrict__ proj_w, const float* __restrict__ proj_b,
                                             int k, int d, int Dl, int A) {
    if (d < Dl) return learnable[(size_t)k * Dl + d];
    const float* w = proj_w + (size_t)(d - Dl) * A;
    const float* a = attr + (size_t)k * A;
    float acc = 0.f;
    for (int i = 0; i < A; ++i) acc = fmaf(a[i], w[i], acc);
    return acc + proj_b[d - Dl];
}

__global__ void __launch_bounds__(128)
assemble_table_kernel(const float* __restrict__ learnable, const float* __restrict__ attr,
                      const float* __restrict__ proj_w, const float* __restrict__ proj_b,
                      int K, int D, int A, int Da, float* __restrict__ table,
                      float* __restrict__ enorm, __nv_bfloat16* __restrict__ table_bf16) {
    const int k = blockIdx.x;
    const int Dl = D - Da;
    pdl_launch();                                      // the forward kernel may start its prologue (it waits before reading)
    float sq = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float v = table_entry(learnable, attr, proj_w, proj_b, k, d, Dl, A);
        table[(size_t)k * D + d] = v;
        if (table_bf16) table_bf16[(size_t)k * D + d] = __float2bfloat16_rn(v);
        sq = fmaf(v, v, sq);
    }
    __shared__ float red[4];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0 && enorm) enorm[k] = (red[0] + red[1]) + (red[2] + red[3]);
}

// Small tables (K <= 64, D <= 64: every semi-tts configuration) in ONE CTA: the table, |e|^2 and -- in the same launch --
// the fp16x2 operand image (vqb_f16x2.cuh) that the parity-mode forward and backward kernels consume.  The image needs
// the table-wide maximum, which a single CTA has without a grid-wide dependency.  The inputs are staged in shared memory
// with one round of coalesced loads, so the A-term projection sums (:110) never wait on global memory.
__global__ void __launch_bounds__(512)
assemble_small_kernel(const float* __restrict__ learnable, const float* __restrict__ attr,
                      const float* __restrict__ proj_w, const float* __restrict__ proj_b,
                      int K, int D, int A, int Da, float* __restrict__ table,
                      float* __restrict__ enorm, __nv_bfloat16* __restrict__ table_bf16, uint8_t* __restrict__ img) {
    __shared__ float s_tab[64 * 64];
    __shared__ float s_en[64];
    __shared__ float s_max[16], s_nrm[16];
    extern __shared__ float s_dyn[];                   // attr [K][A] | proj_w [Da][A] | proj_b [Da]
    float* s_attr = s_dyn;
    float* s_w = s_attr + K * A;
    float* s_b = s_w + Da * A;
    const int Dl = D - Da;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    pdl_launch();                                      // the forward kernel may start its prologue (it waits before reading)
    for (int i = threadIdx.x; i < K * Dl; i += blockDim.x) {
        const int k = i / Dl, d = i - k * Dl;
        s_tab[k * D + d] = learnable[i];
    }
    for (int i = threadIdx.x; i < K * A; i += blockDim.x) s_attr[i] = attr[i];
    for (int i = threadIdx.x; i < Da * A; i += blockDim.x) s_w[i] = proj_w[i];
    if ((int)threadIdx.x < Da) s_b[threadIdx.x] = proj_b[threadIdx.x];
    __syncthreads();
    for (int i = threadIdx.x; i < K * Da; i += blockDim.x) {       // projected columns, same fmaf order as table_entry
        const int k = i / Da, j = i - k * Da;
        const float* a = s_attr + k * A;
        const float* w = s_w + j * A;
        float acc = 0.f;
        for (int t = 0; t < A; ++t) acc = fmaf(a[t], w[t], acc);
        s_tab[k * D + Dl + j] = acc + s_b[j];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
        const float v = s_tab[i];
        table[i] = v;
        if (table_bf16) table_bf16[i] = __float2bfloat16_rn(v);
    }
    float gmax = 0.f, nmax = 0.f;
    for (int k = warp; k < K; k += nwarp) {            // one warp per row: |e|^2, the table maximum, the largest row norm
        float sq = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float v = s_tab[k * D + d];
            gmax = fmaxf(gmax, fabsf(v));
            sq = fmaf(v, v, sq);
        }
        sq = warp_sum(sq);
        if (lane == 0) { s_en[k] = sq; if (enorm) enorm[k] = sq; }
        nmax = fmaxf(nmax, sq);
    }
    gmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(gmax)));   // non-negative floats order like uints
    if (lane == 0) { s_max[warp] = gmax; s_nrm[warp] = nmax; }
    __syncthreads();
    gmax = 0.f; nmax = 0.f;
    for (int i = 0; i < nwarp; ++i) { gmax = fmaxf(gmax, s_max[i]); nmax = fmaxf(nmax, s_nrm[i]); }
    write_image(s_tab, D, K, D, gmax, sqrtf(nmax), s_en, img);
}

// d_learnable = eff[:, :Dl];  d_proj_w = eff[:, Dl:]^T @ attr;  d_proj_b = colsum(eff[:, Dl:])
// with eff = dtable + 2 * table * colsum[:, None] (the |e|^2 term of the L2 distance route).
__global__ void __launch_bounds__(256)
table_backward_kernel(const float* __restrict__ dtable, const float* __restrict__ table,
                      const float* __restrict__ colsum, const float* __restrict__ attr, int K, int D,
                      int A, int Da, int n_elem_blocks, float* __restrict__ d_learnable,
                      float* __restrict__ d_proj_w, float* __restrict__ d_proj_b) {
    const int Dl = D - Da;
    if ((int)blockIdx.x < n_elem_blocks) {
        const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < (int64_t)K * Dl) {
            const int k = (int)(i / Dl), d = (int)(i % Dl);
            float v = dtable[(size_t)k * D + d];
            if (colsum) v = fmaf(2.f * table[(size_t)k * D + d], colsum[k], v);
            d_learnable[i] = v;
        }
        return;
    }
    // projection part: one warp per output (j, a) / bias element j; lanes stride over the K codes
    const int o = (((int)blockIdx.x - n_elem_blocks) * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int n_w = Da * A;
    if (o >= n_w + Da) return;
    const int j = o < n_w ? o / A : o - n_w;
    const int a = o < n_w ? o % A : -1;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) {
        float v = dtable[(size_t)k * D + Dl + j];
        if (colsum) v = fmaf(2.f * table[(size_t)k * D + Dl + j], colsum[k], v);
        acc += (a >= 0) ? v * attr[(size_t)k * A + a] : v;
    }
    acc = warp_sum(acc);
    if (lane == 0) { if (a >= 0) d_proj_w[(size_t)j * A + a] = acc; else d_proj_b[j] = acc; }
}

}  // namespace vqb

using namespace vqb;

static bool small_table(int64_t n_codes, int64_t dim) { return n_codes <= 64 && (dim == 32 || dim == 64); }

extern "C" size_t vqb_operand_cache_bytes(int64_t n_codes, int64_t dim) {
    return small_table(n_codes, dim) ? (size_t)IMG_BYTES : 0;
}

extern "C" int vqb_assemble_table(const float* learnable, const float* phn_attr, const float* proj_w,
                                  const float* proj_b, int64_t n_codes, int64_t dim, int64_t n_attr,
                                  int64_t dim_attr, float* table, float* enorm, void* table_bf16,
                                  void* operand_cache, void* stream) {
    if (!learnable || !table) return invalid("vqb_assemble_table: learnable/table is NULL");
    if (n_codes <= 0 || dim <= 0) return invalid("vqb_assemble_table: bad shape K=%lld D=%lld",
                                                 (long long)n_codes, (long long)dim);
    const bool has_attr = phn_attr != nullptr;
    if (has_attr && (!proj_w || !proj_b || n_attr <= 0 || dim_attr <= 0 || dim_attr >= dim))
        return invalid("vqb_assemble_table: phn_attr given but projection is missing or 0 < D_a < D violated");
    if (!has_attr) { n_attr = 0; dim_attr = 0; }
    if (operand_cache && small_table(n_codes, dim) && n_attr <= 63) {
        const size_t dyn = (size_t)(n_codes * n_attr + dim_attr * n_attr + dim_attr) * 4;      // <= 32.5 KB
        assemble_small_kernel<<<1, 512, dyn, (cudaStream_t)stream>>>(
            learnable, phn_attr, proj_w, proj_b, (int)n_codes, (int)dim, (int)n_attr, (int)dim_attr, table,
            enorm, (__nv_bfloat16*)table_bf16, reinterpret_cast<uint8_t*>(operand_cache));
        VQB_CHECK_LAUNCH("assemble_small_kernel");
        return VQB_OK;
    }
    assemble_table_kernel<<<(unsigned)n_codes, 128, 0, (cudaStream_t)stream>>>(
        learnable, phn_attr, proj_w, proj_b, (int)n_codes, (int)dim, (int)n_attr, (int)dim_attr, table,
        enorm, (__nv_bfloat16*)table_bf16);
    VQB_CHECK_LAUNCH("assemble_table_kernel");
    return VQB_OK;
}

extern "C" int vqb_table_backward(const float* dtable, const float* table, const float* colsum,
                                  const float* phn_attr, int64_t n_codes, int64_t dim, int64_t n_attr,
                                  int64_t dim_attr, float* d_learnable, float* d_proj_w,
                                  float* d_proj_b, void* stream) {
    if (!dtable || !d_learnable) return invalid("vqb_table_backward: dtable/d_learnable is NULL");
    if (colsum && !table) return invalid("vqb_table_backward: colsum given without table");
    const bool has_attr = phn_attr != nullptr;
    if (has_attr && (!d_proj_w || !d_proj_b || n_attr <= 0 || dim_attr <= 0 || dim_attr >= dim))
        return invalid("vqb_table_backward: phn_attr given but projection outputs are missing");
    if (!has_attr) { n_attr = 0; dim_attr = 0; }
    const int64_t n_elem = n_codes * (dim - dim_attr);
    const int n_elem_blocks = (int)ceil_div(n_elem, 256);
    const int n_proj_blocks = has_attr ? (int)ceil_div((dim_attr * n_attr + dim_attr) * 32, 256) : 0;
    table_backward_kernel<<<n_elem_blocks + n_proj_blocks, 256, 0, (cudaStream_t)stream>>>(
        dtable, table, colsum, phn_attr, (int)n_codes, (int)dim, (int)n_attr, (int)dim_attr,
        n_elem_blocks, d_learnable, d_proj_w, d_proj_b);
    VQB_CHECK_LAUNCH("table_backward_kernel");
    return VQB_OK;
}
