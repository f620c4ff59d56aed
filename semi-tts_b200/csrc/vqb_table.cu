// Codebook table assembly and its backward (reference: src/embed.py:109-112, :87-94).
#include <cuda_bf16.h>
#include "vqb_common.cuh"

namespace vqb {

// One CTA per code row: table[k,:] = cat(learnable[k,:], attr[k,:] @ W^T + b); enorm[k] = |table[k,:]|^2.
// With an operand cache the same launch also writes the tf32 hi/lo operand copies used by the tensor-core
// kernels (forward: -2 e with |e|^2 as the bias block; backward: e), including the padded rows K..Kp-1.
__global__ void __launch_bounds__(128)
assemble_table_kernel(const float* __restrict__ learnable, const float* __restrict__ attr,
                      const float* __restrict__ proj_w, const float* __restrict__ proj_b,
                      int K, int D, int A, int Da, float* __restrict__ table,
                      float* __restrict__ enorm, __nv_bfloat16* __restrict__ table_bf16,
                      float* __restrict__ f_hi, float* __restrict__ f_lo, float* __restrict__ b_hi,
                      float* __restrict__ b_lo) {
    const int k = blockIdx.x;
    const int Dl = D - Da;
    pdl_launch();                                      // the forward kernel may start its prologue (it waits before reading)
    if (k >= K) {                                      // padded operand rows (only launched with a cache)
        for (int d = threadIdx.x; d < D + 32; d += blockDim.x) {
            f_hi[(size_t)k * (D + 32) + d] = d == D ? 1e30f : 0.f;
            b_hi[(size_t)k * (D + 32) + d] = 0.f;
            if (d < D) { f_lo[(size_t)k * D + d] = 0.f; b_lo[(size_t)k * D + d] = 0.f; }
        }
        return;
    }
    float sq = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float v;
        if (d < Dl) {
            v = learnable[(size_t)k * Dl + d];
        } else {
            const float* w = proj_w + (size_t)(d - Dl) * A;
            const float* a = attr + (size_t)k * A;
            float acc = 0.f;
            for (int i = 0; i < A; ++i) acc = fmaf(a[i], w[i], acc);
            v = acc + proj_b[d - Dl];
        }
        table[(size_t)k * D + d] = v;
        if (table_bf16) table_bf16[(size_t)k * D + d] = __float2bfloat16_rn(v);
        if (f_hi) {
            const float m2 = -2.f * v, h = tf32_rn(m2), hb = tf32_rn(v);
            f_hi[(size_t)k * (D + 32) + d] = h;  f_lo[(size_t)k * D + d] = m2 - h;
            b_hi[(size_t)k * (D + 32) + d] = hb; b_lo[(size_t)k * D + d] = v - hb;
        }
        sq = fmaf(v, v, sq);
    }
    __shared__ float red[4];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    const float ee = (red[0] + red[1]) + (red[2] + red[3]);
    if (threadIdx.x == 0 && enorm) enorm[k] = ee;
    if (f_hi && threadIdx.x < 32) {                    // bias block: |e|^2 as three tf32-exact words, then zeros
        const float b0 = tf32_trunc(ee), r1 = ee - b0, b1 = tf32_trunc(r1), b2 = r1 - b1;
        const int j = threadIdx.x;
        f_hi[(size_t)k * (D + 32) + D + j] = j == 0 ? b0 : (j == 1 ? b1 : (j == 2 ? b2 : 0.f));
        b_hi[(size_t)k * (D + 32) + D + j] = 0.f;
    }
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

extern "C" size_t vqb_operand_cache_bytes(int64_t n_codes, int64_t dim) {
    return 2 * (cache_hi_bytes(n_codes, dim) + cache_lo_bytes(n_codes, dim));
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
    float *f_hi = nullptr, *f_lo = nullptr, *b_hi = nullptr, *b_lo = nullptr;
    unsigned rows = (unsigned)n_codes;
    if (operand_cache) {
        uint8_t* oc = reinterpret_cast<uint8_t*>(operand_cache);
        const size_t hb = cache_hi_bytes(n_codes, dim), lb = cache_lo_bytes(n_codes, dim);
        f_hi = reinterpret_cast<float*>(oc);          f_lo = reinterpret_cast<float*>(oc + hb);
        b_hi = reinterpret_cast<float*>(oc + hb + lb); b_lo = reinterpret_cast<float*>(oc + 2 * hb + lb);
        rows = (unsigned)cache_rows(n_codes);
    }
    assemble_table_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(
        learnable, phn_attr, proj_w, proj_b, (int)n_codes, (int)dim, (int)n_attr, (int)dim_attr, table,
        enorm, (__nv_bfloat16*)table_bf16, f_hi, f_lo, b_hi, b_lo);
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
