// Exact-fp32 forward of the quantizer on CUDA cores ("parity mode" and the generic fallback).
//
// Replaces the ATen sequence of L2Embedding.forward (src/embed.py:105-147: neg_batch_l2 :208-213,
// temperature :115-124, softmax :127, argmax :130, gather :134, straight-through :145) and of
// SeperateEmbedding.forward (:187-205) with ONE kernel: the N x K score matrix lives in registers,
// p_code is written once, the codeword gather / straight-through / usage histogram / squared-error
// sum are fused behind it.
//
// Mapping: a CTA owns a tile of TILE=128 rows; the x tile is staged in shared memory with coalesced
// 128-bit loads; thread t then owns row t (its K scores are thread-local, so softmax/argmax need no
// shuffles) and reads codebook values as shared-memory broadcasts.  Outputs go back through shared
// memory so that global stores are coalesced 128-bit as well.
#include <math.h>
#include "vqb_common.cuh"

namespace vqb {

constexpr int TILE = 128;

struct FwdP {
    const float* x; const float* w; const float* b; const float* tab; const float* temp;
    float* p; long long* idx; float* q; unsigned long long* hist; double* sqerr;
    int N, D, K; unsigned flags;
};

// cooperative, coalesced load of the x tile (rows beyond N are zero-filled)
__device__ __forceinline__ void load_x_tile(const FwdP& p, int row0, int rows, float* sX, int XS) {
    const int D4 = p.D >> 2;
    for (int i = threadIdx.x; i < TILE * D4; i += TILE) {
        const int r = i / D4, c = i - r * D4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows) v = ldg4_stream(p.x + (size_t)(row0 + r) * p.D + 4 * c);
        *reinterpret_cast<float4*>(sX + r * XS + 4 * c) = v;
    }
}

// codebook chunk [k0, k0+KC) -> sW4[d4][k] (float4 per (d4,k)), zero padded; sB[k]
template <int KC>
__device__ __forceinline__ void load_w_chunk(const FwdP& p, int k0, float4* sW4, float* sB) {
    const int D4 = p.D >> 2;
    for (int i = threadIdx.x; i < KC * D4; i += TILE) {
        const int k = i / D4, c = i - k * D4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + k < p.K) v = ldg4(p.w + (size_t)(k0 + k) * p.D + 4 * c);
        sW4[c * KC + k] = v;
    }
    for (int k = threadIdx.x; k < KC; k += TILE) sB[k] = (k0 + k < p.K) ? __ldg(p.b + k0 + k) : 0.f;
}

// acc[k] = x_row . w_k for the KC codes of the chunk; returns |x_row|^2
template <int KC>
__device__ __forceinline__ float dot_chunk(const float* xr, const float4* sW4, int D, float (&acc)[KC]) {
#pragma unroll
    for (int k = 0; k < KC; ++k) acc[k] = 0.f;
    float xx = 0.f;
    const int D4 = D >> 2;
    for (int c = 0; c < D4; ++c) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + 4 * c);
        xx = fmaf(xv.x, xv.x, xx); xx = fmaf(xv.y, xv.y, xx);
        xx = fmaf(xv.z, xv.z, xx); xx = fmaf(xv.w, xv.w, xx);
        const float4* wrow = sW4 + c * KC;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const float4 w = wrow[k];
            acc[k] = fmaf(xv.x, w.x, acc[k]); acc[k] = fmaf(xv.y, w.y, acc[k]);
            acc[k] = fmaf(xv.z, w.z, acc[k]); acc[k] = fmaf(xv.w, w.w, acc[k]);
        }
    }
    return xx;
}

template <bool L2>
__device__ __forceinline__ float score_of(float dot, float xx, float b, float tau) {
    if (L2) {
        const float dist = __fsub_rn(__fadd_rn(xx, b), 2.f * dot);   // (|x|^2 + |e|^2) - 2 x.e   (:210-212)
        return tau * (-dist);                                        // relu(temp) * -dist        (:115,:213)
    }
    return dot + b;                                                  // F.linear                   (:190)
}

// gather + straight-through + squared error for one row; result overwrites the row's x in smem
template <bool L2>
__device__ __forceinline__ float finish_row(const FwdP& p, float* xr, int code) {
    const bool skip = (p.flags & VQB_SKIP) != 0;
    const float* crow = p.tab + (size_t)code * p.D;
    float se = 0.f;
    for (int d = 0; d < p.D; d += 4) {
        const float4 xv = *reinterpret_cast<const float4*>(xr + d);
        const float4 c = ldg4(crow + d);
        float4 q;
        if (L2) {
            // new_latent = enc_embs + picked_code - enc_embs.detach()  (:145): fl(fl(x + c) - x)
            q.x = __fsub_rn(__fadd_rn(xv.x, c.x), xv.x); q.y = __fsub_rn(__fadd_rn(xv.y, c.y), xv.y);
            q.z = __fsub_rn(__fadd_rn(xv.z, c.z), xv.z); q.w = __fsub_rn(__fadd_rn(xv.w, c.w), xv.w);
            if (skip) q = xv;                                        // (:142)
        } else {
            q = c;                                                   // (:194-197)
        }
        const float dx0 = xv.x - c.x, dx1 = xv.y - c.y, dx2 = xv.z - c.z, dx3 = xv.w - c.w;
        se = fmaf(dx0, dx0, se); se = fmaf(dx1, dx1, se); se = fmaf(dx2, dx2, se); se = fmaf(dx3, dx3, se);
        *reinterpret_cast<float4*>(xr + d) = q;
    }
    return se;
}

__device__ __forceinline__ void store_q_tile(const FwdP& p, int row0, int rows, const float* sX, int XS) {
    const int D4 = p.D >> 2;
    for (int i = threadIdx.x; i < rows * D4; i += TILE) {
        const int r = i / D4, c = i - r * D4;
        stg4_stream(p.q + (size_t)(row0 + r) * p.D + 4 * c, *reinterpret_cast<const float4*>(sX + r * XS + 4 * c));
    }
}

__device__ __forceinline__ void block_add_double(double* dst, float v, float* sRed) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sRed[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < TILE / 32; ++w) s += (double)sRed[w];
        atomicAdd(dst, s);
    }
}

// ------------------------------------------------------------------------------------------------
// K <= KC: every score of a row is held in registers
// ------------------------------------------------------------------------------------------------
template <int KC, bool L2>
__global__ void __launch_bounds__(TILE)
vqb_fwd_simt_small_kernel(FwdP p) {
    extern __shared__ __align__(16) float smem[];
    const int D = p.D, K = p.K, XS = D + 4;
    float4* sW4 = reinterpret_cast<float4*>(smem);          // [D/4][KC]
    float* sB = smem + KC * D;                               // [KC]
    float* sX = sB + KC;                                     // [TILE][XS]
    float* sP = sX + TILE * XS;                              // [TILE][K]   (only if p.p)
    __shared__ int sHist[KC];
    __shared__ float sRed[TILE / 32];

    const int row0 = blockIdx.x * TILE;
    const int rows = min(TILE, p.N - row0);
    const int t = threadIdx.x;
    if (t < KC) sHist[t] = 0;
    load_w_chunk<KC>(p, 0, sW4, sB);
    load_x_tile(p, row0, rows, sX, XS);
    __syncthreads();

    float* xr = sX + t * XS;
    float acc[KC];
    const float xx = dot_chunk<KC>(xr, sW4, D, acc);
    const float tau = L2 ? fmaxf(__ldg(p.temp), 0.f) : 1.f;

    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        float s = score_of<L2>(acc[k], xx, sB[k], tau);
        if (k >= K) s = -INFINITY;
        acc[k] = s;
        m = fmaxf(m, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const float e = (k < K) ? expf(acc[k] - m) : 0.f;
        acc[k] = e;
        sum += e;
    }
    int best = 0;
    float bv = -1.f;
#pragma unroll
    for (int k = 0; k < KC; ++k) {
        const float pk = acc[k] / sum;                       // softmax (:127)
        acc[k] = pk;
        if (k < K && pk > bv) { bv = pk; best = k; }         // argmax over p_code, first max (:130)
    }
    if (p.p) {
#pragma unroll
        for (int k = 0; k < KC; ++k) if (k < K) sP[t * K + k] = acc[k];
    }
    float se = 0.f;
    if (t < rows) {
        se = finish_row<L2>(p, xr, best);
        p.idx[row0 + t] = best;
        if (p.hist) atomicAdd(&sHist[best], 1);
    }
    __syncthreads();
    store_q_tile(p, row0, rows, sX, XS);
    if (p.p) {
        float* dst = p.p + (size_t)row0 * K;
        const int n = rows * K, n4 = n >> 2;
        for (int i = t; i < n4; i += TILE)
            stg4_stream(dst + 4 * i, *reinterpret_cast<const float4*>(sP + 4 * i));
        for (int i = 4 * n4 + t; i < n; i += TILE) dst[i] = sP[i];
    }
    if (p.hist && t < K && sHist[t]) atomicAdd(p.hist + t, (unsigned long long)sHist[t]);
    if (p.sqerr) block_add_double(p.sqerr, se, sRed);
}

// ------------------------------------------------------------------------------------------------
// any K: the codebook streams through shared memory in chunks of KC codes; online softmax
// statistics; raw scores are parked in p_code (if requested) and normalised in a second sweep.
// ------------------------------------------------------------------------------------------------
template <int KC, bool L2>
__global__ void __launch_bounds__(TILE)
vqb_fwd_simt_generic_kernel(FwdP p) {
    extern __shared__ __align__(16) float smem[];
    const int D = p.D, K = p.K, XS = D + 4;
    float4* sW4 = reinterpret_cast<float4*>(smem);
    float* sB = smem + KC * D;
    float* sX = sB + KC;
    __shared__ float sRed[TILE / 32];

    const int row0 = blockIdx.x * TILE;
    const int rows = min(TILE, p.N - row0);
    const int t = threadIdx.x;
    const bool valid = t < rows;
    load_x_tile(p, row0, rows, sX, XS);
    float* xr = sX + t * XS;
    const float tau = L2 ? fmaxf(__ldg(p.temp), 0.f) : 1.f;
    float* prow = p.p ? p.p + (size_t)(row0 + t) * K : nullptr;

    float m = -INFINITY, sum = 0.f, best_s = -INFINITY;
    int best = 0;
    for (int k0 = 0; k0 < K; k0 += KC) {
        __syncthreads();
        load_w_chunk<KC>(p, k0, sW4, sB);
        __syncthreads();
        float acc[KC];
        const float xx = dot_chunk<KC>(xr, sW4, D, acc);
        float cm = -INFINITY;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            float s = score_of<L2>(acc[k], xx, sB[k], tau);
            if (k0 + k >= K) s = -INFINITY;
            acc[k] = s;
            cm = fmaxf(cm, s);
            if (s > best_s) { best_s = s; best = k0 + k; }
        }
        if (cm > m) { sum *= expf(m - cm); m = cm; }
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            if (k0 + k < K) {
                sum += expf(acc[k] - m);
                if (prow && valid) prow[k0 + k] = acc[k];
            }
        }
    }
    if (prow && valid) {
        float bv = -1.f;
        for (int k = 0; k < K; ++k) {
            const float pk = expf(prow[k] - m) / sum;
            prow[k] = pk;
            if (pk > bv) { bv = pk; best = k; }               // argmax over p_code (:130)
        }
    }
    float se = 0.f;
    if (valid) {
        se = finish_row<L2>(p, xr, best);
        p.idx[row0 + t] = best;
        if (p.hist) atomicAdd(p.hist + best, 1ull);
    }
    __syncthreads();
    store_q_tile(p, row0, rows, sX, XS);
    if (p.sqerr) block_add_double(p.sqerr, se, sRed);
}

template <int KC, bool L2>
static int launch_small(const FwdP& p, cudaStream_t s) {
    const size_t smem = ((size_t)KC * p.D + KC + (size_t)TILE * (p.D + 4) + (p.p ? (size_t)TILE * p.K : 0)) * 4;
    if ((int)smem > max_optin_smem()) return invalid("vqb_forward: D=%d needs %zu B of shared memory", p.D, smem);
    auto kern = vqb_fwd_simt_small_kernel<KC, L2>;
    VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel_event_begin(s);
    kern<<<(unsigned)ceil_div(p.N, TILE), TILE, smem, s>>>(p);
    kernel_event_end(s);
    VQB_CHECK_LAUNCH("vqb_fwd_simt_small_kernel");
    return VQB_OK;
}

template <bool L2>
static int launch_generic(const FwdP& p, cudaStream_t s) {
    constexpr int KC = 32;
    const size_t smem = ((size_t)KC * p.D + KC + (size_t)TILE * (p.D + 4)) * 4;
    if ((int)smem > max_optin_smem()) return invalid("vqb_forward: D=%d needs %zu B of shared memory", p.D, smem);
    auto kern = vqb_fwd_simt_generic_kernel<KC, L2>;
    VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel_event_begin(s);
    kern<<<(unsigned)ceil_div(p.N, TILE), TILE, smem, s>>>(p);
    kernel_event_end(s);
    VQB_CHECK_LAUNCH("vqb_fwd_simt_generic_kernel");
    return VQB_OK;
}

template <bool L2>
static int dispatch_fwd(const FwdP& p, cudaStream_t s) {
    if (p.K <= 16) return launch_small<16, L2>(p, s);
    if (p.K <= 32) return launch_small<32, L2>(p, s);
    if (p.K <= 48) return launch_small<48, L2>(p, s);
    if (p.K <= 64) return launch_small<64, L2>(p, s);
    return launch_generic<L2>(p, s);
}

int launch_forward_simt(const vqb_fwd_args* a, cudaStream_t s) {
    FwdP p;
    p.x = a->x; p.w = a->score_w; p.b = a->score_b; p.tab = a->gather_table; p.temp = a->temp;
    p.p = a->p_code; p.idx = (long long*)a->idx; p.q = a->new_latent;
    p.hist = (unsigned long long*)a->hist; p.sqerr = a->sq_err_sum;
    p.N = (int)a->n_rows; p.D = (int)a->dim; p.K = (int)a->n_codes; p.flags = a->flags;
    if (p.N == 0) return VQB_OK;
    if (a->flags & VQB_SCORE_L2) return dispatch_fwd<true>(p, s);
    return dispatch_fwd<false>(p, s);
}

}  // namespace vqb
