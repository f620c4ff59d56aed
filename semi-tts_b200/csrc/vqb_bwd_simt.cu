// Exact-fp32 backward of the quantizer on CUDA cores (both gradient routes).
//
// The reference has no source for this: it is autograd of src/embed.py:105-147 / :187-205, entered
// from src/solver.py:144.  Algebra (restated in oracle/vq_oracle.py):
//   G  = g_p (+ g_q @ T^T without stop_grad);   Gs = P * (G - rowsum(G * P))
//   L2:     Gd = -tau Gs;  dx = g_q + 2 x rowsum(Gd) - 2 Gd @ E;
//           dE += -2 Gd*^T @ x + scatter_add(idx, g_q);  colsum += colsum(Gd*)   (Gd*: rows < n_real)
//           dtemp += sum Gs * (-dist)
//   LINEAR: dx = Gs @ W;  dW += Gs^T @ x;  colsum += colsum(Gs);  dT += scatter_add(idx, g_q)
//
// One persistent kernel.  Per tile of BT=64 rows: (1a) thread-per-row softmax backward with the
// row's K coefficients in registers; (2) the K x D reductions over rows as a register-tiled
// outer-product accumulation (thread = 1/8 of the codes x 8 columns) that persists across the
// CTA's tiles, with the index-keyed scatter accumulated by the owning thread in shared memory;
// (1b) dx, staged through shared memory for coalesced stores.  Partial K x D sums are flushed
// once per CTA with 128-bit vector reductions (red.global.add.v4.f32).
#include <math.h>
#include "vqb_common.cuh"

namespace vqb {

constexpr int BT = 64;

struct BwdP {
    const float *x, *w, *b, *tab, *temp, *p, *gp, *gq;
    const long long* idx;
    float *dx, *dW, *colsum, *dG, *dtemp;
    int N, D, K, n_real, ntiles;
    unsigned flags;
};

template <int KC>
__device__ __forceinline__ void load_codes(const float* src, int K, int D, float4* dst4) {
    const int D4 = D >> 2;
    for (int i = threadIdx.x; i < KC * D4; i += BT) {
        const int k = i / D4, c = i - k * D4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < K) v = ldg4(src + (size_t)k * D + 4 * c);
        dst4[c * KC + k] = v;
    }
}

__device__ __forceinline__ void load_rows(const float* src, int row0, int rows, int D, float* dst, int XS) {
    const int D4 = D >> 2;
    for (int i = threadIdx.x; i < BT * D4; i += BT) {
        const int r = i / D4, c = i - r * D4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src && r < rows) v = ldg4_stream(src + (size_t)(row0 + r) * D + 4 * c);
        *reinterpret_cast<float4*>(dst + r * XS + 4 * c) = v;
    }
}

// [rows, K] block of a row-major [N, K] matrix -> dst[r * KS + k] (zero filled)
__device__ __forceinline__ void load_nk(const float* src, int row0, int rows, int K, float* dst, int KS, int KC) {
    for (int i = threadIdx.x; i < BT * KC; i += BT) {
        const int r = i / KC, k = i - r * KC;
        dst[r * KS + k] = 0.f;
    }
    __syncthreads();
    if (!src) return;
    const float* base = src + (size_t)row0 * K;
    const int n = rows * K;
    for (int i = threadIdx.x; i < n; i += BT) {
        const int r = i / K, k = i - r * K;
        dst[r * KS + k] = __ldcs(base + i);
    }
}

template <int KC, bool L2, int NDB>
__global__ void __launch_bounds__(BT)
vqb_bwd_simt_kernel(BwdP p) {
    constexpr int KPT = KC / 8;                 // codes per thread in the reduction phase
    constexpr int KS = KC + 1;                  // odd row stride: thread-per-row access is conflict free
    extern __shared__ __align__(16) float smem[];
    const int D = p.D, K = p.K, XS = D + 4;
    const bool stop_grad = (p.flags & VQB_STOP_GRAD) != 0;
    const bool skip = (p.flags & VQB_SKIP) != 0;
    const bool want_temp = L2 && (p.flags & VQB_TEMP_GRAD) != 0;
    const bool sep_tab = !stop_grad && p.tab != p.w;

    float4* sW4 = reinterpret_cast<float4*>(smem);              // [D/4][KC]
    float4* sT4 = sep_tab ? sW4 + (KC * D) / 4 : sW4;           // [D/4][KC] gather table if distinct
    float* sAcc = reinterpret_cast<float*>(sT4 + (KC * D) / 4); // [KC][D] index-keyed scatter sums
    float* sB = sAcc + KC * D;                                  // [KC]
    float* sX = sB + KC;                                        // [BT][XS]
    float* sG = sX + BT * XS;                                   // [BT][XS]  g_q, later dx
    float* sP = sG + BT * XS;                                   // [BT][KS]
    float* sC = sP + BT * KS;                                   // [BT][KS]  g_p, later coefficients
    __shared__ int sIdx[BT];
    __shared__ float sRed[BT / 32];

    const int t = threadIdx.x;
    const int kg = t >> 3, dg = t & 7;
    load_codes<KC>(p.w, K, D, sW4);
    if (sep_tab) load_codes<KC>(p.tab, K, D, sT4);
    for (int i = t; i < KC * D; i += BT) sAcc[i] = 0.f;
    for (int k = t; k < KC; k += BT) sB[k] = (want_temp && k < K) ? __ldg(p.b + k) : 0.f;
    const float temp = L2 ? __ldg(p.temp) : 1.f;
    const float tau = L2 ? fmaxf(temp, 0.f) : 1.f;

    float acc[NDB][KPT][8];
    float cs[KPT];
#pragma unroll
    for (int a = 0; a < NDB; ++a)
#pragma unroll
        for (int i = 0; i < KPT; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[a][i][j] = 0.f;
#pragma unroll
    for (int i = 0; i < KPT; ++i) cs[i] = 0.f;
    float dtemp_acc = 0.f;

    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int row0 = tile * BT;
        const int rows = min(BT, p.N - row0);
        __syncthreads();                                        // previous tile fully consumed
        load_rows(p.x, row0, rows, D, sX, XS);
        load_rows(p.gq, row0, rows, D, sG, XS);
        load_nk(p.p, row0, rows, K, sP, KS, KC);
        load_nk(p.gp, row0, rows, K, sC, KS, KC);
        if (t < BT) sIdx[t] = (t < rows) ? (int)p.idx[row0 + t] : -1;
        __syncthreads();

        // ---- phase 1a: softmax backward for row t ------------------------------------------
        const bool valid = t < rows;
        const bool real = valid && (p.n_real <= 0 || row0 + t < p.n_real);
        float G[KC];
        const float* xr = sX + t * XS;
        float* gr = sG + t * XS;
        if (!stop_grad) {
            // ST-onehot: d p_hard = g_q @ T^T joins the softmax route (:137-138 / :199-203)
#pragma unroll
            for (int k = 0; k < KC; ++k) G[k] = 0.f;
            for (int c = 0; c < (D >> 2); ++c) {
                const float4 gv = *reinterpret_cast<const float4*>(gr + 4 * c);
                const float4* trow = sT4 + c * KC;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const float4 w = trow[k];
                    G[k] = fmaf(gv.x, w.x, G[k]); G[k] = fmaf(gv.y, w.y, G[k]);
                    G[k] = fmaf(gv.z, w.z, G[k]); G[k] = fmaf(gv.w, w.w, G[k]);
                }
            }
            if (skip) {
#pragma unroll
                for (int k = 0; k < KC; ++k) G[k] = 0.f;       // skip branch: the gather is unused (:142)
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) G[k] += sC[t * KS + k];
        } else {
#pragma unroll
            for (int k = 0; k < KC; ++k) G[k] = sC[t * KS + k];
        }
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < KC; ++k) s = fmaf(G[k], sP[t * KS + k], s);
        float r = 0.f;
#pragma unroll
        for (int k = 0; k < KC; ++k) {
            const float gs = sP[t * KS + k] * (G[k] - s);       // softmax backward (:127)
            float coef;
            if (L2) { G[k] = -tau * gs; coef = real ? -2.f * G[k] : 0.f; }
            else    { G[k] = gs;        coef = valid ? gs : 0.f; }
            r += G[k];
            sC[t * KS + k] = coef;
        }
        float gb = 0.f;                                         // sum_k Gd[k] * |e_k|^2 (for d temp)
        if (want_temp) {
#pragma unroll
            for (int k = 0; k < KC; ++k) gb = fmaf(G[k], sB[k], gb);
        }
        __syncthreads();

        // ---- phase 2: K x D reductions over the tile's rows -----------------------------------
#pragma unroll
        for (int a = 0; a < NDB; ++a) {
            const int d0 = a * 64 + dg * 8;
            if (d0 < D) {
                for (int n = 0; n < BT; ++n) {
                    float c[KPT];
#pragma unroll
                    for (int i = 0; i < KPT; ++i) c[i] = sC[n * KS + kg * KPT + i];
                    const float4 x0 = *reinterpret_cast<const float4*>(sX + n * XS + d0);
                    const float4 x1 = *reinterpret_cast<const float4*>(sX + n * XS + d0 + 4);
#pragma unroll
                    for (int i = 0; i < KPT; ++i) {
                        acc[a][i][0] = fmaf(c[i], x0.x, acc[a][i][0]); acc[a][i][1] = fmaf(c[i], x0.y, acc[a][i][1]);
                        acc[a][i][2] = fmaf(c[i], x0.z, acc[a][i][2]); acc[a][i][3] = fmaf(c[i], x0.w, acc[a][i][3]);
                        acc[a][i][4] = fmaf(c[i], x1.x, acc[a][i][4]); acc[a][i][5] = fmaf(c[i], x1.y, acc[a][i][5]);
                        acc[a][i][6] = fmaf(c[i], x1.z, acc[a][i][6]); acc[a][i][7] = fmaf(c[i], x1.w, acc[a][i][7]);
                        if (a == 0 && dg == 0) cs[i] += c[i];
                    }
                }
                if (p.gq && !skip) {
                    // index-keyed scatter: the thread that owns (code, d0..d0+7) adds g_q[n, d0..d0+7]
                    for (int n = 0; n < rows; ++n) {
                        const int code = sIdx[n];
                        const unsigned rel = (unsigned)(code - kg * KPT);
                        if (rel < (unsigned)KPT) {
                            float4* dst = reinterpret_cast<float4*>(sAcc + code * D + d0);
                            const float4 g0 = *reinterpret_cast<const float4*>(sG + n * XS + d0);
                            const float4 g1 = *reinterpret_cast<const float4*>(sG + n * XS + d0 + 4);
                            float4 a0 = dst[0], a1 = dst[1];
                            a0.x += g0.x; a0.y += g0.y; a0.z += g0.z; a0.w += g0.w;
                            a1.x += g1.x; a1.y += g1.y; a1.z += g1.z; a1.w += g1.w;
                            dst[0] = a0; dst[1] = a1;
                        }
                    }
                }
            }
        }
        __syncthreads();

        // ---- phase 1b: dx for row t, staged over the g_q tile ----------------------------------
        if (p.dx) {
            float xx = 0.f, xdot = 0.f;
            for (int c = 0; c < (D >> 2); ++c) {
                float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
                const float4* wrow = sW4 + c * KC;
#pragma unroll
                for (int k = 0; k < KC; ++k) {
                    const float4 w = wrow[k];
                    a4.x = fmaf(G[k], w.x, a4.x); a4.y = fmaf(G[k], w.y, a4.y);
                    a4.z = fmaf(G[k], w.z, a4.z); a4.w = fmaf(G[k], w.w, a4.w);
                }
                float4 o;
                if (L2) {
                    const float4 xv = *reinterpret_cast<const float4*>(xr + 4 * c);
                    const float4 gv = *reinterpret_cast<const float4*>(gr + 4 * c);
                    const float r2 = 2.f * r;
                    o.x = fmaf(xv.x, r2, gv.x) - 2.f * a4.x; o.y = fmaf(xv.y, r2, gv.y) - 2.f * a4.y;
                    o.z = fmaf(xv.z, r2, gv.z) - 2.f * a4.z; o.w = fmaf(xv.w, r2, gv.w) - 2.f * a4.w;
                    xx = fmaf(xv.x, xv.x, xx); xx = fmaf(xv.y, xv.y, xx); xx = fmaf(xv.z, xv.z, xx); xx = fmaf(xv.w, xv.w, xx);
                    xdot = fmaf(xv.x, a4.x, xdot); xdot = fmaf(xv.y, a4.y, xdot);
                    xdot = fmaf(xv.z, a4.z, xdot); xdot = fmaf(xv.w, a4.w, xdot);
                } else {
                    o = a4;
                }
                *reinterpret_cast<float4*>(gr + 4 * c) = o;
            }
            // d temp = sum_k Gs (-dist), Gs = Gd / (-tau), dist_k = |x|^2 + |e_k|^2 - 2 x.e_k
            //        = (|x|^2 sum_k Gd + sum_k Gd |e_k|^2 - 2 x . (Gd @ E)) / tau      (relu'(temp) = [temp > 0])
            if (want_temp && valid && temp > 0.f) dtemp_acc += (fmaf(xx, r, gb) - 2.f * xdot) / tau;
            __syncthreads();
            const int D4 = D >> 2;
            for (int i = t; i < rows * D4; i += BT) {
                const int rr = i / D4, c = i - rr * D4;
                stg4_stream(p.dx + (size_t)(row0 + rr) * D + 4 * c,
                            *reinterpret_cast<const float4*>(sG + rr * XS + 4 * c));
            }
        }
    }

    // ---- flush the CTA's partial sums ---------------------------------------------------------
    __syncthreads();
#pragma unroll
    for (int a = 0; a < NDB; ++a) {
        const int d0 = a * 64 + dg * 8;
        if (d0 < D) {
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const int k = kg * KPT + i;
                if (k < K) {
                    float4 v0 = make_float4(acc[a][i][0], acc[a][i][1], acc[a][i][2], acc[a][i][3]);
                    float4 v1 = make_float4(acc[a][i][4], acc[a][i][5], acc[a][i][6], acc[a][i][7]);
                    const float4 s0 = *reinterpret_cast<const float4*>(sAcc + k * D + d0);
                    const float4 s1 = *reinterpret_cast<const float4*>(sAcc + k * D + d0 + 4);
                    if (L2) {
                        v0.x += s0.x; v0.y += s0.y; v0.z += s0.z; v0.w += s0.w;
                        v1.x += s1.x; v1.y += s1.y; v1.z += s1.z; v1.w += s1.w;
                    } else if (p.dG) {
                        red_add_v4(p.dG + (size_t)k * D + d0, s0);
                        red_add_v4(p.dG + (size_t)k * D + d0 + 4, s1);
                    }
                    red_add_v4(p.dW + (size_t)k * D + d0, v0);
                    red_add_v4(p.dW + (size_t)k * D + d0 + 4, v1);
                }
            }
        }
    }
    if (dg == 0 && p.colsum) {
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const int k = kg * KPT + i;
            if (k < K) atomicAdd(p.colsum + k, L2 ? -0.5f * cs[i] : cs[i]);
        }
    }
    if (want_temp) {
        float v = warp_sum(dtemp_acc);
        if ((t & 31) == 0) sRed[t >> 5] = v;
        __syncthreads();
        if (t == 0) atomicAdd(p.dtemp, sRed[0] + sRed[1]);
    }
}

template <int KC, bool L2, int NDB>
static int launch_bwd(BwdP& p, cudaStream_t s) {
    const bool sep_tab = !(p.flags & VQB_STOP_GRAD) && p.tab != p.w;
    const size_t fl = (size_t)KC * p.D * (sep_tab ? 3 : 2) + KC + 2 * (size_t)BT * (p.D + 4) + 2 * (size_t)BT * (KC + 1);
    const size_t smem = fl * 4;
    if ((int)smem > max_optin_smem()) return invalid("vqb_backward: D=%d K=%d needs %zu B of shared memory", p.D, p.K, smem);
    auto kern = vqb_bwd_simt_kernel<KC, L2, NDB>;
    VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 1;
    VQB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, smem));
    if (per_sm < 1) per_sm = 1;
    const int grid = (int)min((int64_t)p.ntiles, (int64_t)sm_count() * per_sm);
    kernel_event_begin(s);
    kern<<<grid, BT, smem, s>>>(p);
    kernel_event_end(s);
    VQB_CHECK_LAUNCH("vqb_bwd_simt_kernel");
    return VQB_OK;
}

template <bool L2>
static int dispatch_bwd(BwdP& p, cudaStream_t s) {
    const int ndb = (p.D + 63) / 64;
#define VQB_BWD_CASE(KC_)                                             \
    if (p.K <= KC_) {                                                 \
        if (ndb == 1) return launch_bwd<KC_, L2, 1>(p, s);            \
        return launch_bwd<KC_, L2, 2>(p, s);                          \
    }
    VQB_BWD_CASE(16) VQB_BWD_CASE(32) VQB_BWD_CASE(48) VQB_BWD_CASE(64)
#undef VQB_BWD_CASE
    return invalid("vqb_backward: the p_code-route backward supports K <= 64 (got K=%d)", p.K);
}

int launch_backward_simt(const vqb_bwd_args* a, cudaStream_t s) {
    BwdP p;
    p.x = a->x; p.w = a->score_w; p.b = a->score_b; p.tab = a->gather_table; p.temp = a->temp;
    p.p = a->p_code; p.gp = a->g_p; p.gq = a->g_q; p.idx = (const long long*)a->idx;
    p.dx = a->dx; p.dW = a->d_score_w; p.colsum = a->colsum; p.dG = a->d_gather; p.dtemp = a->d_temp;
    p.N = (int)a->n_rows; p.D = (int)a->dim; p.K = (int)a->n_codes;
    p.n_real = (int)(a->n_real_rows > 0 ? a->n_real_rows : 0);
    p.ntiles = (int)ceil_div(a->n_rows, BT);
    p.flags = a->flags;
    if (p.N == 0) return VQB_OK;
    if (p.D % 8 != 0 || p.D > 128)
        return invalid("vqb_backward: the p_code-route backward supports D %% 8 == 0 and D <= 128 (got D=%d)", p.D);
    if (a->flags & VQB_SCORE_L2) return dispatch_bwd<true>(p, s);
    return dispatch_bwd<false>(p, s);
}

}  // namespace vqb
