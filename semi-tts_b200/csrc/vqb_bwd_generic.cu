// Backward of the quantizer's p_code route for ANY codebook size (K > 64, or D outside the register-tiled kernels):
// exact fp32 on CUDA cores, three tiled kernels around a coefficient matrix in the caller's workspace.
//
// Autograd of src/embed.py:105-147 / :187-205 with stop_grad (entered from src/solver.py:144); algebra as in
// vqb_bwd_simt.cu / DESIGN.md:
//   Gs = P * (g_p - rowsum(g_p * P));   C = -tau Gs (L2)  |  Gs (LINEAR)
//   dx  = g_q + 2 x rowsum(C) - 2 C @ E          (L2)     |  C @ W           (LINEAR)
//   dE += -2 C*^T @ x + scatter_add(idx, g_q)     (L2)     |  dW += C*^T @ x ; dT += scatter_add(idx, g_q)
//   colsum += colsum(C*)                          (C* = rows below n_real_rows, first_n_real_mel)
// Without stop_grad (ST-onehot, src/embed.py:137-138 / :199-203) d p_hard = g_q @ T^T joins the softmax route: a third
// tiled contraction writes it into C before the coefficient kernel runs (G = g_p + g_q @ T^T).  A learnable temperature
// (temp < 0 in the config, src/embed.py:70-74) gets  d temp = sum Gs * (-dist) = (1/tau) sum Gs * log P  -- the two are
// equal because -dist_k = (log P_k + logsumexp) / tau and sum_k Gs_k = 0 -- so no distance is recomputed here.
// The small-codebook kernels (vqb_bwd_pc.cu, vqb_bwd_simt.cu) keep a row's K coefficients in registers
// and the whole codebook in shared memory; here C[N,K] goes through HBM once (as the reference's own autograd graph
// does) and the K x D contractions are shared-memory tiled.
#include "vqb_common.cuh"

namespace vqb {

constexpr int GT = 32;          // tile edge: 32 rows x 32 codes
constexpr int GMAXV = 4;        // float4 per thread along D: D <= 8 * 4 * GMAXV = 128 per column pass

// ---- G0[row,k] = sum_d g_q[row,d] T[k,d]   (ST-onehot only)      block = 32 rows x 32 codes, thread = (row, 4 codes) ----
__global__ void __launch_bounds__(256)
bwdg_gqt_kernel(const float* __restrict__ gq, const float* __restrict__ T, long long N, int K, int D, float* __restrict__ G0) {
    __shared__ float sG[GT][GT + 1];                                // [row][d]
    __shared__ float sT[GT][GT + 1];                                // [code][d]
    const int t = threadIdx.x, r = t >> 3, c8 = t & 7;
    const long long row0 = (long long)blockIdx.x * GT;
    const int k0 = blockIdx.y * GT;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int d0 = 0; d0 < D; d0 += GT) {
        __syncthreads();
        for (int i = t; i < GT * GT; i += 256) {
            const int a = i >> 5, d = i & 31;
            const bool din = d0 + d < D;
            sG[a][d] = (din && row0 + a < N) ? __ldg(gq + (size_t)(row0 + a) * D + d0 + d) : 0.f;
            sT[a][d] = (din && k0 + a < K) ? __ldg(T + (size_t)(k0 + a) * D + d0 + d) : 0.f;
        }
        __syncthreads();
#pragma unroll 8
        for (int d = 0; d < GT; ++d) {
            const float g = sG[r][d];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j] = fmaf(g, sT[c8 + 8 * j][d], acc[j]);
        }
    }
    const long long row = row0 + r;
    if (row < N) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int k = k0 + c8 + 8 * j;
            if (k < K) G0[(size_t)row * K + k] = acc[j];
        }
    }
}

// ---- C[row,:] = cmul * P * (G - sum_k G P),  G = g_p (+ what bwdg_gqt_kernel left in C),  rowsum[row] = sum_k C[row,k] ----
// one warp per row; d temp (L2, learnable temperature) = (1/tau) sum_k Gs log P, one atomic per block of 8 rows
__global__ void __launch_bounds__(256)
bwdg_coef_kernel(const float* __restrict__ p, const float* __restrict__ gp, int add_g0, long long N, int K,
                 const float* __restrict__ temp, int l2, float* C, float* __restrict__ rowsum, float* __restrict__ dtemp) {
    __shared__ float sDt[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const bool live = row < N;
    const float tval = l2 ? __ldg(temp) : 1.f;
    const float tau = fmaxf(tval, 0.f);
    const float cmul = l2 ? -tau : 1.f;
    float dt = 0.f;
    if (live) {
        const float* pr = p + (size_t)row * K;
        const float* gr = gp ? gp + (size_t)row * K : nullptr;
        float* cr = C + (size_t)row * K;
        float s = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float g = (gr ? __ldg(gr + k) : 0.f) + (add_g0 ? cr[k] : 0.f);
            s = fmaf(g, __ldg(pr + k), s);
        }
        s = warp_sum(s);
        float rs = 0.f;
        for (int k = lane; k < K; k += 32) {
            const float g = (gr ? __ldg(gr + k) : 0.f) + (add_g0 ? cr[k] : 0.f);
            const float pk = __ldg(pr + k);
            const float gs = pk * (g - s);                          // softmax backward (:127)
            const float c = cmul * gs;
            cr[k] = c;
            rs += c;
            if (dtemp && pk > 0.f) dt = fmaf(gs, logf(pk), dt);     // P == 0: the term's limit is 0
        }
        rs = warp_sum(rs);
        if (lane == 0) rowsum[row] = rs;
    }
    if (dtemp) {                                                    // uniform over the grid
        dt = warp_sum(dt);
        if (lane == 0) sDt[w] = dt;
        __syncthreads();
        if (threadIdx.x == 0 && tval > 0.f) {                       // relu'(temp) = [temp > 0]
            float v = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) v += sDt[i];
            if (v != 0.f) atomicAdd(dtemp, v / tau);
        }
    }
}

// ---- dx[row,:] = base + ad * sum_k C[row,k] E[k,:]        block = 32 rows, thread = (row, 1/8 of the columns) -----------
// Columns are processed in passes of 128 (col0); a thread owns float4 columns c8 + 8 j (j < GMAXV) of the pass.
__global__ void __launch_bounds__(256)
bwdg_dx_kernel(const float* __restrict__ C, const float* __restrict__ rowsum, const float* __restrict__ E,
               const float* __restrict__ x, const float* __restrict__ gq, long long N, int K, int D, int l2,
               float* __restrict__ dx) {
    __shared__ float sC[GT][GT + 1];
    __shared__ __align__(16) float sE[GT][128];
    const int t = threadIdx.x, r = t >> 3, c8 = t & 7;
    const long long row0 = (long long)blockIdx.x * GT;
    const long long row = row0 + r;
    const bool live = row < N;
    const float ad = l2 ? -2.f : 1.f;
    for (int col0 = 0; col0 < D; col0 += 128) {
        const int w4 = min(128, D - col0) >> 2;                     // float4 columns of this pass
        float4 acc[GMAXV];
#pragma unroll
        for (int j = 0; j < GMAXV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k0 = 0; k0 < K; k0 += GT) {
            __syncthreads();
            for (int i = t; i < GT * GT; i += 256) {
                const int rr = i >> 5, kk = i & 31;
                sC[rr][kk] = (row0 + rr < N && k0 + kk < K) ? C[(size_t)(row0 + rr) * K + k0 + kk] : 0.f;
            }
            for (int i = t; i < GT * w4; i += 256) {
                const int kk = i / w4, c = i - kk * w4;
                *reinterpret_cast<float4*>(&sE[kk][4 * c]) =
                    (k0 + kk < K) ? ldg4(E + (size_t)(k0 + kk) * D + col0 + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
#pragma unroll 8
            for (int kk = 0; kk < GT; ++kk) {
                const float cv = sC[r][kk];
#pragma unroll
                for (int j = 0; j < GMAXV; ++j) {
                    const int c = c8 + 8 * j;
                    if (c < w4) {
                        const float4 e = *reinterpret_cast<const float4*>(&sE[kk][4 * c]);
                        acc[j].x = fmaf(cv, e.x, acc[j].x); acc[j].y = fmaf(cv, e.y, acc[j].y);
                        acc[j].z = fmaf(cv, e.z, acc[j].z); acc[j].w = fmaf(cv, e.w, acc[j].w);
                    }
                }
            }
        }
        if (live) {
            const float r2 = l2 ? 2.f * rowsum[row] : 0.f;
#pragma unroll
            for (int j = 0; j < GMAXV; ++j) {
                const int c = c8 + 8 * j;
                if (c < w4) {
                    const size_t o = (size_t)row * D + col0 + 4 * c;
                    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (l2) {
                        const float4 xv = ldg4(x + o);
                        b = make_float4(xv.x * r2, xv.y * r2, xv.z * r2, xv.w * r2);
                        if (gq) { const float4 g = ldg4(gq + o); b.x += g.x; b.y += g.y; b.z += g.z; b.w += g.w; }
                    }
                    *reinterpret_cast<float4*>(dx + o) = make_float4(fmaf(ad, acc[j].x, b.x), fmaf(ad, acc[j].y, b.y),
                                                                     fmaf(ad, acc[j].z, b.z), fmaf(ad, acc[j].w, b.w));
                }
            }
        }
    }
}

// ---- dE[k,:] += ad * sum_{row < n_eff} C[row,k] x[row,:],  colsum[k] += sum_{row < n_eff} C[row,k] ------------------------
// block = 32 codes x one chunk of rows (blockIdx.y), thread = (code, 1/8 of the columns); the chunk's partial sums go to
// dE / colsum by global reductions (the order of the chunks is not fixed: this generic route is not bit-reproducible).
__global__ void __launch_bounds__(256)
bwdg_de_kernel(const float* __restrict__ C, const float* __restrict__ x, long long n_eff, int K, int D, int l2,
               int rows_per_chunk, float* __restrict__ dE, float* __restrict__ colsum) {
    __shared__ float sC[GT][GT + 1];                                // [row][code]
    __shared__ __align__(16) float sX[GT][128];
    const int t = threadIdx.x, kq = t >> 3, c8 = t & 7;
    const int k0 = blockIdx.x * GT;
    const long long beg = (long long)blockIdx.y * rows_per_chunk;
    const long long end = min(n_eff, beg + (long long)rows_per_chunk);
    const float ad = l2 ? -2.f : 1.f;
    for (int col0 = 0; col0 < D; col0 += 128) {
        const int w4 = min(128, D - col0) >> 2;
        float4 acc[GMAXV];
#pragma unroll
        for (int j = 0; j < GMAXV; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        float cs = 0.f;
        for (long long r0 = beg; r0 < end; r0 += GT) {
            __syncthreads();
            for (int i = t; i < GT * GT; i += 256) {
                const int rr = i >> 5, kk = i & 31;
                sC[rr][kk] = (r0 + rr < end && k0 + kk < K) ? C[(size_t)(r0 + rr) * K + k0 + kk] : 0.f;
            }
            for (int i = t; i < GT * w4; i += 256) {
                const int rr = i / w4, c = i - rr * w4;
                *reinterpret_cast<float4*>(&sX[rr][4 * c]) =
                    (r0 + rr < end) ? ldg4(x + (size_t)(r0 + rr) * D + col0 + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            __syncthreads();
#pragma unroll 8
            for (int rr = 0; rr < GT; ++rr) {
                const float cv = sC[rr][kq];
                cs += cv;
#pragma unroll
                for (int j = 0; j < GMAXV; ++j) {
                    const int c = c8 + 8 * j;
                    if (c < w4) {
                        const float4 xv = *reinterpret_cast<const float4*>(&sX[rr][4 * c]);
                        acc[j].x = fmaf(cv, xv.x, acc[j].x); acc[j].y = fmaf(cv, xv.y, acc[j].y);
                        acc[j].z = fmaf(cv, xv.z, acc[j].z); acc[j].w = fmaf(cv, xv.w, acc[j].w);
                    }
                }
            }
        }
        const int k = k0 + kq;
        if (k < K) {
#pragma unroll
            for (int j = 0; j < GMAXV; ++j) {
                const int c = c8 + 8 * j;
                if (c < w4)
                    red_add_v4(dE + (size_t)k * D + col0 + 4 * c, make_float4(ad * acc[j].x, ad * acc[j].y, ad * acc[j].z, ad * acc[j].w));
            }
            if (col0 == 0 && c8 == 0) atomicAdd(colsum + k, cs);
        }
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------
static size_t align256g(size_t v) { return (v + 255) & ~(size_t)255; }

bool backward_generic_needed(const vqb_bwd_args* a) {
    return a->n_codes > 64 || a->dim % 8 != 0 || a->dim > 128;      // what the register-tiled kernels cannot take
}

size_t backward_generic_workspace(const vqb_bwd_args* a) {
    const size_t N = (size_t)a->n_rows, K = (size_t)a->n_codes;
    return align256g(N * K * 4) + align256g(N * 4) + scatter_workspace_bytes(a->n_rows, a->n_codes, a->dim);
}

int launch_backward_generic(const vqb_bwd_args* a, cudaStream_t s) {
    const int64_t N = a->n_rows, K = a->n_codes, D = a->dim;
    const bool l2 = (a->flags & VQB_SCORE_L2) != 0;
    const size_t need = backward_generic_workspace(a);
    if (!a->workspace || a->workspace_bytes < need) {
        set_error("vqb_backward: workspace too small (%zu < %zu bytes)", a->workspace_bytes, need);
        return VQB_ERR_WORKSPACE;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
    float* C = reinterpret_cast<float*>(ws);
    float* rowsum = reinterpret_cast<float*>(ws + align256g((size_t)N * K * 4));
    uint8_t* sws = ws + align256g((size_t)N * K * 4) + align256g((size_t)N * 4);
    const size_t sws_bytes = scatter_workspace_bytes(N, K, D);

    // ST-onehot: d p_hard = g_q @ T^T joins the softmax route (not in the L2 skip branch, where the gather is unused)
    const bool add_g0 = !(a->flags & VQB_STOP_GRAD) && a->g_q && !(l2 && (a->flags & VQB_SKIP));
    float* dtemp = (l2 && (a->flags & VQB_TEMP_GRAD)) ? a->d_temp : nullptr;
    const int64_t kblocks = ceil_div(K, GT);
    if (kblocks > 65535) return invalid("vqb_backward: the generic p_code-route backward supports K <= %d (got %lld)", 65535 * GT, (long long)K);
    kernel_event_begin(s);
    if (add_g0) {
        bwdg_gqt_kernel<<<dim3((unsigned)ceil_div(N, GT), (unsigned)kblocks), 256, 0, s>>>(a->g_q, a->gather_table, N, (int)K, (int)D, C);
        VQB_CHECK_LAUNCH("bwdg_gqt_kernel");
    }
    bwdg_coef_kernel<<<(unsigned)ceil_div(N, 8), 256, 0, s>>>(a->p_code, a->g_p, add_g0 ? 1 : 0, N, (int)K, a->temp, l2 ? 1 : 0, C, rowsum, dtemp);
    VQB_CHECK_LAUNCH("bwdg_coef_kernel");
    bwdg_dx_kernel<<<(unsigned)ceil_div(N, GT), 256, 0, s>>>(C, rowsum, a->score_w, a->x, a->g_q, N, (int)K, (int)D, l2 ? 1 : 0, a->dx);
    VQB_CHECK_LAUNCH("bwdg_dx_kernel");
    const int64_t n_eff = (a->n_real_rows > 0 && a->n_real_rows < N) ? a->n_real_rows : N;
    // enough row chunks to fill the machine a few times over, at least 256 rows each
    int64_t chunks = ceil_div((int64_t)sm_count() * 4, kblocks);
    int64_t rpc = ceil_div(n_eff, chunks < 1 ? 1 : chunks);
    rpc = ceil_div(rpc < 256 ? 256 : rpc, GT) * GT;
    chunks = ceil_div(n_eff, rpc);
    if (chunks > 65535) { rpc = ceil_div(ceil_div(n_eff, 65535), GT) * GT; chunks = ceil_div(n_eff, rpc); }
    bwdg_de_kernel<<<dim3((unsigned)kblocks, (unsigned)chunks), 256, 0, s>>>(C, a->x, n_eff, (int)K, (int)D, l2 ? 1 : 0, (int)rpc,
                                                                           a->d_score_w, a->colsum);
    VQB_CHECK_LAUNCH("bwdg_de_kernel");
    kernel_event_end(s);
    if (a->g_q && !(l2 && (a->flags & VQB_SKIP))) {
        float* dst = l2 ? a->d_score_w : a->d_gather;
        if (!dst) return invalid("vqb_backward: the scatter destination (d_score_w for L2, d_gather for LINEAR) is NULL");
        return launch_scatter_add(a->idx, N, a->g_q, K, D, dst, nullptr, sws_bytes ? sws : nullptr, sws_bytes, s);
    }
    return VQB_OK;
}

}  // namespace vqb
