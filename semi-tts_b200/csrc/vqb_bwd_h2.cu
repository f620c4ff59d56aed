// Backward of the quantizer, parity mode (K <= 64, D = 64, stop_grad), second generation: the two K x D GEMMs run
// as kind::f16 tcgen05 MMAs on operands split into two fp16 pieces ("fp16x2": v = hi + lo, 22 significant bits)
// after an exact power-of-two rescale, so one MMA burst per tile gives fp32-level accuracy.
//
// Autograd of src/embed.py:105-147 / :187-205 (entered from src/solver.py:144); algebra in DESIGN.md:
//   Gs = P * (g_p - rowsum(g_p * P));   C = -tau Gs (L2)  |  Gs (LINEAR)
//   dx  = g_q + 2 x rowsum(C) - 2 C @ E          (L2)     |  C @ W           (LINEAR)
//   dE += -2 C*^T @ x + scatter_add(idx, g_q)     (L2)     |  dW += C^T @ x ; dT += scatter_add(idx, g_q)
//   colsum += colsum(C*)                          (C* = rows below n_real_rows, first_n_real_mel)
//
// Why fp16 pieces and not tf32 (vqb_bwd_tc.cu):  a 16-bit operand tile [128 rows][64 values] is 128 bytes per row,
// and for 16-bit types the K-major and the MN-major 128-byte-swizzle layouts of that tile are the SAME bytes -- the
// coefficient tile C is written once and serves GEMM 1 (contraction over codes, K-major A) and GEMM 2 (contraction
// over rows, MN-major B).  hi and lo pieces sit in separate 16 KB tiles, so there is one MMA burst per tile instead of
// two dependent passes, the x tile is converted in place, and the shared memory saved double-buffers the x tile and
// lets the next tile's p_code / g_p blocks stream in while the current tile is still being processed.
//
// Scaling (all factors are powers of two, hence exact):
//   C'[r][:] = C[r][:] * 2^-e_r      e_r = exponent(max_k |C[r][k]|) - 14      (row maximum lands in [2^14, 2^15))
//   E'       = E * 2^-g              g   = exponent(max |E|) - 14
//   x''[r][:]= x[r][:] * 2^(e_r - t) t   = max_r(exponent(max_d |x[r][d]|) + e_r) - 14 over the tile's real rows
//   GEMM 1:  D1 = C' @ E'  ->  C @ E = D1 * 2^(e_r + g);    GEMM 2:  D2 = x''^T @ C'  ->  x^T @ C = D2 * 2^t
//   D2 is flushed into registers once per tile (the scale t differs from tile to tile).
//
// One persistent CTA per SM, 8 warps:
//   warp 0     TMA producer: x tiles (double-buffered), g_q tile, the contiguous [128 x K] blocks of p_code and g_p
//   warp 1     MMA issuer: GEMM 1 (3 MMAs per 16 codes: hi.hi + lo.hi + hi.lo) then GEMM 2 (A = [x_hi ; x_lo]
//              stacked along M, B = C_hi then C_lo: all four cross terms)
//   warps 2-3  index-keyed scatter of g_q into a shared-memory accumulator [K][64]: stable counting sort of the
//              tile's rows by code, then four half-warps walk four segments of the sorted list
//   warps 4-7  thread = row: softmax backward, rescale + split, column sums (warp butterfly, while the MMAs run),
//              dx epilogue staged in the (by then dead) x buffer and sent out by TMA
// The K x D sums leave the CTA once, as plain stores into a per-CTA partial record; reduce_partials_h2_kernel adds
// the records in a fixed order (no atomics: gradients are bit-reproducible).
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include "vqb_common.cuh"
#include "vqb_tc.cuh"

namespace vqb {
using namespace tc;

constexpr int HM = 128;                 // rows per tile
constexpr int HBLK = HM * 128;          // one [128 rows][128 B] block = 16 KB
constexpr int H_THREADS = 256;
constexpr int H_KD = 64 * 64;
constexpr int H_PARTIAL_FLOATS = 2 * H_KD + 64;   // [0] d_score_w part, [1] scatter part (LINEAR) / transposed projected
                                                  // columns (L2 with the fused tail), [2] column sums

struct BwdH2P {
    const float* p;
    const float* gp;
    const float* gq;          // may be NULL
    const long long* idx;
    const float* temp;
    const float* E;           // [K][64] score table (L2: the codebook; LINEAR: W)
    float* partial;           // [grid][H_PARTIAL_FLOATS]
    unsigned long long* dbg;  // optional timeline buffer (developer hook)
    int N, K, n_real, num_tiles;
    int t_first;              // L2 + fused tail: columns d >= t_first are also stored transposed (record plane 1), else 64
    unsigned flags;
};

#define VQB_HTL(tag) do { if (p.dbg && r == 0 && blockIdx.x == 0 && tl_n < 40) { p.dbg[tl_n++] = ((unsigned long long)(tag) << 56) | (globaltimer_ns() & 0x00FFFFFFFFFFFFFFull); } } while (0)
// marks of the scatter warps: slots [base, base + 40)
#define VQB_HTW(base, tag) do { if (p.dbg && lane == 0 && blockIdx.x == 0 && tw_n < 40) { p.dbg[(base) + tw_n++] = ((unsigned long long)(tag) << 56) | (globaltimer_ns() & 0x00FFFFFFFFFFFFFFull); } } while (0)

// floor(log2(|v|)) of a positive normal float; subnormals and zero give -127, inf/nan give 128
__device__ __forceinline__ int exp_of(float v) { return (int)((__float_as_uint(v) >> 23) & 0xFFu) - 127; }
// 2^n for n in [-126, 127] (clamped)
__device__ __forceinline__ float pow2i(int n) {
    n = n < -126 ? -126 : (n > 127 ? 127 : n);
    return __uint_as_float((uint32_t)(n + 127) << 23);
}
__device__ __forceinline__ float f16_trunc(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}
// eight scaled values -> their fp16 hi pieces (exact: 11 significant bits) and lo pieces (rounded remainder)
__device__ __forceinline__ void split8(const float* v, float s, uint4& hi, uint4& lo) {
    float h[8], l[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float w = v[i] * s;
        h[i] = f16_trunc(w);
        l[i] = w - h[i];
    }
    hi = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
    lo = make_uint4(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]), pack_h2(l[4], l[5]), pack_h2(l[6], l[7]));
}

__global__ void __launch_bounds__(H_THREADS, 1)
vqb_bwd_h2_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                  const __grid_constant__ CUtensorMap tm_dx, BwdH2P p, int stage_bytes) {
    constexpr int D = 64;
    constexpr int TILE = 2 * HBLK;                   // one [128][64] fp32 tile = two 16 KB blocks

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem;                              // [2 buffers][2 blocks][16 KB]: raw x, then x_hi | x_lo (fp16) in place
    uint8_t* sG = sX + 2 * TILE;                     // [2][16 KB] g_q tile, then dx in place
    uint8_t* sCh = sG + TILE;                        // [16 KB] C_hi  fp16 [128 rows][64 codes]
    uint8_t* sCl = sCh + HBLK;                       // [16 KB] C_lo
    uint8_t* sEh = sCl + HBLK;                       // [8 KB]  E_hi  fp16 [64 codes][64 d]
    uint8_t* sEl = sEh + 64 * 128;                   // [8 KB]  E_lo
    float* stP = reinterpret_cast<float*>(sEl + 64 * 128);                                   // staging of p_code [128 x K]
    float* stG = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stP) + stage_bytes);    // staging of g_p
    float* sAcc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(stG) + stage_bytes);   // [64][64] scatter accumulator
    int* sOrd = reinterpret_cast<int*>(sAcc + 64 * D);                                       // [2][128] rows sorted by code
    int* sCnt = sOrd + 2 * HM;                                                                   // [96]
    int* sSeg = sCnt + 96;                                                                       // [2][8] segment cuts
    int* sRed = sSeg + 16;                                                                       // [8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 8);
    uint64_t* pg_full = bars;
    uint64_t* pg_free = bars + 1;
    uint64_t* x_full = bars + 2;      // [2]
    uint64_t* x_free = bars + 4;      // [2]
    uint64_t* g_full = bars + 6;
    uint64_t* g_free = bars + 7;
    uint64_t* coef_ready = bars + 8;
    uint64_t* d1_done = bars + 9;
    uint64_t* mma_done = bars + 10;
    uint64_t* sort_full = bars + 11;  // [2]
    uint64_t* sort_free = bars + 13;  // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool l2 = (p.flags & VQB_SCORE_L2) != 0;
    const bool do_scatter = p.gq != nullptr && !(l2 && (p.flags & VQB_SKIP));
    const int K = p.K;
    const int n_my = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_g); tma_prefetch_desc(&tm_dx);
        mbar_init(pg_full, 1); mbar_init(pg_free, 4);
        mbar_init(&x_full[0], 1); mbar_init(&x_full[1], 1); mbar_init(&x_free[0], 1); mbar_init(&x_free[1], 1);
        mbar_init(g_full, 1); mbar_init(g_free, do_scatter ? 6 : 4);   // four row warps + the two scatter warps
        mbar_init(coef_ready, 4); mbar_init(d1_done, 1); mbar_init(mma_done, 1);
        mbar_init(&sort_full[0], 1); mbar_init(&sort_full[1], 1); mbar_init(&sort_free[0], 1); mbar_init(&sort_free[1], 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<128>(tmem_slot);
    for (int i = threadIdx.x; i < 64 * D; i += H_THREADS) sAcc[i] = 0.f;
    // PDL: the fused tail kernel (launched behind this one with programmatic serialization) may queue up now; it waits
    // for this grid to complete before it reads the partial records.
    pdl_launch();

    // ---- E -> E' = E * 2^-g as fp16 hi / lo tiles [64 codes][128 B], 128-byte swizzle (MN-major B of GEMM 1) --------
    float emx = 0.f;
    for (int i = threadIdx.x; i < K * D / 4; i += H_THREADS) {
        const float4 v = ldg4(p.E + 4 * i);
        emx = fmaxf(fmaxf(emx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
    }
    {
        const unsigned wm = __reduce_max_sync(0xffffffffu, __float_as_uint(emx));   // non-negative floats order like uints
        if (lane == 0) sRed[warp] = (int)wm;
    }
    __syncthreads();
    unsigned emb = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) emb = max(emb, (unsigned)sRed[w]);
    const int gE = emb ? exp_of(__uint_as_float(emb)) - 14 : 0;
    {
        const float sE = pow2i(-gE);
        for (int i = threadIdx.x; i < 64 * 8; i += H_THREADS) {
            const int k = i >> 3, j = i & 7;
            uint4 hi = make_uint4(0, 0, 0, 0), lo = hi;
            if (k < K) {
                const float4 a = ldg4(p.E + k * D + 8 * j), b = ldg4(p.E + k * D + 8 * j + 4);
                const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                split8(v, sE, hi, lo);
            }
            *reinterpret_cast<uint4*>(sEh + sw128_offset(k, j)) = hi;
            *reinterpret_cast<uint4*>(sEl + sw128_offset(k, j)) = lo;
        }
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d1 = tmem_base, d2 = tmem_base + 64;

    if (warp == 0) {
        // =============================== TMA producer =====================================================
        if (lane == 0 && n_my > 0) {
            auto tile_of = [&](int it) { return (int)blockIdx.x + it * (int)gridDim.x; };
            auto issue_pg = [&](int it) {
                const int row0 = tile_of(it) * HM;
                const int rows = min(HM, p.N - row0);
                const uint32_t bulk = (uint32_t)(rows * K * 4) & ~15u;
                mbar_arrive_expect_tx(pg_full, 2 * bulk);
                if (bulk) {
                    bulk_load_1d(stP, p.p + (size_t)row0 * K, bulk, pg_full);
                    bulk_load_1d(stG, p.gp + (size_t)row0 * K, bulk, pg_full);
                }
            };
            auto issue_x = [&](int it) {
                const int b = it & 1;
                mbar_arrive_expect_tx(&x_full[b], TILE);
                tma_load_2d(sX + b * TILE, &tm_x, 0, tile_of(it) * HM, &x_full[b]);
                tma_load_2d(sX + b * TILE + HBLK, &tm_x, 32, tile_of(it) * HM, &x_full[b]);
            };
            auto issue_g = [&](int it) {
                if (!p.gq) return;
                mbar_arrive_expect_tx(g_full, TILE);
                tma_load_2d(sG, &tm_g, 0, tile_of(it) * HM, g_full);
                tma_load_2d(sG + HBLK, &tm_g, 32, tile_of(it) * HM, g_full);
            };
            issue_pg(0); issue_x(0); issue_g(0);
            if (n_my > 1) issue_x(1);
            for (int it = 0; it < n_my; ++it) {
                if (it + 1 < n_my) {
                    mbar_wait(pg_free, it & 1);          // the row threads have consumed the staging of tile `it`
                    issue_pg(it + 1);
                    if (p.gq) {
                        mbar_wait(g_free, it & 1);       // row threads and scatter warp are done with the g_q tile
                        issue_g(it + 1);
                    }
                }
                if (it + 2 < n_my) {
                    mbar_wait(&x_free[it & 1], (it >> 1) & 1);
                    issue_x(it + 2);
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer =======================================================
        if (lane == 0) {
            const int ks1 = (K + 15) >> 4;                       // 16 codes per K-step of GEMM 1
            const uint32_t n2 = (uint32_t)ks1 * 16;              // codes covered by GEMM 2's N
            const uint32_t IDESC1 = umma_idesc(0u, HM, D) | UMMA_B_MN;
            const uint32_t IDESC2 = umma_idesc(0u, HM, n2) | UMMA_A_MN | UMMA_B_MN;
            for (int it = 0; it < n_my; ++it) {
                const uint8_t* xb = sX + (it & 1) * TILE;
                mbar_wait(coef_ready, it & 1);
                tcgen05_fence_after();
                // GEMM 1: D1[r][d] = sum_k C'[r][k] E'[k][d]   (hi.hi + lo.hi + hi.lo)
                for (int ks = 0; ks < ks1; ++ks) {
                    const uint64_t ah = umma_desc_sw128(sCh) + 2 * ks, al = umma_desc_sw128(sCl) + 2 * ks;
                    const uint64_t bh = umma_desc_sw128_mn(sEh + ks * 2048, 8192, 1024);
                    const uint64_t bl = umma_desc_sw128_mn(sEl + ks * 2048, 8192, 1024);
                    umma_bf16(d1, ah, bh, IDESC1, ks != 0);
                    umma_bf16(d1, al, bh, IDESC1, true);
                    umma_bf16(d1, ah, bl, IDESC1, true);
                }
                umma_commit(d1_done);
                // GEMM 2: D2[d][k] = sum_r x''[r][d] C'[r][k]: TMEM lanes 0..63 take x_hi, lanes 64..127 x_lo
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {                 // 16 rows per K-step
                    const uint64_t a = umma_desc_sw128_mn(xb + ks * 2048, HBLK, 1024);
                    umma_bf16(d2, a, umma_desc_sw128_mn(sCh + ks * 2048, HBLK, 1024), IDESC2, ks != 0);
                    umma_bf16(d2, a, umma_desc_sw128_mn(sCl + ks * 2048, HBLK, 1024), IDESC2, true);
                }
                umma_commit(mma_done);
            }
        }
    } else if (warp == 2 || warp == 3) {
        // =============================== index-keyed scatter of g_q =======================================
        // Step 1 (warp 3): stable counting sort of the tile's 128 rows by code.  ord[pos] = row-offset | code << 16 |
        //   (last row of its code ? 1 << 24 : 0), rows of equal code adjacent and in row order; rows beyond N carry
        //   the sentinel code K and sort last.  The sorted list is cut into four segments at run boundaries.
        // Step 2 (both warps): each half-warp walks one segment, 16 lanes x float4 = one g_q row per step.  A run is
        //   summed in registers; its accumulator row is loaded when the run ends and written back when the NEXT run
        //   ends (every code has one run per tile and segments own disjoint codes, so nothing aliases): no
        //   shared-memory round trip sits on the per-row path, there are no atomics, and the summation order is the
        //   row order -- reproducible.
        if (do_scatter) {
            const unsigned lt = (1u << lane) - 1u;
            const int half = lane >> 4, j16 = lane & 15;
            const uint32_t g_s = smem_u32(sG) + (j16 >> 3) * HBLK + ((j16 & 7) << 4);   // this lane's 16-byte column of a row
            const uint32_t acc_s = smem_u32(sAcc) + j16 * 16;
            const int seg = 2 * (warp - 2) + half;
            int tw_n = 0;
            for (int it = 0; it < n_my; ++it) {
                const int b = it & 1;
                int* ord = sOrd + b * HM;
                int* segs = sSeg + b * 8;
                if (warp == 3) {
                    VQB_HTW(80, 30);
                    if (it >= 2) mbar_wait(&sort_free[b], ((it >> 1) - 1) & 1);   // warp 2 has finished with this list
                    const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * HM;
                    const int rows = min(HM, p.N - row0);
                    int cj[4], before[4], rank[4], last[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rr = 32 * j + lane;
                        const long long k = rr < rows ? p.idx[row0 + rr] : (long long)K;
                        cj[j] = k < 0 ? 0 : (k > K ? K - 1 : (int)k);
                    }
                    for (int i = lane; i < 96; i += 32) sCnt[i] = 0;
                    __syncwarp();
                    VQB_HTW(80, 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const unsigned peers = __match_any_sync(0xffffffffu, cj[j]);
                        rank[j] = __popc(peers & lt);
                        before[j] = sCnt[cj[j]];
                        __syncwarp();
                        if (rank[j] == 0) sCnt[cj[j]] = before[j] + __popc(peers);
                        __syncwarp();
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) last[j] = (before[j] + rank[j] == sCnt[cj[j]] - 1) ? (1 << 24) : 0;
                    __syncwarp();
                    // exclusive scan over the K+1 counters (three per lane)
                    const int a0 = sCnt[3 * lane], a1 = sCnt[3 * lane + 1], a2 = sCnt[3 * lane + 2];
                    int incl = a0 + a1 + a2;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int v = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += v;
                    }
                    const int excl = incl - (a0 + a1 + a2);
                    __syncwarp();
                    sCnt[3 * lane] = excl; sCnt[3 * lane + 1] = excl + a0; sCnt[3 * lane + 2] = excl + a0 + a1;
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rr = 32 * j + lane;
                        ord[sCnt[cj[j]] + before[j] + rank[j]] = (rr * 128) | ((rr & 7) << 4) | (cj[j] << 16) | last[j];
                    }
                    __syncwarp();
                    // segment cuts: the start of the run that holds position 32 q (q = 1..3), clipped to the valid rows
                    if (lane < 5) {
                        const int n_valid = sCnt[K];
                        int cut = lane == 0 ? 0 : n_valid;
                        if (lane >= 1 && lane <= 3 && 32 * lane < n_valid) cut = sCnt[(ord[32 * lane] >> 16) & 0xFF];
                        segs[lane] = cut;
                    }
                    __syncwarp();
                    VQB_HTW(80, 33);
                    if (lane == 0) mbar_arrive(&sort_full[b]);
                } else {
                    VQB_HTW(40, 20);
                    mbar_wait(&sort_full[b], (it >> 1) & 1);
                    VQB_HTW(40, 21);
                }
                mbar_wait(g_full, it & 1);
                VQB_HTW(40, 22);
                const int beg = segs[seg], end = segs[seg + 1];
                const int len = end - beg;
                const int steps = max(len, __shfl_xor_sync(0xffffffffu, len, 16));
                int pend = -1;
                float4 sum = make_float4(0.f, 0.f, 0.f, 0.f), psum = sum, pbase = sum;
#pragma unroll 1
                for (int i = 0; i < steps; i += 4) {
                    int e[4];
                    float4 g[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) e[u] = i + u < len ? ord[beg + i + u] : -1;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        g[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (e[u] >= 0) {
                            const uint32_t a_ = (g_s + (e[u] & 0x3F80)) ^ (e[u] & 0x70);
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g[u].x), "=f"(g[u].y), "=f"(g[u].z), "=f"(g[u].w) : "r"(a_));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        if (e[u] >= 0) {
                            sum.x += g[u].x; sum.y += g[u].y; sum.z += g[u].z; sum.w += g[u].w;
                            if (e[u] & (1 << 24)) {                // this row closes the run of its code
                                const int code = (e[u] >> 16) & 0xFF;
                                if (pend >= 0)
                                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(acc_s + pend * 256), "f"(pbase.x + psum.x),
                                                 "f"(pbase.y + psum.y), "f"(pbase.z + psum.z), "f"(pbase.w + psum.w) : "memory");
                                pend = code; psum = sum;
                                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(pbase.x), "=f"(pbase.y), "=f"(pbase.z), "=f"(pbase.w)
                                             : "r"(acc_s + code * 256) : "memory");
                                sum = make_float4(0.f, 0.f, 0.f, 0.f);
                            }
                        }
                    }
                }
                if (pend >= 0)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(acc_s + pend * 256), "f"(pbase.x + psum.x),
                                 "f"(pbase.y + psum.y), "f"(pbase.z + psum.z), "f"(pbase.w + psum.w) : "memory");
                __syncwarp();
                VQB_HTW(40, 23);
                if (lane == 0) {
                    mbar_arrive(g_free);
                    if (warp == 2) mbar_arrive(&sort_free[b]);
                }
            }
            asm volatile("bar.arrive 3, 192;" ::: "memory");
        }
    } else if (warp >= 4) {
        // =============================== row threads ======================================================
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;                              // row within the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const float tau = l2 ? fmaxf(__ldg(p.temp), 0.f) : 1.f;
        const float cmul = l2 ? -tau : 1.f;
        const float uE = pow2i(gE);
        float acc[64];                                             // D2 of all tiles: lane r = d (r < 64: x_hi part, else x_lo part)
#pragma unroll
        for (int k = 0; k < 64; ++k) acc[k] = 0.f;
        float cs0 = 0.f, cs1 = 0.f;                                // two column sums per lane (see the butterfly below)
        int tl_n = 0;
        VQB_HTL(1);
        for (int it = 0; it < n_my; ++it) {
            const uint32_t ph = it & 1;
            const int tile = (int)blockIdx.x + it * (int)gridDim.x;
            const int row0 = tile * HM;
            const int rows = min(HM, p.N - row0);
            const bool valid = r < rows;
            const bool real = valid && (p.n_real <= 0 || row0 + r < p.n_real);
            uint8_t* xb = sX + (it & 1) * TILE;
            VQB_HTL(2);
            mbar_wait(pg_full, ph);
            VQB_HTL(3);
            const int nfl = rows * K, nbulk = ((nfl * 4) & ~15) >> 2;
            if (nbulk != nfl) {                                    // last < 16 bytes of a ragged tile
                if (r < nfl - nbulk) {
                    stP[nbulk + r] = p.p[(size_t)row0 * K + nbulk + r];
                    stG[nbulk + r] = p.gp[(size_t)row0 * K + nbulk + r];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            // ---- softmax backward for row r -----------------------------------------------------------------
            float c[64];
            float rsum, m;
            {
                float gg[64];
#pragma unroll
                for (int k = 0; k < 64; ++k) {
                    const bool on = valid && k < K;
                    c[k] = on ? stP[r * K + k] : 0.f;
                    gg[k] = on ? stG[r * K + k] : 0.f;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(pg_free);               // staging consumed: next tile's blocks may land
                float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 64; ++k) s4[k & 3] = fmaf(gg[k], c[k], s4[k & 3]);
                const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
                float r4[4] = {0.f, 0.f, 0.f, 0.f}, m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 64; ++k) {
                    c[k] = cmul * (c[k] * (gg[k] - s));
                    r4[k & 3] += c[k];
                    m4[k & 3] = fmaxf(m4[k & 3], fabsf(c[k]));
                }
                rsum = (r4[0] + r4[1]) + (r4[2] + r4[3]);
                m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
            }
            VQB_HTL(4);
            // ---- x row, scales ---------------------------------------------------------------------------------
            mbar_wait(&x_full[it & 1], (it >> 1) & 1);
            float xr[64];
            float mx = 0.f;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const float4 v = *reinterpret_cast<const float4*>(xb + kb * HBLK + sw128_offset(r, ch));
                    xr[kb * 32 + 4 * ch] = v.x; xr[kb * 32 + 4 * ch + 1] = v.y;
                    xr[kb * 32 + 4 * ch + 2] = v.z; xr[kb * 32 + 4 * ch + 3] = v.w;
                    mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
                }
            }
            const bool nz = m > 0.f;                               // (a NaN row maximum scales by 1 and propagates)
            int er = nz ? exp_of(m) - 14 : 0;
            er = er < -126 ? -126 : (er > 126 ? 126 : er);
            const int tr = (real && nz && mx > 0.f) ? exp_of(mx) + er : INT_MIN / 2;
            {
                const int wm = __reduce_max_sync(0xffffffffu, tr);
                if (lane == 0) sRed[q4] = wm;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
            const int tmax = max(max(sRed[0], sRed[1]), max(sRed[2], sRed[3]));
            const int t = tmax > INT_MIN / 4 ? tmax - 14 : 0;
            const float sc = pow2i(-er);
            const float sx = (real && nz && er - t >= -126) ? pow2i(er - t) : 0.f;
            // x'' -> fp16 hi | lo, in place over the raw tile (this thread has read the whole of row r above)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 hi, lo;
                split8(xr + 8 * j, sx, hi, lo);
                *reinterpret_cast<uint4*>(xb + sw128_offset(r, j)) = hi;
                *reinterpret_cast<uint4*>(xb + HBLK + sw128_offset(r, j)) = lo;
            }
            // C' -> fp16 hi | lo
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 hi, lo;
                split8(c + 8 * j, sc, hi, lo);
                *reinterpret_cast<uint4*>(sCh + sw128_offset(r, j)) = hi;
                *reinterpret_cast<uint4*>(sCl + sw128_offset(r, j)) = lo;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(coef_ready);
            VQB_HTL(5);
            // ---- column sums of C* while the MMAs run: butterfly transpose-reduce over the warp ----------------
            // after the five steps lane L holds the sums of columns 2L and 2L+1 in c[0], c[1]
            {
#pragma unroll
                for (int k = 0; k < 64; ++k) c[k] = real ? c[k] : 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const bool up = (lane & 16) != 0;
                    const float send = up ? c[i] : c[i + 32], keep = up ? c[i + 32] : c[i];
                    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const bool up = (lane & 8) != 0;
                    const float send = up ? c[i] : c[i + 16], keep = up ? c[i + 16] : c[i];
                    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const bool up = (lane & 4) != 0;
                    const float send = up ? c[i] : c[i + 8], keep = up ? c[i + 8] : c[i];
                    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool up = (lane & 2) != 0;
                    const float send = up ? c[i] : c[i + 4], keep = up ? c[i + 4] : c[i];
                    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const bool up = (lane & 1) != 0;
                    const float send = up ? c[i] : c[i + 2], keep = up ? c[i + 2] : c[i];
                    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
                }
                cs0 += c[0]; cs1 += c[1];
            }
            VQB_HTL(6);
            // ---- dx = g_q + 2 x rowsum(C) - 2 (C @ E)   |   C @ W ---------------------------------------------------
            // Formed in registers (over the dead coefficient array) while GEMM 2 may still be running, then staged
            // in this tile's x buffer, which is dead once GEMM 2 has completed: the g_q tile is only read here, so
            // the scatter warps never sit on this critical path.
            mbar_wait(d1_done, ph);
            tcgen05_fence_after();
            VQB_HTL(7);
            if (p.gq) mbar_wait(g_full, ph);
            VQB_HTL(8);
            {
                const float u1 = pow2i(er);
                const float r2 = l2 ? 2.f * rsum : 0.f;
                const float ad = l2 ? -2.f : 1.f;
                const bool add_g = l2 && p.gq != nullptr;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    float a[32];
                    tmem_ld_32x32(d1 + lane_addr + kb * 32, a);
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (add_g) g = *reinterpret_cast<const float4*>(sG + kb * HBLK + sw128_offset(r, ch));
                        c[kb * 32 + 4 * ch] = fmaf(ad, a[4 * ch] * u1 * uE, fmaf(xr[kb * 32 + 4 * ch], r2, g.x));
                        c[kb * 32 + 4 * ch + 1] = fmaf(ad, a[4 * ch + 1] * u1 * uE, fmaf(xr[kb * 32 + 4 * ch + 1], r2, g.y));
                        c[kb * 32 + 4 * ch + 2] = fmaf(ad, a[4 * ch + 2] * u1 * uE, fmaf(xr[kb * 32 + 4 * ch + 2], r2, g.z));
                        c[kb * 32 + 4 * ch + 3] = fmaf(ad, a[4 * ch + 3] * u1 * uE, fmaf(xr[kb * 32 + 4 * ch + 3], r2, g.w));
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && p.gq) mbar_arrive(g_free);            // this warp has read its rows of the g_q tile
            VQB_HTL(9);
            // ---- D2 of this tile -> registers; the x buffer becomes the dx staging tile ---------------------------
            mbar_wait(mma_done, ph);
            tcgen05_fence_after();
            VQB_HTL(10);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    *reinterpret_cast<float4*>(xb + kb * HBLK + sw128_offset(r, ch)) =
                        make_float4(c[kb * 32 + 4 * ch], c[kb * 32 + 4 * ch + 1], c[kb * 32 + 4 * ch + 2], c[kb * 32 + 4 * ch + 3]);
            fence_proxy_async_smem();
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (r == 0) {
                tma_store_2d(&tm_dx, xb, 0, row0);
                tma_store_2d(&tm_dx, xb + HBLK, 32, row0);
                tma_store_commit();
            }
            {
                const float t1 = pow2i(t >> 1), t2 = pow2i(t - (t >> 1));
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    float a[32];
                    tmem_ld_32x32(d2 + lane_addr + hb * 32, a);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[hb * 32 + j] = fmaf(a[j] * t1, t2, acc[hb * 32 + j]);
                }
            }
            tcgen05_fence_before();
            if (r == 0) {
                tma_store_wait_read();                             // dx has left shared memory
                mbar_arrive(&x_free[it & 1]);                      // ... and GEMM 2 has long finished with this buffer
            }
            VQB_HTL(11);
        }
        if (r == 0) tma_store_wait_all();
        // every row thread must be behind r == 0's wait before the dx staging tile (sX) is reused below -- without the
        // scatter warps (skip branch) barrier 3 is not taken, so the row group synchronises on its own barrier
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (do_scatter) asm volatile("bar.sync 3, 192;" ::: "memory");   // both scatter warps have finished their last tile
        VQB_HTL(12);

        // ---- once per CTA: the K x D sums -> this CTA's partial record -------------------------------------------
        // (every row thread has waited for the last tile's mma_done: sX is dead; sAcc is complete after barrier 3)
        float* part = p.partial + (size_t)blockIdx.x * H_PARTIAL_FLOATS;
        float* sXch = reinterpret_cast<float*>(sX);                // [64][64] x_lo halves, then [4][64] column sums
        float* sCs = sXch + H_KD;
        if (r >= 64) {
#pragma unroll
            for (int k = 0; k < 64; ++k) sXch[k * 64 + (r - 64)] = acc[k];
        }
        sCs[q4 * 64 + 2 * lane] = cs0;                             // the butterfly leaves columns 2L, 2L+1 in lane L
        sCs[q4 * 64 + 2 * lane + 1] = cs1;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (r < 64) {
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                if (k < K) {
                    const float v = acc[k] + sXch[k * 64 + r];
                    if (l2) {
                        const float e = fmaf(-2.f, v, sAcc[k * 64 + r]);
                        part[k * 64 + r] = e;
                        // projected columns once more, column-major: the tail's projection blocks read them coalesced
                        if (r >= p.t_first) part[H_KD + r * 64 + k] = e;
                    } else {
                        part[k * 64 + r] = v;
                        part[H_KD + k * 64 + r] = sAcc[k * 64 + r];
                    }
                }
            }
            part[2 * H_KD + r] = (sCs[r] + sCs[64 + r]) + (sCs[128 + r] + sCs[192 + r]);
        }
        VQB_HTL(13);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// out[i] += sum over the CTAs' partial records, in a fixed order (deterministic).  One block = 32 consecutive
// outputs x 32 slices of the CTA list (independent loads, combined through shared memory in slice order).
__global__ void __launch_bounds__(1024)
reduce_partials_h2_kernel(const float* __restrict__ partial, int n_cta, int K, float* __restrict__ dW,
                          float* __restrict__ dG, float* __restrict__ colsum) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int n_kd = K * 64;
    const int n_planes = dG ? 2 : 1;
    const int o = blockIdx.x * 32 + tx;                            // output index over [planes x n_kd | 64 column sums]
    const float* src = nullptr;
    float* dst = nullptr;
    if (o < n_planes * n_kd) {
        const int plane = o / n_kd, i = o - plane * n_kd;
        src = partial + plane * H_KD + i;
        dst = (plane ? dG : dW) + i;
    } else if (o - n_planes * n_kd < K) {
        const int k = o - n_planes * n_kd;
        src = partial + 2 * H_KD + k;
        dst = colsum + k;
    }
    float a = 0.f;
    if (src) {
#pragma unroll 5
        for (int cta = ty; cta < n_cta; cta += 32) a += __ldg(src + (size_t)cta * H_PARTIAL_FLOATS);
    }
    red[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && dst) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) s += red[j][tx];
        *dst += s;
    }
}

// -----------------------------------------------------------------------------------------------------------
// Fused tail (vqb_bwd_tail, L2 score): partial sums -> parameter gradients -> sum over GPUs, block-parallel.
//
// Every block owns a few outputs of the flat gradient  d_learnable | d_proj_w | d_proj_b  and finishes them alone:
//   learnable blocks (32 consecutive outputs (k, d < D_l)):   fixed-order sum over the CTAs' partial records of
//       dE[k][d] and of the column sum cs[k];  out = dE + 2 E[k][d] cs[k]                     (src/embed.py:109-112, :211)
//   projection blocks (one per projected column j):  eff[k] = dE[k][D_l+j] + 2 E[k][D_l+j] cs[k] for all k, then
//       d_proj_w[j][a] = sum_k eff[k] attr[k][a],  d_proj_b[j] = sum_k eff[k]
// Data-parallel runs: the block then PUSHES its outputs into every peer's exchange buffer over NVLink as 8-byte
// (value, epoch) words -- the data carries its own flag (NCCL's "LL" idea), so there is no fence and no separate
// signal -- and polls its own buffer until the same outputs of every peer have arrived; the sum runs in rank order, so
// all GPUs end with the same bits.  No block waits for another block of its own GPU, remote blocks push before they
// poll: no deadlock whatever the residency.  Two slots alternate by epoch parity (a rank cannot run two exchanges
// ahead of a peer, because each exchange needs that peer's data of the same epoch).  Waits are bounded (tail.timeout_ms,
// minutes by default, like a collective library's watchdog); a timed-out wait raises counter[2] and returns -- no trap.
// The last block to finish (ticket counter) hands the ticket back and publishes the epoch for the next call.
// -----------------------------------------------------------------------------------------------------------
struct TailP {
    const float* table;       // [K][64]
    const float* attr;        // [K][A] or NULL
    float* d_flat;
    unsigned int* counter;    // [0] ticket, [1] epoch
    void* const* peer_bufs;   // device array [world]
    unsigned long long* dbg;  // optional timeline buffer (developer hook), slots 100..
    unsigned long long timeout_ns;
    int A, Da, world, rank, n_learn_blocks;
};
#define VQB_TTL(slot) do { if (t.dbg && tid == 0 && blockIdx.x == gridDim.x - 1) t.dbg[100 + (slot)] = globaltimer_ns(); } while (0)

__device__ __forceinline__ void st_relaxed_sys_b64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_b64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// exchange buffer of one rank: [2 slots][world senders][n_pad] 8-byte words (value | epoch << 32)
__host__ __device__ inline size_t exch_words(int64_t n_flat, int world) { return 2 * (size_t)world * (size_t)((n_flat + 3) & ~3ll); }

__global__ void __launch_bounds__(1024)
bwd_tail_h2_kernel(const float* __restrict__ partial, int n_cta, int K, TailP t) {
    __shared__ float red[32][33], red2[32][33], red3[32][33], red4[32][33];
    __shared__ float s_eff[64], s_tab[64];
    extern __shared__ float s_attr[];        // [K][A] (projection blocks)
    __shared__ float s_out[64];              // this block's finished outputs
    __shared__ int s_idx[64];                // their positions in the flat gradient
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 32 + tx;
    const int Dl = 64 - t.Da;
    const int n_l = K * Dl, n_w = t.Da * t.A, n_flat = n_l + n_w + t.Da;
    const bool exchange = t.world > 1;
    pdl_launch();                                                   // a PDL-launched successor (the next table assembly) may queue up
    if (t.dbg && tid == 0 && blockIdx.x == gridDim.x - 1) t.dbg[100] = globaltimer_ns();
    if ((int)blockIdx.x >= t.n_learn_blocks) {
        // constants of a projection block (frozen attribute table, this step's codebook column): fetched while the main
        // backward kernel is still running -- neither is written by it
        const int j = blockIdx.x - t.n_learn_blocks;
        for (int i = tid; i < K * t.A; i += 1024) s_attr[i] = __ldg(t.attr + i);
        if (tid < K) s_tab[tid] = __ldg(t.table + tid * 64 + Dl + j);
    }
    // learnable blocks: output tx of this block and its codebook entry (a constant of this step), also ahead of the wait
    const int li = blockIdx.x * 32 + tx;
    const bool live = (int)blockIdx.x < t.n_learn_blocks && li < n_l;
    const int lk = live ? li / Dl : 0, ld = live ? li - lk * Dl : 0;
    const float ltab = live ? __ldg(t.table + lk * 64 + ld) : 0.f;
    pdl_wait();                                                     // the main backward kernel has completed
    VQB_TTL(1);
    const unsigned int epoch = exchange ? *reinterpret_cast<volatile unsigned int*>(t.counter + 1) + 1u : 0u;
    int n_mine = 0;                                                 // outputs finished by this block (block-uniform)

    if ((int)blockIdx.x < t.n_learn_blocks) {
        // ---- 32 consecutive learnable outputs --------------------------------------------------------------------
        const int i = li, k = lk, d = ld;
        float a = 0.f, c = 0.f;
        if (live) {
#pragma unroll 5
            for (int cta = ty; cta < n_cta; cta += 32) {
                const float* rec = partial + (size_t)cta * H_PARTIAL_FLOATS;
                a += __ldg(rec + k * 64 + d);
                c += __ldg(rec + 2 * H_KD + k);
            }
        }
        red[ty][tx] = a; red2[ty][tx] = c;
        __syncthreads();
        if (ty == 0) {
            float sa = 0.f, sc = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { sa += red[j][tx]; sc += red2[j][tx]; }
            s_out[tx] = live ? fmaf(2.f * ltab, sc, sa) : 0.f;
            s_idx[tx] = live ? i : -1;
        }
        n_mine = 32;
    } else {
        // ---- one projected column j: eff[k] for all codes, then its A weights and its bias ------------------------
        const int j = blockIdx.x - t.n_learn_blocks;
        float a[2] = {0.f, 0.f}, c[2] = {0.f, 0.f};                  // codes tx and tx + 32 (K <= 64)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = tx + 32 * h;
            if (k < K) {
#pragma unroll 5
                for (int cta = ty; cta < n_cta; cta += 32) {
                    const float* rec = partial + (size_t)cta * H_PARTIAL_FLOATS;
                    a[h] += __ldg(rec + H_KD + (Dl + j) * 64 + k);   // column-major copy: lanes read consecutive codes
                    c[h] += __ldg(rec + 2 * H_KD + k);
                }
            }
        }
        red[ty][tx] = a[0]; red2[ty][tx] = c[0]; red3[ty][tx] = a[1]; red4[ty][tx] = c[1];
        __syncthreads();
        if (ty < 2) {
            const int k = tx + 32 * ty;
            if (k < K) {
                float sa = 0.f, sc = 0.f;
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) { sa += (ty ? red3 : red)[jj][tx]; sc += (ty ? red4 : red2)[jj][tx]; }
                s_eff[k] = fmaf(2.f * s_tab[k], sc, sa);
            }
        }
        __syncthreads();
        if (tid <= t.A) {                                           // tid < A: weight (j, a = tid);  tid == A: bias j
            float acc = 0.f;
            for (int k = 0; k < K; ++k) acc = tid < t.A ? fmaf(s_eff[k], s_attr[k * t.A + tid], acc) : acc + s_eff[k];
            s_out[tid] = acc;
            s_idx[tid] = tid < t.A ? n_l + j * t.A + tid : n_l + n_w + j;
        }
        n_mine = t.A + 1;                                           // A <= 63 (checked by the host)
    }
    __syncthreads();
    VQB_TTL(2);

    if (!exchange) {
        if (tid < n_mine && s_idx[tid] >= 0) t.d_flat[s_idx[tid]] = s_out[tid];
        return;
    }

    // ---- push to every peer (own buffer included), then gather the peers' words of the same outputs ---------------
    const int n_pad = (n_flat + 3) & ~3;
    const size_t slot_off = (size_t)(epoch & 1u) * t.world * n_pad;
    for (int w = tid; w < n_mine * t.world; w += 1024) {
        const int o = w % n_mine, r = w / n_mine;                   // consecutive threads: consecutive words of one peer
        const int i = s_idx[o];
        if (i >= 0) {
            unsigned long long* dst = reinterpret_cast<unsigned long long*>(t.peer_bufs[r]) + slot_off + (size_t)t.rank * n_pad + i;
            st_relaxed_sys_b64(dst, ((unsigned long long)epoch << 32) | __float_as_uint(s_out[o]));
        }
    }
    VQB_TTL(3);
    if (tid < n_mine && s_idx[tid] >= 0) {
        const int i = s_idx[tid];
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(t.peer_bufs[t.rank]) + slot_off + i;
        float sum = 0.f;
        const unsigned long long t0 = globaltimer_ns();
        for (int r = 0; r < t.world; ++r) {                         // rank order: identical bits on every GPU
            unsigned long long w = ld_relaxed_sys_b64(mine + (size_t)r * n_pad);
            bool gave_up = false;
            while ((unsigned int)(w >> 32) != epoch) {
                if (globaltimer_ns() - t0 > t.timeout_ns) {
                    // no trap: the context survives; the host finds the flag when it synchronises (dist.check_exchange)
                    atomicMax(t.counter + 2, (unsigned int)(r + 1));
                    gave_up = true;
                    break;
                }
                w = ld_relaxed_sys_b64(mine + (size_t)r * n_pad);
            }
            if (!gave_up) sum += __uint_as_float((unsigned int)w);
        }
        t.d_flat[i] = sum;
    }
    VQB_TTL(4);
    // ---- housekeeping: the last block hands the ticket back and publishes the epoch ----------------------------------
    __syncthreads();
    if (tid == 0) {
        if (atomicAdd(t.counter, 1u) == gridDim.x - 1) { t.counter[0] = 0u; t.counter[1] = epoch; }
    }
}

// -----------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------
unsigned long long* get_debug_timeline();

size_t exchange_bytes(int64_t n_flat, int world) { return exch_words(n_flat, world) * 8; }

bool backward_h2_supported(const vqb_bwd_args* a) {
    if (!(a->flags & VQB_TENSOR_CORES)) return false;
    if (!a->g_p || !(a->flags & VQB_STOP_GRAD) || (a->flags & VQB_TEMP_GRAD)) return false;
    if (a->n_codes > 64 || a->dim != 64) return false;
    return aligned16(a->p_code) && aligned16(a->g_p);
}

int backward_h2_workspace(const vqb_bwd_args* a, size_t* bytes) {
    *bytes = backward_h2_supported(a) ? (size_t)sm_count() * H_PARTIAL_FLOATS * 4 : 0;
    return VQB_OK;
}

int launch_backward_h2(const vqb_bwd_args* a, cudaStream_t s) {
    const int64_t N = a->n_rows, K = a->n_codes, D = a->dim;
    const size_t need = (size_t)sm_count() * H_PARTIAL_FLOATS * 4;
    if (!a->workspace || a->workspace_bytes < need) {
        set_error("vqb_backward: workspace too small (%zu < %zu bytes)", a->workspace_bytes, need);
        return VQB_ERR_WORKSPACE;
    }
    const bool l2 = (a->flags & VQB_SCORE_L2) != 0;
    CUtensorMap tx, tg, td;
    int rc;
    if ((rc = make_tmap_2d_f32(&tx, a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, HM))) return rc;
    if ((rc = make_tmap_2d_f32(&tg, a->g_q ? a->g_q : a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, HM))) return rc;
    if ((rc = make_tmap_2d_f32(&td, a->dx, (uint64_t)N, (uint64_t)D, (uint64_t)D, HM))) return rc;

    BwdH2P p;
    p.p = a->p_code; p.gp = a->g_p; p.gq = a->g_q; p.idx = (const long long*)a->idx; p.temp = a->temp;
    p.E = a->score_w;
    p.partial = reinterpret_cast<float*>(a->workspace);
    p.dbg = get_debug_timeline();
    p.N = (int)N; p.K = (int)K; p.n_real = (int)(a->n_real_rows > 0 && a->n_real_rows < N ? a->n_real_rows : 0);
    p.num_tiles = (int)ceil_div(N, HM);
    p.flags = a->flags;
    p.t_first = (a->tail && l2) ? 64 - (int)a->tail->dim_attr : 64;

    const int stage_bytes = (int)((HM * K * 4 + 127) & ~127);
    const size_t smem = (size_t)2 * 2 * HBLK + 2 * HBLK + 2 * HBLK + 2 * 64 * 128 + 2 * (size_t)stage_bytes + 64 * 64 * 4 +
                        2 * HM * 4 + 96 * 4 + 16 * 4 + 8 * 4 + 16 * 8 + 16 + 1024;
    if ((int)smem > max_optin_smem()) return invalid("vqb_backward: fp16x2 kernel needs %zu B of shared memory", smem);
    VQB_CUDA(cudaFuncSetAttribute(vqb_bwd_h2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    // a plain stream-ordered launch: what precedes the backward in the stream (the producer of g_p / g_q, possibly a
    // copy) is not ours to overlap.  The kernel still releases its own successor early (the fused tail).
    kernel_event_begin(s);
    vqb_bwd_h2_kernel<<<grid, H_THREADS, smem, s>>>(tx, tg, td, p, stage_bytes);
    kernel_event_end(s);
    VQB_CHECK_LAUNCH("vqb_bwd_h2_kernel");
    return launch_bwd_reduce(a, p.partial, grid, s, p.dbg);
}

// Behind either main backward kernel (vqb_bwd_h2_kernel, vqb_bwd_pcode_kernel): the per-CTA partial records -> gradients.
// With a tail: ONE kernel (fixed-order sum, table backward, sum over GPUs), PDL-chained to the main kernel; otherwise the
// fixed-order reduction into the caller's accumulators.
int launch_bwd_reduce(const vqb_bwd_args* a, const float* partial, int grid, cudaStream_t s, unsigned long long* dbg) {
    const int64_t K = a->n_codes;
    const bool l2 = (a->flags & VQB_SCORE_L2) != 0;
    if (a->tail) {
        const vqb_bwd_tail* tl = a->tail;
        TailP t;
        t.table = a->gather_table; t.attr = tl->phn_attr; t.d_flat = tl->d_flat; t.counter = tl->counter;
        t.timeout_ns = (unsigned long long)(tl->timeout_ms ? tl->timeout_ms : 120000u) * 1000000ull;
        t.peer_bufs = tl->peer_bufs; t.dbg = dbg; t.A = (int)tl->n_attr; t.Da = (int)tl->dim_attr; t.world = tl->world; t.rank = tl->rank;
        const int Dl = 64 - t.Da;
        t.n_learn_blocks = (int)ceil_div(K * Dl, 32);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(t.n_learn_blocks + t.Da)); cfg.blockDim = dim3(32, 32); cfg.stream = s;
        cfg.dynamicSmemBytes = (size_t)K * t.A * 4;                  // <= 64 * 63 * 4 B
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = getenv("VQB_NO_PDL") ? 0 : 1;
        VQB_CUDA(cudaLaunchKernelEx(&cfg, bwd_tail_h2_kernel, partial, grid, (int)K, t));
        VQB_CHECK_LAUNCH("bwd_tail_h2_kernel");
        return VQB_OK;
    }
    float* dG = l2 ? nullptr : a->d_gather;
    const int n_out = (int)((dG ? 2 : 1) * K * 64 + K);
    reduce_partials_h2_kernel<<<(unsigned)ceil_div(n_out, 32), dim3(32, 32), 0, s>>>(partial, grid, (int)K, a->d_score_w, dG,
                                                                                  a->colsum);
    VQB_CHECK_LAUNCH("reduce_partials_h2_kernel");
    return VQB_OK;
}

}  // namespace vqb
