// Backward of the quantizer on the 5th-generation tensor cores (parity mode: K <= 64, D = 64, stop_grad).
//
// Autograd of src/embed.py:105-147 / :187-205 (entered from src/solver.py:144); algebra in DESIGN.md:
//   Gs = P * (g_p - rowsum(g_p * P));   C = -tau Gs (L2)  |  Gs (LINEAR)
//   dx  = g_q + 2 x rowsum(C) - 2 C @ E          (L2)     |  C @ W           (LINEAR)
//   dE += -2 C*^T @ x + scatter_add(idx, g_q)     (L2)     |  dW += C^T @ x ; dT += scatter_add(idx, g_q)
//   colsum += colsum(C*)                          (C* = rows of "real" tiles, first_n_real_mel)
//
// One persistent CTA per SM, 8 warps, per tile of 128 rows:
//   warp 0     TMA producer: x and g_q tiles ([128][D] fp32, 128-byte swizzle) and the contiguous
//              [128 x K] blocks of p_code and g_p (1-D bulk copies) into a staging area
//   warps 4-7  thread = row: softmax backward from the staging area -> coefficient tile C (and its tf32
//              remainder C_lo) written as a swizzled operand tile; x_lo = x - trunc_tf32(x)
//   warp 1     MMA issuer (tcgen05.mma kind::tf32), two GEMMs with split operands stacked along N / M so that
//              two passes give all four hi/lo cross terms (fp32-level accuracy):
//                GEMM 1  D1[128 x 2D]  = [C ; then C_lo] (K-major A)  x  [E_hi | E_lo] (MN-major B, the table itself)
//                GEMM 2  D2[2D x 64]  += [x | x_lo]^T (MN-major A, the row tiles themselves) x [C ; then C_lo] (MN-major B)
//              D2 stays in TMEM for the whole kernel: the K x D reduction over all of the CTA's rows costs one
//              flush per CTA.  tf32 MN-major operands must use the 128B-swizzle/32B-atom layout, K-major ones the
//              plain 128B swizzle, so C is written in both layouts; the C_lo pass reuses the same two buffers.
//                GEMM 3  D3[* x 64]   += ones^T x [C ; then C_lo]: the column sums of C, also resident in TMEM
//   warp 2     index-keyed scatter of g_q into a shared-memory accumulator [64][D]
//   The K x D sums are flushed once per CTA as plain stores into a per-CTA partial buffer; a small reduce kernel
//   adds the partials in a fixed order (no atomics anywhere: gradients are bit-reproducible).
//   warps 4-7  then read D1 from TMEM, form dx and stage it for coalesced 128-bit stores.
#include <cudaTypedefs.h>
#include <math.h>
#include "vqb_common.cuh"
#include "vqb_tc.cuh"

namespace vqb {
using namespace tc;

constexpr int BBM = 128;                // rows per tile
constexpr int BXBLK = BBM * 128;        // one row-tile K-block: [128 rows][32 fp32] = 16 KB
constexpr int BEBLK = 64 * 128;         // one table group: [64 codes][32 fp32] = 8 KB
constexpr int BWD_THREADS = 256;
// per-CTA partial record: [0] x part of C^T x, [1] x_lo part, [2] scatter sums (each 64 x 64), [3] column sums (64)
constexpr int PART_KD = 64 * 64;
constexpr int PARTIAL_FLOATS = 3 * PART_KD + 64;

struct BwdTcP {
    const float* p;
    const float* gp;
    const float* gq;          // may be NULL
    const long long* idx;
    const float* temp;
    float* dx;
    float* dW;
    float* colsum;
    float* dG;                // LINEAR: scatter destination; L2: NULL (scatter goes to dW)
    float* partial;           // [grid][PARTIAL_FLOATS] per-CTA partial sums
    unsigned long long* dbg;  // optional timeline buffer (developer hook)
    int N, K, n_real, num_tiles;
    unsigned flags;
};

#define VQB_BTL(tag) do { if (p.dbg && r == 0 && blockIdx.x == 0 && tl_n < 120) { p.dbg[tl_n++] = ((unsigned long long)(tag) << 56) | (globaltimer_ns() & 0x00FFFFFFFFFFFFFFull); } } while (0)

__device__ __forceinline__ float tf32_lo(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }

template <int KB>   // D = 32 * KB
__global__ void __launch_bounds__(BWD_THREADS, 1)
vqb_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                  const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo, BwdTcP p) {
    constexpr int D = 32 * KB;
    constexpr int TILE = KB * BXBLK;                 // one [128][D] tile
    constexpr int TMEM_COLS = 256;                   // D1: 2D columns at 0, D2: 64 columns at 128
    static_assert(KB == 2, "instantiated for D = 64");

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS)
    uint8_t* sX = smem;                              // [KB][16 KB] x tile            (GEMM 2: A groups 0..KB-1)
    uint8_t* sXlo = sX + TILE;                       // [KB][16 KB] x_lo, later dx    (GEMM 2: A groups KB..2KB-1)
    uint8_t* sG = sXlo + TILE;                       // [KB][16 KB] g_q tile
    uint8_t* sC = sG + TILE;                         // [2][16 KB]  C, then C_lo: K-major, SW128       (GEMM 1 A); staging of p_code before
    uint8_t* sCmn = sC + 2 * BXBLK;                  // [2][16 KB]  C, then C_lo: MN-major, SW128/32B  (GEMM 2 B); staging of g_p before
    uint8_t* sE = sCmn + 2 * BXBLK;                  // [2KB][8 KB] E_hi groups, then E_lo groups (SW128/32B)
    float* sAcc = reinterpret_cast<float*>(sE + 2 * KB * BEBLK);    // [64][D] scatter accumulator
    uint8_t* sOnes = reinterpret_cast<uint8_t*>(sAcc + 64 * D);     // [128 rows][32 fp32] of 1.0 (GEMM 3 A operand)
    int* sIdx = reinterpret_cast<int*>(sOnes + BXBLK);              // [128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sIdx + BBM);
    uint64_t* e_full = bars;
    uint64_t* in_full = bars + 1;
    uint64_t* coef_ready = bars + 2;
    uint64_t* mma_done = bars + 3;
    uint64_t* scat_done = bars + 4;
    uint64_t* tile_free = bars + 5;
    uint64_t* p1_done = bars + 6;
    uint64_t* coef2_ready = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool l2 = (p.flags & VQB_SCORE_L2) != 0;
    const bool do_scatter = p.gq != nullptr && !(l2 && (p.flags & VQB_SKIP));
    const int K = p.K;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_g); tma_prefetch_desc(&tm_hi); tma_prefetch_desc(&tm_lo);
        mbar_init(e_full, 1); mbar_init(in_full, 1); mbar_init(coef_ready, 4);
        mbar_init(mma_done, 1); mbar_init(scat_done, 1); mbar_init(tile_free, 4);
        mbar_init(p1_done, 1); mbar_init(coef2_ready, 4);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<TMEM_COLS>(tmem_slot);
    for (int i = threadIdx.x; i < 64 * D; i += BWD_THREADS) sAcc[i] = 0.f;
    for (int i = threadIdx.x; i < BXBLK / 16; i += BWD_THREADS)
        reinterpret_cast<float4*>(sOnes)[i] = make_float4(1.f, 1.f, 1.f, 1.f);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d1 = tmem_base, d2 = tmem_base + 128, d3 = tmem_base + 192;

    auto tile_is_real = [&](int tile) { return p.n_real <= 0 || (tile + 1) * BBM <= p.n_real || p.n_real >= p.N; };

    if (warp == 0) {
        // =============================== TMA producer =====================================================
        if (lane == 0) {
            mbar_arrive_expect_tx(e_full, 2 * KB * BEBLK);
            for (int g = 0; g < KB; ++g) {
                tma_load_2d(sE + (size_t)g * BEBLK, &tm_hi, g * 32, 0, e_full);
                tma_load_2d(sE + (size_t)(KB + g) * BEBLK, &tm_lo, g * 32, 0, e_full);
            }
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const uint32_t ph = it & 1;
                mbar_wait(tile_free, ph ^ 1);
                mbar_wait(scat_done, ph ^ 1);
                const int row0 = tile * BBM;
                const int rows = min(BBM, p.N - row0);
                const uint32_t bulk = (uint32_t)(rows * K * 4) & ~15u;
                mbar_arrive_expect_tx(in_full, TILE + (p.gq ? TILE : 0) + 2 * bulk);
                for (int kb = 0; kb < KB; ++kb) {
                    tma_load_2d(sX + kb * BXBLK, &tm_x, kb * 32, row0, in_full);
                    if (p.gq) tma_load_2d(sG + kb * BXBLK, &tm_g, kb * 32, row0, in_full);
                }
                if (bulk) {
                    bulk_load_1d(sC, p.p + (size_t)row0 * K, bulk, in_full);
                    bulk_load_1d(sCmn, p.gp + (size_t)row0 * K, bulk, in_full);
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer =======================================================
        if (lane == 0) {
            constexpr uint32_t IDESC1 = umma_idesc(2u, BBM, 2 * D) | UMMA_B_MN;
            constexpr uint32_t IDESC2 = umma_idesc(2u, 2 * D, 64) | UMMA_A_MN | UMMA_B_MN;
            mbar_wait(e_full, 0);
            bool d2_started = false;
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                const bool real = tile_is_real(tile);
#pragma unroll
                for (int pass = 0; pass < 2; ++pass) {             // pass 0: C, pass 1: C_lo (same buffers)
                    mbar_wait(pass ? coef2_ready : coef_ready, it & 1);
                    tcgen05_fence_after();
                    // GEMM 1: D1[r][0:D) += C.E_hi, D1[r][D:2D) += C.E_lo
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {              // 8 codes per K-step
                        const uint64_t a = umma_desc_sw128(sC + (ks >> 2) * BXBLK) + 2 * (ks & 3);
                        const uint64_t b = umma_desc_mn_32b(sE + ks * 1024, BEBLK, 512);
                        umma_tf32(d1, a, b, IDESC1, (pass | ks) != 0);
                    }
                    // GEMM 2: D2[d][k] += sum_r x[r][d] C[r][k] (TMEM lanes 0..D-1) and x_lo[r][d] C[r][k] (lanes D..2D-1)
                    if (real) {
#pragma unroll
                        for (int ks = 0; ks < 16; ++ks) {         // 8 rows per K-step
                            const uint64_t a = umma_desc_mn_32b(sX + ks * 1024, BXBLK, 512);
                            const uint64_t b = umma_desc_mn_32b(sCmn + ks * 1024, BXBLK, 512);
                            umma_tf32(d2, a, b, IDESC2, d2_started || (pass | ks) != 0);
                            // GEMM 3: every TMEM lane of D3 accumulates sum_r C[r][k] (all A groups alias the ones block)
                            umma_tf32(d3, umma_desc_mn_32b(sOnes + ks * 1024, 0, 512), b, IDESC2, d2_started || (pass | ks) != 0);
                        }
                    }
                    if (pass == 0) umma_commit(p1_done);
                }
                if (real) d2_started = true;
                umma_commit(mma_done);
            }
        }
    } else if (warp == 2) {
        // =============================== index-keyed scatter of g_q =======================================
        const int half = lane >> 4, c4 = lane & 15;               // two rows per step, 16 float4 columns each
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            mbar_wait(in_full, it & 1);
            if (do_scatter) {
                const int row0 = tile * BBM;
                const int rows = min(BBM, p.N - row0);
                for (int j = 0; j < 4; ++j) {
                    const int rr = 32 * j + lane;
                    sIdx[rr] = rr < rows ? (int)p.idx[row0 + rr] : -1;
                }
                __syncwarp();
                for (int s2 = 0; s2 < BBM / 2; ++s2) {
                    const int rr = 2 * s2 + half;
                    const int code = sIdx[rr];
                    const int other = __shfl_xor_sync(0xffffffffu, code, 16);
                    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (code >= 0) g = *reinterpret_cast<const float4*>(sG + (c4 >> 3) * BXBLK + sw128_offset(rr, c4 & 7));
                    float4 o;
                    o.x = __shfl_xor_sync(0xffffffffu, g.x, 16); o.y = __shfl_xor_sync(0xffffffffu, g.y, 16);
                    o.z = __shfl_xor_sync(0xffffffffu, g.z, 16); o.w = __shfl_xor_sync(0xffffffffu, g.w, 16);
                    const bool same = code == other;
                    if (same) { g.x += o.x; g.y += o.y; g.z += o.z; g.w += o.w; }
                    if (code >= 0 && !(same && half == 1)) {
                        float4* acc = reinterpret_cast<float4*>(sAcc + code * D + 4 * c4);
                        float4 a = *acc;
                        a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
                        *acc = a;
                    }
                    __syncwarp();
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(scat_done);
        }
        {
            float4* dst = reinterpret_cast<float4*>(p.partial + (size_t)blockIdx.x * PARTIAL_FLOATS + 2 * PART_KD);
            for (int i = lane; i < PART_KD / 4; i += 32) dst[i] = reinterpret_cast<const float4*>(sAcc)[i];
        }
    } else if (warp >= 4) {
        // =============================== row threads ======================================================
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;                              // row within the tile == TMEM lane
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        const float tau = l2 ? fmaxf(__ldg(p.temp), 0.f) : 1.f;
        const float cmul = l2 ? -tau : 1.f;
        int real_tiles = 0;
        uint32_t it = 0;
        int tl_n = 0;
        VQB_BTL(1);
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            const uint32_t ph = it & 1;
            const int row0 = tile * BBM;
            const int rows = min(BBM, p.N - row0);
            const bool valid = r < rows;
            const bool real = tile_is_real(tile);
            real_tiles += real;
            VQB_BTL(2);
            mbar_wait(in_full, ph);
            VQB_BTL(3);
            float* stP = reinterpret_cast<float*>(sC);
            float* stG = reinterpret_cast<float*>(sCmn);
            const int nfl = rows * K, nbulk = ((nfl * 4) & ~15) >> 2;
            if (nbulk != nfl) {                                    // last < 16 bytes of a ragged tile
                if (r < nfl - nbulk) {
                    stP[nbulk + r] = p.p[(size_t)row0 * K + nbulk + r];
                    stG[nbulk + r] = p.gp[(size_t)row0 * K + nbulk + r];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");
            }
            // ---- softmax backward for row r (all loads of the row first, then the arithmetic) ------------------
            float c[64], gg[64];
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                const bool on = valid && k < K;
                c[k] = on ? stP[r * K + k] : 0.f;
                gg[k] = on ? stG[r * K + k] : 0.f;
            }
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 64; ++k) s4[k & 3] = fmaf(gg[k], c[k], s4[k & 3]);
            const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            float r4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 64; ++k) {
                c[k] = cmul * (c[k] * (gg[k] - s));
                r4[k & 3] += c[k];
            }
            const float rsum = (r4[0] + r4[1]) + (r4[2] + r4[3]);
            VQB_BTL(4);
            asm volatile("bar.sync 1, 128;" ::: "memory");        // staging fully consumed: C may overwrite it
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const int k0 = kb * 32 + ch * 4;
                    const float4 v = make_float4(c[k0], c[k0 + 1], c[k0 + 2], c[k0 + 3]);
                    *reinterpret_cast<float4*>(sC + kb * BXBLK + sw128_offset(r, ch)) = v;
                    *reinterpret_cast<float4*>(sCmn + kb * BXBLK + sw32b_offset(r, ch)) = v;
                }
            }
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                float4 xv[8];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) xv[ch] = *reinterpret_cast<const float4*>(sX + kb * BXBLK + sw32b_offset(r, ch));
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    *reinterpret_cast<float4*>(sXlo + kb * BXBLK + sw32b_offset(r, ch)) =
                        make_float4(tf32_lo(xv[ch].x), tf32_lo(xv[ch].y), tf32_lo(xv[ch].z), tf32_lo(xv[ch].w));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(coef_ready);
            VQB_BTL(5);
            // second pass: the tf32 remainders of C through the same two buffers, once pass 0 has been consumed
            mbar_wait(p1_done, ph);
            VQB_BTL(6);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const int k0 = kb * 32 + ch * 4;
                    const float4 v = make_float4(tf32_lo(c[k0]), tf32_lo(c[k0 + 1]), tf32_lo(c[k0 + 2]), tf32_lo(c[k0 + 3]));
                    *reinterpret_cast<float4*>(sC + kb * BXBLK + sw128_offset(r, ch)) = v;
                    *reinterpret_cast<float4*>(sCmn + kb * BXBLK + sw32b_offset(r, ch)) = v;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(coef2_ready);
            VQB_BTL(7);

            // ---- dx = g_q + 2 x rowsum(C) - 2 (C @ E)   |   C @ W -----------------------------------------
            mbar_wait(mma_done, ph);
            tcgen05_fence_after();
            VQB_BTL(8);
            const float r2 = 2.f * rsum;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
                float a[32], b[32];
                tmem_ld_32x32(d1 + lane_addr + kb * 32, a);
                tmem_ld_32x32(d1 + lane_addr + D + kb * 32, b);
                float4 xv[8], gv[8];
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {                   // all loads first: the stores below may alias
                    xv[ch] = *reinterpret_cast<const float4*>(sX + kb * BXBLK + sw32b_offset(r, ch));
                    gv[ch] = p.gq ? *reinterpret_cast<const float4*>(sG + kb * BXBLK + sw128_offset(r, ch))
                                  : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    float4 o;
                    o.x = a[4 * ch] + b[4 * ch]; o.y = a[4 * ch + 1] + b[4 * ch + 1];
                    o.z = a[4 * ch + 2] + b[4 * ch + 2]; o.w = a[4 * ch + 3] + b[4 * ch + 3];
                    if (l2) {
                        o.x = fmaf(xv[ch].x, r2, gv[ch].x) - 2.f * o.x; o.y = fmaf(xv[ch].y, r2, gv[ch].y) - 2.f * o.y;
                        o.z = fmaf(xv[ch].z, r2, gv[ch].z) - 2.f * o.z; o.w = fmaf(xv[ch].w, r2, gv[ch].w) - 2.f * o.w;
                    }
                    *reinterpret_cast<float4*>(sXlo + kb * BXBLK + sw32b_offset(r, ch)) = o;   // x_lo is dead after mma_done
                }
            }
            tcgen05_fence_before();
            VQB_BTL(9);
            asm volatile("bar.sync 1, 128;" ::: "memory");
            constexpr int D4 = D / 4;
            for (int i = r; i < rows * D4; i += 128) {
                const int rr = i / D4, cc = i % D4;
                stg4_stream(p.dx + (size_t)(row0 + rr) * D + 4 * cc,
                            *reinterpret_cast<const float4*>(sXlo + (cc >> 3) * BXBLK + sw32b_offset(rr, cc & 7)));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(tile_free);
            VQB_BTL(10);
        }
        // ---- once per CTA: the K x D sums and the column sums held in TMEM -> this CTA's partial record -------
        {
            float* part = p.partial + (size_t)blockIdx.x * PARTIAL_FLOATS;
            tcgen05_fence_after();
            const float scale = l2 ? -2.f : 1.f;
            const int d = r & (D - 1);
            float* dst = part + (r >= D ? PART_KD : 0);            // TMEM lanes D..2D-1 hold the x_lo part of the same d
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
                float a[32];
                if (real_tiles > 0) {
                    tmem_ld_32x32(d2 + lane_addr + hb * 32, a);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) a[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) dst[(hb * 32 + j) * D + d] = scale * a[j];
            }
            if (q4 == 0) {                                         // every lane of D3 holds the same column sums
#pragma unroll
                for (int hb = 0; hb < 2; ++hb) {
                    float a[32];
                    if (real_tiles > 0) {
                        tmem_ld_32x32(d3 + lane_addr + hb * 32, a);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) a[j] = 0.f;
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) part[3 * PART_KD + hb * 32 + j] = a[j];
                    }
                }
            }
        }
        VQB_BTL(11);
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// out[i] += sum over the CTAs' partial records, in a fixed order (deterministic).  grid.x covers the K*D elements
// (+ one block for the column sums), blockDim = (64, 16): 16 slices of the CTA list per element (independent
// accumulators, loads unrolled for memory-level parallelism), combined through shared memory.
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const float* __restrict__ partial, int n_cta, int K, int l2, float* __restrict__ dW,
                       float* __restrict__ dG, float* __restrict__ colsum) {
    __shared__ float red[16][64];
    __shared__ float red2[16][64];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int n_kd = K * 64;
    const bool cs_block = (int)blockIdx.x * 64 >= n_kd;            // last block: column sums
    const int i = cs_block ? tx : blockIdx.x * 64 + tx;
    float a0 = 0.f, a1 = 0.f, b0 = 0.f;
    if (cs_block) {
#pragma unroll 4
        for (int cta = ty; cta < n_cta; cta += 16) a0 += __ldg(partial + (size_t)cta * PARTIAL_FLOATS + 3 * PART_KD + tx);
    } else if (i < n_kd) {
#pragma unroll 4
        for (int cta = ty; cta < n_cta; cta += 16) {
            const float* rec = partial + (size_t)cta * PARTIAL_FLOATS + i;
            a0 += __ldg(rec);
            a1 += __ldg(rec + PART_KD);
            b0 += __ldg(rec + 2 * PART_KD);
        }
    }
    red[ty][tx] = a0 + a1; red2[ty][tx] = b0;
    __syncthreads();
    if (ty == 0) {
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) { sa += red[j][tx]; sb += red2[j][tx]; }
        if (cs_block) {
            if (tx < K) colsum[tx] += sa;
        } else if (i < n_kd) {
            if (l2) { dW[i] += sa + sb; } else { dW[i] += sa; if (dG) dG[i] += sb; }
        }
    }
}

// -----------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------
unsigned long long* get_debug_timeline();
void launch_build_operands(const float* w, const float* bias, int K, int Kpad, int D, float scale, float pad_bias,
                           float* hi, float* lo, float* emax, cudaStream_t s);

static size_t b_align256(size_t v) { return (v + 255) & ~(size_t)255; }
static size_t b_hi_bytes(int64_t K, int64_t D) { return cache_hi_bytes(K, D); }
static size_t b_lo_bytes(int64_t K, int64_t D) { return cache_lo_bytes(K, D); }

bool backward_tensor_supported(const vqb_bwd_args* a) {
    if (!(a->flags & VQB_TENSOR_CORES)) return false;
    if (!a->g_p || !(a->flags & VQB_STOP_GRAD) || (a->flags & VQB_TEMP_GRAD)) return false;
    if (a->n_codes > 64 || a->dim != 64) return false;
    const int64_t nr = a->n_real_rows;
    if (nr > 0 && nr < a->n_rows && nr % BBM != 0) return false;   // the real/fake split must fall on a tile edge
    return aligned16(a->p_code) && aligned16(a->g_p);
}

int backward_tensor_workspace(const vqb_bwd_args* a, size_t* bytes) {
    const bool cached = a->operand_cache && (a->flags & VQB_SCORE_L2);
    *bytes = backward_tensor_supported(a) ? (cached ? 0 : b_hi_bytes(a->n_codes, a->dim) + b_lo_bytes(a->n_codes, a->dim)) + 256 +
                                                 (size_t)sm_count() * PARTIAL_FLOATS * 4 : 0;
    return VQB_OK;
}

int launch_backward_tensor(const vqb_bwd_args* a, cudaStream_t s) {
    const int64_t N = a->n_rows, K = a->n_codes, D = a->dim;
    const bool cached = a->operand_cache && (a->flags & VQB_SCORE_L2);
    const size_t op_bytes = cached ? 0 : b_hi_bytes(K, D) + b_lo_bytes(K, D);
    const size_t need = op_bytes + 256 + (size_t)sm_count() * PARTIAL_FLOATS * 4;
    if (!a->workspace || a->workspace_bytes < need) {
        set_error("vqb_backward: workspace too small (%zu < %zu bytes)", a->workspace_bytes, need);
        return VQB_ERR_WORKSPACE;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
    float* hi;
    float* lo;
    if (cached) {     // [fwd_hi | fwd_lo | bwd_hi | bwd_lo]
        uint8_t* oc = reinterpret_cast<uint8_t*>(const_cast<void*>(a->operand_cache)) + b_hi_bytes(K, D) + b_lo_bytes(K, D);
        hi = reinterpret_cast<float*>(oc);
        lo = reinterpret_cast<float*>(oc + b_hi_bytes(K, D));
    } else {
        hi = reinterpret_cast<float*>(ws);
        lo = reinterpret_cast<float*>(ws + b_hi_bytes(K, D));
        launch_build_operands(a->score_w, nullptr, (int)K, (int)K, (int)D, 1.f, 0.f, hi, lo, nullptr, s);
        VQB_CHECK_LAUNCH("build_operands_kernel");
    }

    CUtensorMap tx, tg, th, tl;
    int rc;
    if ((rc = make_tmap_2d_f32(&tx, a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, BBM, true))) return rc;
    if ((rc = make_tmap_2d_f32(&tg, a->g_q ? a->g_q : a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, BBM))) return rc;
    if ((rc = make_tmap_2d_f32(&th, hi, (uint64_t)K, (uint64_t)(D + 32), (uint64_t)(D + 32), 64, true))) return rc;
    if ((rc = make_tmap_2d_f32(&tl, lo, (uint64_t)K, (uint64_t)D, (uint64_t)D, 64, true))) return rc;

    BwdTcP p;
    p.p = a->p_code; p.gp = a->g_p; p.gq = a->g_q; p.idx = (const long long*)a->idx; p.temp = a->temp;
    p.dx = a->dx; p.dW = a->d_score_w; p.colsum = a->colsum; p.dG = (a->flags & VQB_SCORE_L2) ? nullptr : a->d_gather;
    p.N = (int)N; p.K = (int)K; p.n_real = (int)(a->n_real_rows > 0 ? a->n_real_rows : 0);
    p.num_tiles = (int)ceil_div(N, BBM);
    p.flags = a->flags;
    p.dbg = get_debug_timeline();
    p.partial = reinterpret_cast<float*>(ws + op_bytes + 256);

    constexpr int KB = 2;
    const size_t smem = (size_t)3 * KB * BXBLK + 4 * BXBLK + 2 * KB * BEBLK + 64 * 64 * 4 + BXBLK + BBM * 4 + 256 + 1024;
    auto kern = vqb_bwd_tc_kernel<KB>;
    VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    kernel_event_begin(s);
    kern<<<grid, BWD_THREADS, smem, s>>>(tx, tg, th, tl, p);
    kernel_event_end(s);
    VQB_CHECK_LAUNCH("vqb_bwd_tc_kernel");
    const int n_blocks = (int)ceil_div(K * 64, 64) + 1;
    reduce_partials_kernel<<<n_blocks, dim3(64, 16), 0, s>>>(p.partial, grid, (int)K, (a->flags & VQB_SCORE_L2) ? 1 : 0,
                                                           a->d_score_w, p.dG, a->colsum);
    VQB_CHECK_LAUNCH("reduce_partials_kernel");
    return VQB_OK;
}

}  // namespace vqb
