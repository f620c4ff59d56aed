// Parity-mode backward of the quantizer (K <= 64, D = 64, stop_grad) on tcgen05 / TMEM / TMA, third generation.
//
// Autograd of src/embed.py:105-147 / :187-205 (entered from src/solver.py:144); algebra in DESIGN.md:
//   Gs = P * (g_p - rowsum(g_p * P));   C = -tau Gs (L2)  |  Gs (LINEAR)
//   dx  = g_q + 2 x rowsum(C) - 2 C @ E          (L2)     |  C @ W           (LINEAR)
//   dE += -2 C*^T @ x + scatter_add(idx, g_q)     (L2)     |  dW += C^T @ x ; dT += scatter_add(idx, g_q)
//   colsum += colsum(C*)                          (C* = rows below n_real_rows, first_n_real_mel)
//
// Shape of the kernel.  Like the forward (vqb_fwd_pc.cu) this path is HBM-bound and latency-shaped, so the design goal
// is rows in flight per SM: a CTA is 128 threads working on ONE tile of TR = 96 rows at a time -- thread = row = TMEM lane,
// no warp specialisation, a straight chain per tile -- and TWO CTAs are resident per SM (~100 KB of shared memory each),
// each a persistent loop over its tiles.  While one CTA waits for its 82 KB of inputs the other one computes.
//
// Everything that is a sum over rows or over codes runs on the tensor cores as kind::f16 MMAs over fp16x2 operands
// (vqb_f16x2.cuh: v * 2^s = hi + lo, 22 significant bits, exact power-of-two scales), ONE burst per tile:
//   GEMM 1  D1[r][d]  = sum_k C'[r][k] E'[k][d]          -> dx      (A = C' K-major, B = the table image, MN-major)
//   GEMM 2  D2[d][k]  = sum_r x''[r][d] C'[r][k]          -> dE/dW   (A = [x_hi ; x_lo] stacked along M, MN-major)
//   GEMM 3  D3[d][k]  = sum_r gq''[r][d] OH[r][k]         -> the index-keyed scatter-add of g_q as a GEMM with the
//                                                           one-hot matrix of the picked codes (exact products)
// For 16-bit types the K-major and MN-major 128-byte-swizzle layouts of a [rows][64] tile are the same bytes, so the
// coefficient tile C' is written once and serves GEMM 1 (as A) and GEMM 2 (as B); x and g_q are split in place over
// their raw TMA tiles.  Scales:  C'[r] = C[r] 2^-e_r (row maximum in [2^14, 2^15)),  x''[r] = x[r] 2^(e_r - t),
// gq'' = gq 2^-g  with  t, g  per tile;  C @ E = D1 2^(e_r + gE),  x^T C = D2 2^t,  scatter = D3 2^g.
// D2 / D3 are folded into registers once per tile (the scales differ from tile to tile); the K x D sums leave the CTA once,
// as plain stores into a per-CTA partial record that the tail / reduce kernel (vqb_bwd_tail.cu) adds in a fixed order --
// no atomics, gradients are bit-reproducible.
#include <cudaTypedefs.h>
#include <limits.h>
#include <math.h>
#include "vqb_common.cuh"
#include "vqb_tc.cuh"
#include "vqb_f16x2.cuh"

namespace vqb {
using namespace tc;

constexpr int TR = 96;                  // rows per tile (a multiple of 16: GEMM 2 / 3 contract over rows, 16 per MMA)
constexpr int TBLK = TR * 128;          // one [TR rows][128 B] block = 12 KB
constexpr int BP_THREADS = 128;
constexpr int BP_KD = 64 * 64;
constexpr int BP_PARTIAL_FLOATS = 2 * BP_KD + 64;   // same record as vqb_bwd_tail.cu: [0] d_score_w part, [1] scatter part (LINEAR) /
                                                    // transposed projected columns (L2 + fused tail), [2] column sums

struct BwdPcP {
    const float* p;
    const float* gp;
    const float* gq;          // may be NULL
    const long long* idx;
    const float* temp;
    const uint8_t* img;       // operand image of the score table (L2: the codebook; LINEAR: W)
    float* partial;           // [grid][BP_PARTIAL_FLOATS]
    unsigned long long* dbg;  // optional timeline buffer (developer hook)
    int N, K, n_real, num_tiles;
    int t_first;              // L2 + fused tail: columns d >= t_first are also stored transposed (record plane 1), else 64
    int pg_bytes, stage_bytes;
    const float* glogp;       // [S][B][K] upstream gradient of log(p_code + eps), or NULL (then gp is given)
    float eps;
    int gl_rb;                // glogp by TMA: rows per box (8 or 32; needs S % gl_rb == 0); 0 = staged with cp.async
    int gl_ld;                // row stride of the staged g_logp rows in shared memory, floats (TMA: box width, else K)
    const long long* lens;    // [N / S] valid frames per utterance, or NULL (length-aware rows, vqb_bwd_args.row_lengths)
    int S;
    float* dx;                // [N][64] (pad-only tiles are zero-filled directly)
    unsigned flags;
};

#define VQB_BTL(tag) do { if (p.dbg && r == 0 && blockIdx.x == 0 && tl_n < 60) { p.dbg[tl_n++] = ((unsigned long long)(tag) << 56) | (globaltimer_ns() & 0x00FFFFFFFFFFFFFFull); } } while (0)

template <int KP, bool L2>
__global__ void __launch_bounds__(BP_THREADS, 2)
vqb_bwd_pcode_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g,
                     const __grid_constant__ CUtensorMap tm_dx, const __grid_constant__ CUtensorMap tm_gl, BwdPcP p) {
    constexpr int D = 64;
    constexpr int TILE = 2 * TBLK;                   // one [TR][64] fp32 tile = two blocks
    constexpr int EV = KP / 8;                       // 16-byte words per thread that cover the table image (2 * KP * 128 B)

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sPG = smem;                             // p_code / g_p staging, then C_hi | C_lo | E_hi | E_lo
    float* stP = reinterpret_cast<float*>(sPG);
    float* stG = reinterpret_cast<float*>(sPG + p.stage_bytes);
    uint8_t* sCh = sPG;                              // [TR][128 B] C_hi  fp16 [rows][64 codes]
    uint8_t* sCl = sPG + TBLK;                       // C_lo
    uint8_t* sEh = sPG + 2 * TBLK;                   // [KP][128 B] E_hi  fp16 [codes][64 d]
    uint8_t* sEl = sEh + KP * 128;                   // E_lo
    uint8_t* sX = sPG + p.pg_bytes;                  // [2 blocks] raw x, then x_hi | x_lo (fp16) in place
    uint8_t* sG = sX + TILE;                         // [2 blocks] raw g_q, then gq_hi | gq_lo in place, then dx
    uint8_t* sOH = sG + TILE;                        // [TR][128 B] one-hot rows of the picked codes (fp16)
    int* sRed = reinterpret_cast<int*>(sOH + TBLK);  // [8]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 8);
    uint64_t* in_full = bars;
    uint64_t* d1_done = bars + 1;
    uint64_t* mma_done = bars + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int r = threadIdx.x, warp = r >> 5, lane = r & 31;
    const bool have_gq = p.gq != nullptr;
    const bool do_scatter = have_gq && !(L2 && (p.flags & VQB_SKIP));
    const int K = p.K;
    const int n_my = (p.num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA
    int tl_n = 0;
    if (p.dbg && r == 0) p.dbg[128 + 2 * blockIdx.x] = globaltimer_ns();

    if (r == 0) {
        tma_prefetch_desc(&tm_x); tma_prefetch_desc(&tm_g); tma_prefetch_desc(&tm_dx);
        if (p.gl_rb) tma_prefetch_desc(&tm_gl);
        mbar_init(in_full, 1); mbar_init(d1_done, 1); mbar_init(mma_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<256>(tmem_slot);
    // PDL: the fused tail kernel (launched behind this one with programmatic serialization) may queue up now; it waits
    // for this grid to complete before it reads the partial records.
    pdl_launch();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t d1 = tmem_base, d2 = tmem_base + 64, d3 = tmem_base + 64 + KP;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;

    const float tau = L2 ? fmaxf(__ldg(p.temp), 0.f) : 1.f;
    const float cmul = L2 ? -tau : 1.f;
    const int gE = __ldg(reinterpret_cast<const int*>(p.img + IMG_HDR));
    const float uE = pow2i(gE);

    float acc[KP];                                   // lane r = d (r < 64: hi part of x / g_q, else lo part): dE (L2) | dW (LINEAR)
    float accg[L2 ? 1 : KP];                         // LINEAR: the scatter sums, kept apart (they go to the gather table)
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = 0.f;
#pragma unroll
    for (int k = 0; k < (L2 ? 1 : KP); ++k) accg[k] = 0.f;
    float cs0 = 0.f, cs1 = 0.f;                      // two column sums per lane (see the butterfly below)

    // g_logp [S][B][K] seen as a 2-D tensor [S][B * K]: the rows of a tile that belong to utterance b sit in the box
    // (columns (b * K & ~3) .. + gl_ld, frames sf .. + gl_rb) -- TMA wants the box to start on a 16-byte boundary, so it
    // starts up to three words early and the row thread skips them; boxes never straddle an utterance
    // (S % gl_rb == 0 == TR % gl_rb)
    const uint32_t gl_box_bytes = (uint32_t)(p.gl_rb * p.gl_ld * 4);
    auto glogp_bytes = [&](int rows) -> uint32_t { return p.gl_rb ? (uint32_t)((rows + p.gl_rb - 1) / p.gl_rb) * gl_box_bytes : 0u; };
    auto issue_glogp = [&](int row0, int rows) {
        for (int j = 0; j * p.gl_rb < rows; ++j) {
            const int g0 = row0 + j * p.gl_rb, b = g0 / p.S;
            tma_load_2d(reinterpret_cast<uint8_t*>(stG) + j * gl_box_bytes, &tm_gl, (b * K) & ~3, g0 - b * p.S, in_full);
        }
    };
    auto issue_loads = [&](int it) {
        const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * TR;
        const int rows = min(TR, p.N - row0);
        const uint32_t bulk = (uint32_t)(rows * K * 4) & ~15u;
        mbar_arrive_expect_tx(in_full, (p.gp ? 2 : 1) * bulk + TILE + (have_gq ? TILE : 0) + glogp_bytes(rows));
        if (bulk) {
            bulk_load_1d(stP, p.p + (size_t)row0 * K, bulk, in_full);
            if (p.gp) bulk_load_1d(stG, p.gp + (size_t)row0 * K, bulk, in_full);
        }
        if (p.gl_rb) issue_glogp(row0, rows);
        tma_load_2d(sX, &tm_x, 0, row0, in_full);
        tma_load_2d(sX + TBLK, &tm_x, 32, row0, in_full);
        if (have_gq) {
            tma_load_2d(sG, &tm_g, 0, row0, in_full);
            tma_load_2d(sG + TBLK, &tm_g, 32, row0, in_full);
        }
    };
    // length-aware rows: does tile `it` of this CTA hold any real frame?  (a tile spans at most TR / S + 2 utterances)
    auto tile_live = [&](int it) -> bool {
        if (!p.lens) return true;
        const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * TR;
        const int rows = min(TR, p.N - row0);
        for (int b = row0 / p.S; b * p.S < row0 + rows; ++b) {
            const long long lo = max(row0, b * p.S), hi = min((long long)(row0 + rows), (long long)b * p.S + __ldg(p.lens + b));
            if (hi > lo) return true;
        }
        return false;
    };
    auto next_live = [&](int it) -> int {           // first live tile of this CTA at or after `it` (n_my if none)
        while (it < n_my && !tile_live(it)) ++it;
        return it;
    };
    {
        const int first = next_live(0);
        if (r == 0 && first < n_my) issue_loads(first);
    }
    VQB_BTL(1);

    int lt = 0;                                      // live tiles processed so far (drives the mbarrier phases)
    for (int it = 0; it < n_my; ++it) {
        const int row0 = ((int)blockIdx.x + it * (int)gridDim.x) * TR;
        const int rows = min(TR, p.N - row0);
        if (!tile_live(it)) {
            // only padding: nothing is loaded or computed; the dx rows are zero
            float4* d4 = reinterpret_cast<float4*>(p.dx + (size_t)row0 * D);
            for (int i = r; i < rows * (D / 4); i += BP_THREADS) d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            continue;
        }
        const uint32_t ph = lt & 1;
        bool valid = r < rows;
        if (p.lens && valid) {
            const int b = (row0 + r) / p.S;
            valid = (row0 + r) - b * p.S < __ldg(p.lens + b);       // a pad row inside a live tile contributes nothing
        }
        const bool real = valid && (p.n_real <= 0 || row0 + r < p.n_real);
        // the table image, on its way to shared memory through registers (it lands behind the C tiles once the staging
        // area has been consumed)
        uint4 ev[EV];
#pragma unroll
        for (int i = 0; i < EV; ++i) {
            const int i4 = r + BP_THREADS * i;                      // < 2 * KP * 8
            const int piece = i4 / (KP * 8), off = (i4 - piece * KP * 8) * 16;
            ev[i] = __ldg(reinterpret_cast<const uint4*>(p.img + piece * IMG_PIECE + off));
        }
        long long code = 0;
        if (do_scatter && valid) code = __ldg(p.idx + row0 + r);
        const bool gl_staged = p.glogp && !p.gl_rb;
        if (gl_staged) {
            // Gradient of the CTC input folded in, general shapes (the boxes of the TMA route need S % 8 == 0 and 16-byte
            // aligned rows of [S][B * K]): the rows of g_logp [S][B][K] that belong to this tile are staged where the bulk
            // copy would have put g_p, one row per warp step, the lanes along its K codes (contiguous 4 K-byte runs), with
            // 4-byte cp.async: no registers held, every element of the tile in flight at once.
            const int B_ = p.N / p.S;
            int b = (row0 + warp) / p.S, sf = (row0 + warp) - b * p.S;
            for (int rl = warp; rl < rows; rl += BP_THREADS / 32) {
                const float* grow = p.glogp + ((size_t)sf * B_ + b) * K;
                const uint32_t dst = smem_u32(stG + rl * K);
                for (int k = lane; k < K; k += 32)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4u * k), "l"(grow + k) : "memory");
                sf += BP_THREADS / 32;
                while (sf >= p.S) { sf -= p.S; ++b; }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        mbar_wait(in_full, ph);
        VQB_BTL(2);
        {
            const int nfl = rows * K, nbulk = ((nfl * 4) & ~15) >> 2;
            if (nbulk != nfl || gl_staged) {                       // last < 16 bytes of a ragged tile; the staged g_logp rows
                if (nbulk != nfl && r < nfl - nbulk) {
                    stP[nbulk + r] = p.p[(size_t)row0 * K + nbulk + r];
                    if (p.gp) stG[nbulk + r] = p.gp[(size_t)row0 * K + nbulk + r];
                }
                __syncthreads();
            }
        }
        // ---- softmax backward for row r -------------------------------------------------------------------------
        float c[64];
        float rsum, m;
        {
            float gg[KP];
#pragma unroll
            for (int k = 0; k < KP; ++k) { c[k] = 0.f; gg[k] = 0.f; }
            if (valid) {
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    if (k < KP - 15 || k < K) c[k] = stP[r * K + k];
                }
                if (!p.glogp) {
#pragma unroll
                    for (int k = 0; k < KP; ++k) {
                        if (k < KP - 15 || k < K) gg[k] = stG[r * K + k];
                    }
                } else {
                    if (p.gl_rb) {
                        // rows of gl_ld floats (gl_ld / 4 odd: conflict-free 128-bit reads); the row's K words start `sh`
                        // words into it (see issue_glogp), what lies beyond them belongs to the next utterance
                        const float4* g4 = reinterpret_cast<const float4*>(stG + r * p.gl_ld);
                        const int sh = (((row0 + r) / p.S) * K) & 3;
                        const bool s1 = sh == 1, s2 = sh == 2, s3 = sh == 3;
                        float4 a = g4[0];
#pragma unroll
                        for (int k4 = 0; k4 < KP / 4; ++k4) {
                            if (4 * k4 < KP - 15 || 4 * k4 < K) {
                                const float4 n = g4[k4 + 1];
                                gg[4 * k4] = s3 ? a.w : (s2 ? a.z : (s1 ? a.y : a.x));
                                gg[4 * k4 + 1] = s3 ? n.x : (s2 ? a.w : (s1 ? a.z : a.y));
                                gg[4 * k4 + 2] = s3 ? n.y : (s2 ? n.x : (s1 ? a.w : a.z));
                                gg[4 * k4 + 3] = s3 ? n.z : (s2 ? n.y : (s1 ? n.x : a.w));
                                a = n;
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < KP; ++k) {
                            if (k < KP - 15 || k < K) gg[k] = stG[r * K + k];
                        }
                    }
                    // d log(p + eps) / dp  (MUFU.RCP + multiply: 2 ulp)
#pragma unroll
                    for (int k = 0; k < KP; ++k) gg[k] = (k < KP - 15 || k < K) ? __fdividef(gg[k], c[k] + p.eps) : 0.f;
                }
            }
#pragma unroll
            for (int k = KP; k < 64; ++k) c[k] = 0.f;
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < KP; ++k) s4[k & 3] = fmaf(gg[k], c[k], s4[k & 3]);
            const float s = (s4[0] + s4[1]) + (s4[2] + s4[3]);
            float r4[4] = {0.f, 0.f, 0.f, 0.f}, m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                c[k] = cmul * (c[k] * (gg[k] - s));
                r4[k & 3] += c[k];
                m4[k & 3] = fmaxf(m4[k & 3], fabsf(c[k]));
            }
            rsum = (r4[0] + r4[1]) + (r4[2] + r4[3]);
            m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
        }
        __syncthreads();                                           // #1 the staging area has been consumed
        VQB_BTL(3);
        // ---- C' -> fp16 hi | lo over the staging area; the table image behind it -----------------------------------
        const bool nz = m > 0.f && m < INFINITY;                   // (a NaN / inf row scales by 1 and propagates)
        const int er = nz ? scale_exp(m) : 0;
        if (r < TR) {
            const float sc = pow2i(-er);
#pragma unroll
            for (int j = 0; j < KP / 8; ++j) {
                uint4 hi, lo;
                split8(c + 8 * j, sc, hi, lo);
                *reinterpret_cast<uint4*>(sCh + sw128_offset(r, j)) = hi;
                *reinterpret_cast<uint4*>(sCl + sw128_offset(r, j)) = lo;
            }
        }
#pragma unroll
        for (int i = 0; i < EV; ++i) {
            const int i4 = r + BP_THREADS * i;
            const int piece = i4 / (KP * 8), off = (i4 - piece * KP * 8) * 16;
            *reinterpret_cast<uint4*>(sEh + piece * KP * 128 + off) = ev[i];
        }
        // ---- column sums of C*: butterfly transpose-reduce over the warp ---------------------------------------------
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
        // ---- x row and g_q row -> registers; tile scales ----------------------------------------------------------------
        float xr[64], gr[64];
        float mx = 0.f, mg = 0.f;
        if (r < TR) {
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {
                    const float4 v = *reinterpret_cast<const float4*>(sX + kb * TBLK + sw128_offset(r, ch));
                    xr[kb * 32 + 4 * ch] = v.x; xr[kb * 32 + 4 * ch + 1] = v.y;
                    xr[kb * 32 + 4 * ch + 2] = v.z; xr[kb * 32 + 4 * ch + 3] = v.w;
                    mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
                }
            }
            if (have_gq) {
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                    for (int ch = 0; ch < 8; ++ch) {
                        const float4 v = *reinterpret_cast<const float4*>(sG + kb * TBLK + sw128_offset(r, ch));
                        gr[kb * 32 + 4 * ch] = v.x; gr[kb * 32 + 4 * ch + 1] = v.y;
                        gr[kb * 32 + 4 * ch + 2] = v.z; gr[kb * 32 + 4 * ch + 3] = v.w;
                        mg = fmaxf(fmaxf(mg, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
                    }
                }
            } else {
#pragma unroll
                for (int d = 0; d < 64; ++d) gr[d] = 0.f;
            }
        } else {
#pragma unroll
            for (int d = 0; d < 64; ++d) { xr[d] = 0.f; gr[d] = 0.f; }
        }
        const bool xok = real && nz && mx > 0.f && mx < INFINITY;
        const int tr = xok ? exp_of(mx) + er : INT_MIN / 2;
        const int tg = (valid && do_scatter && mg > 0.f && mg < INFINITY) ? exp_of(mg) : INT_MIN / 2;
        {
            const int wm = __reduce_max_sync(0xffffffffu, tr), wg = __reduce_max_sync(0xffffffffu, tg);
            if (lane == 0) { sRed[warp] = wm; sRed[4 + warp] = wg; }
        }
        __syncthreads();                                           // #2 tile maxima
        const int tmax = max(max(sRed[0], sRed[1]), max(sRed[2], sRed[3]));
        const int gmax = max(max(sRed[4], sRed[5]), max(sRed[6], sRed[7]));
        const int t = tmax > INT_MIN / 4 ? tmax - 14 : 0;
        const int g = gmax > INT_MIN / 4 ? gmax - 14 : 0;
        if (r < TR) {
            // x'' -> fp16 hi | lo, in place over the raw tile (this thread has read the whole of row r above)
            const float sx = (xok && er - t >= -126) ? pow2i(er - t) : 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                uint4 hi, lo;
                split8(xr + 8 * j, sx, hi, lo);
                *reinterpret_cast<uint4*>(sX + sw128_offset(r, j)) = hi;
                *reinterpret_cast<uint4*>(sX + TBLK + sw128_offset(r, j)) = lo;
            }
            if (do_scatter) {
                // gq'' likewise, and the one-hot row of the picked code (1.0 = 0x3C00 in its 16-bit slot)
                const float sg = (valid && g >= -126) ? pow2i(-g) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint4 hi, lo;
                    split8(gr + 8 * j, sg, hi, lo);
                    *reinterpret_cast<uint4*>(sG + sw128_offset(r, j)) = hi;
                    *reinterpret_cast<uint4*>(sG + TBLK + sw128_offset(r, j)) = lo;
                }
                const int kc = code < 0 ? 0 : (code >= K ? K - 1 : (int)code);
#pragma unroll
                for (int j = 0; j < KP / 8; ++j) {
                    uint4 oh = make_uint4(0u, 0u, 0u, 0u);
                    if (valid && (kc >> 3) == j) {
                        const uint32_t w = 0x3C00u << (16 * (kc & 1));
                        const int q = (kc & 7) >> 1;
                        oh.x = q == 0 ? w : 0u; oh.y = q == 1 ? w : 0u; oh.z = q == 2 ? w : 0u; oh.w = q == 3 ? w : 0u;
                    }
                    *reinterpret_cast<uint4*>(sOH + sw128_offset(r, j)) = oh;
                }
            }
            // the part of dx that needs no GEMM: g_q + 2 x rowsum(C)   (over the dead g_q registers)
            if (L2) {
                const float r2 = 2.f * rsum;
#pragma unroll
                for (int d = 0; d < 64; ++d) gr[d] = fmaf(xr[d], r2, gr[d]);
            }
        }
        fence_proxy_async_smem();                                  // generic writes -> tcgen05.mma operand reads
        tcgen05_fence_before();
        __syncthreads();                                           // #3 operands complete
        VQB_BTL(4);
        if (r == 0) {
            tcgen05_fence_after();
            constexpr uint32_t IDESC1 = umma_idesc(0u, 128, D) | UMMA_B_MN;
            constexpr uint32_t IDESC2 = umma_idesc(0u, 128, KP) | UMMA_A_MN | UMMA_B_MN;
            // GEMM 1: D1[r][d] = sum_k C'[r][k] E'[k][d]   (hi.hi + lo.hi + hi.lo), 16 codes per K-step
#pragma unroll
            for (int ks = 0; ks < KP / 16; ++ks) {
                const uint64_t ah = umma_desc_sw128(sCh) + 2 * ks, al = umma_desc_sw128(sCl) + 2 * ks;
                const uint64_t bh = umma_desc_sw128_mn(sEh + ks * 2048, 8192, 1024);
                const uint64_t bl = umma_desc_sw128_mn(sEl + ks * 2048, 8192, 1024);
                umma_bf16(d1, ah, bh, IDESC1, ks != 0);
                umma_bf16(d1, al, bh, IDESC1, true);
                umma_bf16(d1, ah, bl, IDESC1, true);
            }
            umma_commit(d1_done);
            // GEMM 2: D2[d][k] = sum_r x''[r][d] C'[r][k]: TMEM lanes 0..63 take x_hi, lanes 64..127 x_lo; 16 rows per K-step
#pragma unroll
            for (int ks = 0; ks < TR / 16; ++ks) {
                const uint64_t a = umma_desc_sw128_mn(sX + ks * 2048, TBLK, 1024);
                umma_bf16(d2, a, umma_desc_sw128_mn(sCh + ks * 2048, TBLK, 1024), IDESC2, ks != 0);
                umma_bf16(d2, a, umma_desc_sw128_mn(sCl + ks * 2048, TBLK, 1024), IDESC2, true);
            }
            // GEMM 3: D3[d][k] = sum_r gq''[r][d] OH[r][k]
            if (do_scatter) {
#pragma unroll
                for (int ks = 0; ks < TR / 16; ++ks)
                    umma_bf16(d3, umma_desc_sw128_mn(sG + ks * 2048, TBLK, 1024), umma_desc_sw128_mn(sOH + ks * 2048, TBLK, 1024),
                              IDESC2, ks != 0);
            }
            umma_commit(mma_done);
        }
        // ---- dx = g_q + 2 x rowsum(C) - 2 (C @ E)   |   C @ W ---------------------------------------------------
        mbar_wait(d1_done, ph);
        tcgen05_fence_after();
        VQB_BTL(5);
        if (r < TR) {
            const float u = pow2i(er) * uE * (L2 ? -2.f : 1.f);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                float a[32];
                tmem_ld_32x32(d1 + lane_addr + kb * 32, a);
#pragma unroll
                for (int j = 0; j < 32; ++j) gr[kb * 32 + j] = fmaf(a[j], u, L2 ? gr[kb * 32 + j] : 0.f);
            }
        }
        // ---- all MMAs retired: dx -> the (dead) g_q tile, D2 / D3 of this tile -> registers ---------------------------
        mbar_wait(mma_done, ph);
        tcgen05_fence_after();
        VQB_BTL(6);
        if (r < TR) {
            if (p.lens && !valid) {
#pragma unroll
                for (int d = 0; d < 64; ++d) gr[d] = 0.f;          // pad row: dx = 0 whatever g_q holds there
            }
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int ch = 0; ch < 8; ++ch)
                    *reinterpret_cast<float4*>(sG + kb * TBLK + sw128_offset(r, ch)) =
                        make_float4(gr[kb * 32 + 4 * ch], gr[kb * 32 + 4 * ch + 1], gr[kb * 32 + 4 * ch + 2], gr[kb * 32 + 4 * ch + 3]);
        }
        fence_proxy_async_smem();
        {
            // (t and g are exponents of ordinary magnitudes; the clamp of pow2i only matters beyond 2^+-126)
            const float ts = pow2i(t) * (L2 ? -2.f : 1.f);
            float a[KP];
            tmem_ld_cols<KP>(d2 + lane_addr, a);
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = fmaf(a[k], ts, acc[k]);
            if (do_scatter) {
                const float gs = pow2i(g);
                tmem_ld_cols<KP>(d3 + lane_addr, a);
#pragma unroll
                for (int k = 0; k < KP; ++k) {
                    if (L2) acc[k] = fmaf(a[k], gs, acc[k]);
                    else accg[k] = fmaf(a[k], gs, accg[k]);
                }
            }
        }
        tcgen05_fence_before();
        __syncthreads();                                           // #4 dx staged; TMEM and every operand tile are free
        VQB_BTL(7);
        if (r == 0) {
            tma_store_2d(&tm_dx, sG, 0, row0);
            tma_store_2d(&tm_dx, sG + TBLK, 32, row0);
            tma_store_commit();
            const int nit = next_live(it + 1);
            if (nit < n_my) {
                // next tile: p_code / g_p / x may land now; g_q only once dx has left its tile
                const int nrow0 = ((int)blockIdx.x + nit * (int)gridDim.x) * TR;
                const int nrows = min(TR, p.N - nrow0);
                const uint32_t bulk = (uint32_t)(nrows * K * 4) & ~15u;
                mbar_arrive_expect_tx(in_full, (p.gp ? 2 : 1) * bulk + TILE + (have_gq ? TILE : 0) + glogp_bytes(nrows));
                if (bulk) {
                    bulk_load_1d(stP, p.p + (size_t)nrow0 * K, bulk, in_full);
                    if (p.gp) bulk_load_1d(stG, p.gp + (size_t)nrow0 * K, bulk, in_full);
                }
                if (p.gl_rb) issue_glogp(nrow0, nrows);
                tma_load_2d(sX, &tm_x, 0, nrow0, in_full);
                tma_load_2d(sX + TBLK, &tm_x, 32, nrow0, in_full);
                tma_store_wait_read();
                if (have_gq) {
                    tma_load_2d(sG, &tm_g, 0, nrow0, in_full);
                    tma_load_2d(sG + TBLK, &tm_g, 32, nrow0, in_full);
                }
            } else {
                tma_store_wait_read();
            }
        }
        VQB_BTL(8);
        ++lt;
    }
    __syncthreads();                                               // dx of the last tile has left shared memory (thread 0 waited)
    VQB_BTL(9);

    // ---- once per CTA: the K x D sums -> this CTA's partial record -----------------------------------------------
    float* part = p.partial + (size_t)blockIdx.x * BP_PARTIAL_FLOATS;
    float* sXch = reinterpret_cast<float*>(sX);                    // [KP][64] lo halves (dE / dW), then [KP][64] (scatter, LINEAR)
    float* sCs = reinterpret_cast<float*>(sPG);                    // [4][64] column sums
    if (r >= 64) {
#pragma unroll
        for (int k = 0; k < KP; ++k) sXch[k * 64 + (r - 64)] = acc[k];
        if (!L2) {
#pragma unroll
            for (int k = 0; k < KP; ++k) sXch[KP * 64 + k * 64 + (r - 64)] = accg[k];
        }
    }
    sCs[warp * 64 + 2 * lane] = cs0;                               // the butterfly leaves columns 2L, 2L+1 in lane L
    sCs[warp * 64 + 2 * lane + 1] = cs1;
    __syncthreads();
    if (r < 64) {
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            if (k < K) {
                const float v = acc[k] + sXch[k * 64 + r];
                part[k * 64 + r] = v;
                if (L2) {
                    // projected columns once more, column-major: the tail's projection blocks read them coalesced
                    if (r >= p.t_first) part[BP_KD + r * 64 + k] = v;
                } else {
                    part[BP_KD + k * 64 + r] = accg[k] + sXch[KP * 64 + k * 64 + r];
                }
            }
        }
        part[2 * BP_KD + r] = (sCs[r] + sCs[64 + r]) + (sCs[128 + r] + sCs[192 + r]);
    }
    VQB_BTL(10);

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_base);
    if (p.dbg && r == 0) p.dbg[129 + 2 * blockIdx.x] = globaltimer_ns();
}

// -----------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------
unsigned long long* get_debug_timeline();

bool backward_pcode_supported(const vqb_bwd_args* a) {
    if (!(a->flags & VQB_TENSOR_CORES)) return false;
    if ((!a->g_p && !a->g_logp) || !(a->flags & VQB_STOP_GRAD) || (a->flags & VQB_TEMP_GRAD)) return false;
    if (a->n_codes > 64 || a->dim != 64) return false;
    return aligned16(a->p_code) && (!a->g_p || aligned16(a->g_p));
}

static int bp_grid(int64_t N) {
    const int64_t tiles = ceil_div(N, TR);
    const int slots = 2 * sm_count();
    return (int)(tiles < slots ? tiles : slots);
}

size_t backward_pcode_workspace(const vqb_bwd_args* a) {
    if (!backward_pcode_supported(a)) return 0;
    return (size_t)bp_grid(a->n_rows) * BP_PARTIAL_FLOATS * 4 + (a->operand_cache ? 0 : (size_t)IMG_BYTES);
}

template <int KP, bool L2>
static int launch_bp(const CUtensorMap& tx, const CUtensorMap& tg, const CUtensorMap& td, const CUtensorMap& tgl, BwdPcP p, int grid,
                     cudaStream_t s) {
    p.stage_bytes = (TR * p.K * 4 + 127) & ~127;
    const int ops = 2 * TBLK + 2 * KP * 128;
    const int stage2 = p.stage_bytes + (p.gl_rb ? TR * p.gl_ld * 4 : p.stage_bytes);    // p_code rows + (g_p | g_logp boxes)
    p.pg_bytes = ((stage2 > ops ? stage2 : ops) + 1023) & ~1023;
    const size_t smem = (size_t)p.pg_bytes + 2 * 2 * TBLK + TBLK + 8 * 4 + 3 * 8 + 16 + 1024;
    if ((int)smem > max_optin_smem()) return invalid("vqb_backward: the parity-mode kernel needs %zu B of shared memory", smem);
    auto kern = vqb_bwd_pcode_kernel<KP, L2>;
    { const int rc_ = ensure_smem(kern, smem, true); if (rc_) return rc_; }
    // a plain stream-ordered launch: what precedes the backward in the stream (the producer of g_p / g_q, possibly a copy)
    // is not ours to overlap -- and chaining it behind this library's own forward with programmatic serialization was
    // measured SLOWER (61.8 vs 57.9 us per step at config 2, profiles/r2_pdl_ab.txt).  The kernel still releases its own
    // successor early (the fused tail).
    kernel_event_begin(s);
    kern<<<grid, BP_THREADS, smem, s>>>(tx, tg, td, tgl, p);
    kernel_event_end(s);
    VQB_CHECK_LAUNCH("vqb_bwd_pcode_kernel");
    return VQB_OK;
}

int launch_backward_pcode(const vqb_bwd_args* a, cudaStream_t s) {
    const int64_t N = a->n_rows, K = a->n_codes, D = a->dim;
    const int grid = bp_grid(N);
    const size_t rec_bytes = (size_t)grid * BP_PARTIAL_FLOATS * 4;
    const bool cached = a->operand_cache != nullptr;
    const size_t need = rec_bytes + (cached ? 0 : (size_t)IMG_BYTES);
    if (!a->workspace || a->workspace_bytes < need) {
        set_error("vqb_backward: workspace too small (%zu < %zu bytes)", a->workspace_bytes, need);
        return VQB_ERR_WORKSPACE;
    }
    const bool l2 = (a->flags & VQB_SCORE_L2) != 0;
    const uint8_t* img = reinterpret_cast<const uint8_t*>(a->operand_cache);
    if (!cached) {
        uint8_t* w = reinterpret_cast<uint8_t*>(a->workspace) + rec_bytes;
        int rc = launch_build_image(a->score_w, a->score_b, (int)K, (int)D, w, s);
        if (rc) return rc;
        img = w;
    }
    CUtensorMap tx, tg, td;
    int rc;
    if ((rc = make_tmap_2d_f32(&tx, a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, TR))) return rc;
    if ((rc = make_tmap_2d_f32(&tg, a->g_q ? a->g_q : a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, TR))) return rc;
    if ((rc = make_tmap_2d_f32(&td, a->dx, (uint64_t)N, (uint64_t)D, (uint64_t)D, TR))) return rc;

    BwdPcP p;
    p.p = a->p_code; p.gp = a->g_p; p.gq = a->g_q; p.idx = (const long long*)a->idx; p.temp = a->temp; p.img = img;
    p.partial = reinterpret_cast<float*>(a->workspace);
    p.dbg = get_debug_timeline();
    p.N = (int)N; p.K = (int)K; p.n_real = (int)(a->n_real_rows > 0 && a->n_real_rows < N ? a->n_real_rows : 0);
    p.num_tiles = (int)ceil_div(N, TR);
    p.flags = a->flags;
    p.t_first = (a->tail && l2) ? 64 - (int)a->tail->dim_attr : 64;
    p.pg_bytes = p.stage_bytes = 0;
    p.lens = (const long long*)a->row_lengths; p.S = (int)a->frames_per_utt; p.dx = a->dx;
    p.glogp = a->g_logp; p.eps = a->ctc_eps;
    p.gl_rb = 0; p.gl_ld = (int)K;
    CUtensorMap tgl = tx;
    if (a->g_logp) {
        // the TMA route of g_logp: [S][B * K] with 16-byte aligned rows, boxes of 8 or 32 frames that never straddle an
        // utterance; any other shape is staged by the kernel itself (cp.async)
        const int64_t S = a->frames_per_utt, B = N / S;
        static const bool no_tma = getenv("VQB_GLOGP_NO_TMA") != nullptr;      // developer A/B
        if (!no_tma && S % 8 == 0 && (B * K) % 4 == 0 && aligned16(a->g_logp)) {
            p.gl_rb = S % 32 == 0 ? 32 : 8;
            int w4 = (int)((K + 3) / 4) + 1;               // K words that may start up to 3 words into the box ...
            w4 += (w4 & 1) ^ 1;                            // ... and an odd number of 16-byte words per row (bank spread)
            p.gl_ld = 4 * w4;
            if ((rc = make_tmap_2d_plain_f32(&tgl, a->g_logp, (uint64_t)(B * K), (uint64_t)S, (uint64_t)(B * K * 4),
                                             (uint32_t)p.gl_ld, (uint32_t)p.gl_rb))) return rc;
        }
    }
    const int KP = (int)((K + 15) / 16 * 16);
    if (l2) {
        switch (KP) {
            case 16: rc = launch_bp<16, true>(tx, tg, td, tgl, p, grid, s); break;
            case 32: rc = launch_bp<32, true>(tx, tg, td, tgl, p, grid, s); break;
            case 48: rc = launch_bp<48, true>(tx, tg, td, tgl, p, grid, s); break;
            default: rc = launch_bp<64, true>(tx, tg, td, tgl, p, grid, s); break;
        }
    } else {
        switch (KP) {
            case 16: rc = launch_bp<16, false>(tx, tg, td, tgl, p, grid, s); break;
            case 32: rc = launch_bp<32, false>(tx, tg, td, tgl, p, grid, s); break;
            case 48: rc = launch_bp<48, false>(tx, tg, td, tgl, p, grid, s); break;
            default: rc = launch_bp<64, false>(tx, tg, td, tgl, p, grid, s); break;
        }
    }
    if (rc) return rc;
    return launch_bwd_reduce(a, p.partial, grid, s, p.dbg);
}

}  // namespace vqb
