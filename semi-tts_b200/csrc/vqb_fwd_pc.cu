// Parity-mode forward of the quantizer (p_code is part of the result; K <= 64, D in {32, 64}) on tcgen05 / TMEM / TMA:
// the whole of L2Embedding.forward (src/embed.py:105-147) or SeperateEmbedding.forward (:187-205) in one kernel.
//
//   scores (neg_batch_l2 :208-213 with the temperature :115, or F.linear :190)  ->  softmax (:127)  ->  argmax over the
//   probabilities (:130)  ->  gather (:134 / :194-197)  ->  straight-through (:145)  ->  usage histogram.
//
// Shape of the kernel.  The path is HBM-bound with 5.5 kflop per row, so what matters is how many rows are in flight per
// SM, not the MMA rate: a CTA is 128 threads and ONE tile of 128 rows at a time, thread = row = TMEM lane, no warp
// specialisation -- a straight chain  TMA load -> split -> MMA -> softmax -> gather -> TMA store  per tile -- and THREE
// CTAs are resident per SM (67 KB of shared memory, <= 168 registers), so every SM works on 384 rows at once and the
// chain of one CTA is hidden behind the other two.  At BASELINE config 2 (51 200 rows = 400 tiles <= 444 CTA slots) the
// whole input is requested from HBM in the first microsecond of the kernel.
//
// Precision.  x and the score table are carried as fp16x2 (vqb_f16x2.cuh): the row thread rescales its row by a power of
// two and splits it IN PLACE over the raw TMA tile (x_hi | x_lo, 16 KB each), the table image comes ready-made from the
// table assembly; acc = x_lo.e_lo + x_lo.e_hi + x_hi.e_lo + x_hi.e_hi, 4 x D/16 MMAs of kind::f16 with fp32 accumulation
// in TMEM.  The distance is then formed in the reference's own association, (|x|^2 + |e|^2) - 2 x.e (:210-212).
// Index exactness.  |acc - exact fp32 dot| is bounded (see `win` below); a row whose two best scores lie within that
// window re-evaluates every code inside it in exact fp32 -- same expression and fmaf order as the CUDA-core kernel
// (vqb_fwd_simt.cu) -- before the softmax, so p_code, its argmax and the gathered codeword agree with the exact kernel.
#include <cudaTypedefs.h>
#include <math.h>
#include "vqb_common.cuh"
#include "vqb_tc.cuh"
#include "vqb_f16x2.cuh"

namespace vqb {
using namespace tc;

constexpr int PM = 128;                 // rows per tile (UMMA M)
constexpr int PBLK = PM * 128;          // one [128 rows][128 B] block = 16 KB

struct PcP {
    const float* x;            // [N][D] (exact re-rank only; the tiles arrive by TMA)
    const float* table;        // [K][D] fp32 score table (exact re-rank)
    const float* gtab;         // [K][D] fp32 gather table
    const float* bias;         // [K]    |e|^2 (L2) or b (LINEAR)
    const float* temp;         // [1]
    const uint8_t* img;        // operand image of the score table (vqb_f16x2.cuh)
    float* pcode;
    float* q;
    long long* idx;
    unsigned long long* hist;
    double* sqerr;
    unsigned int* stats;       // [0] += rows re-ranked in exact fp32 (may be NULL)
    unsigned long long* dbg;   // optional timeline buffer (vqb_debug_set_timeline), NULL in production
    float* logp;               // [S][B][K] log(p_code + eps) for nn.CTCLoss, or NULL (vqb_fwd_args.ctc_logp)
    float eps;
    const long long* lens;     // [N / S] valid frames per utterance, or NULL (length-aware rows, vqb_fwd_args.row_lengths)
    int S;                     // frames per utterance
    int N, K, num_tiles;
    int se_bytes;              // shared-memory bytes of the table region (image, later the fp32 gather table)
    unsigned flags;
};

#define VQB_PTL(tag) do { if (p.dbg && threadIdx.x == 0 && blockIdx.x == 0 && tl_n < 60) { p.dbg[tl_n++] = ((unsigned long long)(tag) << 56) | (globaltimer_ns() & 0x00FFFFFFFFFFFFFFull); } } while (0)

// exact fp32 score of one code for one row, read from global memory: the rare re-rank path.
// Same expression and fmaf order as vqb_fwd_simt.cu (dot_chunk + score_of).
template <int D>
__device__ __noinline__ float exact_score(const float* __restrict__ xrow, const float* __restrict__ e, float xx, float b,
                                          float tau, bool linear) {
    float dot = 0.f;
#pragma unroll 4
    for (int c = 0; c < D / 4; ++c) {
        const float4 xv = ldg4(xrow + 4 * c), w = ldg4(e + 4 * c);
        dot = fmaf(xv.x, w.x, dot); dot = fmaf(xv.y, w.y, dot);
        dot = fmaf(xv.z, w.z, dot); dot = fmaf(xv.w, w.w, dot);
    }
    if (linear) return dot + b;                                      // F.linear                   (:190)
    const float dist = __fsub_rn(__fadd_rn(xx, b), 2.f * dot);       // (|x|^2 + |e|^2) - 2 x.e   (:210-212)
    return tau * (-dist);                                            // relu(temp) * -dist        (:115, :213)
}

template <int KP, int D, bool LINEAR>
__global__ void __launch_bounds__(PM, 3)
vqb_fwd_pcode_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_q, PcP p) {
    constexpr int KB = D / 32;                      // raw fp32 blocks of an x tile
    constexpr int KS = D / 16;                      // MMA K-steps
    constexpr int DP = D + 4;                       // padded row of the fp32 gather table in shared memory
    const int K = p.K;
    const int KO = K | 1;                           // p_code staging row stride (odd: conflict-free)

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* sX = smem;                                             // [2][16 KB] raw x -> x_hi | x_lo -> new_latent
    uint8_t* sE = sX + 2 * PBLK;                                    // e_hi | e_lo image, later the fp32 gather table
    float* sTab = reinterpret_cast<float*>(sE);
    float* sP = reinterpret_cast<float*>(sE + p.se_bytes);          // [128][KO] p_code staging
    uint8_t* sHdr = reinterpret_cast<uint8_t*>(sP + PM * KO);       // image header: gE, emax | bias [64]
    const float* sBias = reinterpret_cast<const float*>(sHdr + 64);
    float* sRed = reinterpret_cast<float*>(sHdr + 64 + 256);        // [4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sRed + 4);
    uint64_t* x_full = bars;
    uint64_t* e_full = bars + 1;
    uint64_t* mma_done = bars + 2;
    uint64_t* t_full = bars + 3;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);
    // (only with p.logp) per row: offset of its [S][B][K] output row (int64), its arg-max code, log(p_top + eps)
    int4* sRow = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(bars) + 64);

    const int r = threadIdx.x, warp = r >> 5, lane = r & 31;
    int tl_n = 0;
    if (p.dbg && r == 0) p.dbg[128 + 2 * blockIdx.x] = globaltimer_ns();
    if (p.dbg && r == 0 && blockIdx.x == 0) p.dbg[60] = globaltimer_ns();

    if (r == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_q);
        mbar_init(x_full, 1); mbar_init(e_full, 1); mbar_init(mma_done, 1); mbar_init(t_full, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<64>(tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
    // PDL: the next kernel in the stream may begin its prologue; everything below that depends on the kernel BEFORE this
    // one (the table assembly: image, table, bias) sits behind pdl_wait().  x is older than that kernel.
    pdl_launch();
    VQB_PTL(1);

    const bool skip = (p.flags & VQB_SKIP) != 0;
    const float LOG2E = 1.4426950408889634f;
    float se_acc = 0.f;
    float temp_raw = 1.f;
    constexpr uint32_t IDESC = umma_idesc(0u, PM, KP);

    uint32_t it = 0;                                // tiles this CTA has COMPUTED (drives the mbarrier phases)
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int row0 = tile * PM;
        const int rows = min(PM, p.N - row0);
        bool valid = r < rows;
        bool live = valid;                          // length-aware rows: this row is a real frame, not padding
        if (p.lens) {
            // does the tile hold any real frame?  (a tile spans at most PM / S + 2 utterances)
            bool any = false;
            for (int b = row0 / p.S; b * p.S < row0 + rows; ++b) {
                const long long lo = max(row0, b * p.S), hi = min((long long)(row0 + rows), (long long)b * p.S + __ldg(p.lens + b));
                any = any || hi > lo;
            }
            if (!any) {
                // only padding: nothing is loaded or computed; the outputs of these rows are zero
                float4* q4 = reinterpret_cast<float4*>(p.q + (size_t)row0 * D);
                for (int i = r; i < rows * (D / 4); i += PM) q4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                float* pc = p.pcode + (size_t)row0 * K;
                for (int i = r; i < rows * K; i += PM) pc[i] = 0.f;
                if (valid) p.idx[row0 + r] = 0;
                if (p.logp && valid) {
                    const int b = (row0 + r) / p.S, sf = (row0 + r) - b * p.S;
                    float* orow = p.logp + ((size_t)sf * (p.N / p.S) + b) * K;
                    const float lz = logf(p.eps);
                    for (int k = 0; k < K; ++k) orow[k] = lz;
                }
                continue;
            }
            if (valid) {
                const int b = (row0 + r) / p.S;
                live = (row0 + r) - b * p.S < __ldg(p.lens + b);
            }
        }
        const uint32_t ph = it & 1;
        // ---- loads: the x tile (older than the previous kernel), then -- behind pdl_wait -- the table image ---------
        if (r == 0) {
            mbar_arrive_expect_tx(x_full, KB * PBLK);
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(sX + kb * PBLK, &tm_x, kb * 32, row0, x_full);
        }
        if (it == 0) {
            pdl_wait();
            if (!LINEAR) temp_raw = __ldg(p.temp);                  // consumed after the MMA: its latency is hidden
        }
        if (r == 0) {
            mbar_arrive_expect_tx(e_full, 2 * KP * 128 + (it == 0 ? 320 : 0));
            bulk_load_1d(sE, p.img, KP * 128, e_full);
            bulk_load_1d(sE + KP * 128, p.img + IMG_PIECE, KP * 128, e_full);
            if (it == 0) bulk_load_1d(sHdr, p.img + IMG_HDR, 320, e_full);   // scale, |e|_max and the 64 bias words
        }
        VQB_PTL(2);
        mbar_wait(x_full, ph);
        VQB_PTL(3);

        // ---- row -> registers; |x|^2 in the exact kernel's fmaf order; rescale; fp16x2 split in place ---------------
        float xr[D];
        float xx = 0.f, mx = 0.f;
#pragma unroll
        for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(sX + kb * PBLK + sw128_offset(r, c));
                xr[kb * 32 + 4 * c] = v.x; xr[kb * 32 + 4 * c + 1] = v.y; xr[kb * 32 + 4 * c + 2] = v.z; xr[kb * 32 + 4 * c + 3] = v.w;
            }
        }
#pragma unroll
        for (int d = 0; d < D; ++d) {
            xx = fmaf(xr[d], xr[d], xx);
            mx = fmaxf(mx, fabsf(xr[d]));
        }
        const int er = scale_exp(mx);
        {
            const float sx = pow2i(-er);
#pragma unroll
            for (int j = 0; j < D / 8; ++j) {
                uint4 hi, lo;
                split8(xr + 8 * j, sx, hi, lo);
                *reinterpret_cast<uint4*>(sX + sw128_offset(r, j)) = hi;
                *reinterpret_cast<uint4*>(sX + PBLK + sw128_offset(r, j)) = lo;
            }
        }
        fence_proxy_async_smem();                   // generic writes -> tcgen05.mma operand reads
        tcgen05_fence_before();
        __syncthreads();
        VQB_PTL(4);
        if (r == 0) {
            mbar_wait(e_full, ph);
            tcgen05_fence_after();
            const uint64_t ah = umma_desc_sw128(sX), al = umma_desc_sw128(sX + PBLK);
            const uint64_t bh = umma_desc_sw128(sE), bl = umma_desc_sw128(sE + KP * 128);
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {       // small terms first
                umma_bf16(tmem_base, al + 2 * ks, bl + 2 * ks, IDESC, ks != 0);
                umma_bf16(tmem_base, al + 2 * ks, bh + 2 * ks, IDESC, true);
                umma_bf16(tmem_base, ah + 2 * ks, bl + 2 * ks, IDESC, true);
                umma_bf16(tmem_base, ah + 2 * ks, bh + 2 * ks, IDESC, true);
            }
            umma_commit(mma_done);
        }
        mbar_wait(mma_done, ph);
        tcgen05_fence_after();
        VQB_PTL(5);
        // the MMAs have retired: the image's shared memory is dead, the fp32 gather table takes its place (bulk copies of
        // one padded row each, issued by warp 0; every thread waits on t_full right before its gather)
        if (warp == 0) {
            if (lane == 0) mbar_arrive_expect_tx(t_full, (uint32_t)(K * D * 4));
            __syncwarp();
            for (int k = lane; k < K; k += 32) bulk_load_1d(sTab + k * DP, p.gtab + (size_t)k * D, D * 4, t_full);
        }
        if (it == 0) mbar_wait(e_full, 0);           // the header (scale, bias) has landed with the image: visible to this thread

        // ---- scores (log2 domain) -> softmax -> p_code, argmax over p_code -----------------------------------------
        //   L2:     score = relu(temp) * -((|x|^2 + |e|^2) - 2 x.e)          (:115, :208-213)
        //   LINEAR: score = x.w + b                                           (:190)
        // Scores are kept in the natural domain, s = relu(temp) * -dist exactly as the exact kernel forms them (dist in the
        // reference's own association; 2 x.e is an exact scaling of the accumulator, so one fma rounds like the reference's
        // subtraction), and the exponent is taken of the exact difference (s - max) * log2(e): two scores that differ by
        // one ulp stay different in p_code, so its arg-max agrees with the exact kernel's down to the last bit.
        const float tau = fmaxf(temp_raw, 0.f);
        const float ntau = LINEAR ? 1.f : -tau;
        const float emax = *reinterpret_cast<const float*>(sHdr + 4);
        const float u = pow2i(er) * pow2i(*reinterpret_cast<const int*>(sHdr)) * (LINEAR ? 1.f : -2.f);
        float v[KP];
        auto load_scores = [&]() -> float {
            tmem_ld_cols<KP>(tmem_base + lane_addr, v);
            float m = -INFINITY;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                float sc = LINEAR ? fmaf(v[k], u, sBias[k]) : ntau * fmaf(v[k], u, __fadd_rn(xx, sBias[k]));
                if (k >= KP - 15) sc = k < K ? sc : -INFINITY;      // padded codes
                v[k] = sc;
                m = fmaxf(m, sc);
            }
            return m;
        };
        float s4[4];
        auto exp_scores = [&](float m) {            // v <- exp(v - m): the maximum itself gives ex2(0) = 1 exactly
            s4[0] = s4[1] = s4[2] = s4[3] = 0.f;
#pragma unroll
            for (int k = 0; k < KP; ++k) {
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((v[k] - m) * LOG2E));     // padded codes: ex2(-inf) = 0
                v[k] = e;
                s4[k & 3] += e;
            }
        };
        float m1 = load_scores();
        if (ntau == 0.f) {                          // temp <= 0: uniform over the K real codes only
#pragma unroll
            for (int k = 0; k < KP; ++k) v[k] = k < K ? 0.f : -INFINITY;
            m1 = 0.f;
        }
        exp_scores(m1);
        // Index exactness.  |acc - exact fp32 dot| <= 5e-6 |x||e| (operand pieces 2 * 2^-22, fp32 accumulation in the tensor
        // core and in the exact kernel's 64-term fmaf chain), twice for the two candidates, doubled in the distance, plus the
        // rounding of (|x|^2 + |e|^2) - 2 x.e itself: `win` in score units.  A second code can only lie inside that window
        // of the best one if the other codes' exponentials sum to at least ex2(-win) -- one compare per row; rows that pass
        // it (a handful per 10^5 at config 2) list the codes inside the window, re-evaluate them in exact fp32 (same
        // expression and fmaf order as vqb_fwd_simt.cu) and redo the softmax.
        {
            const float xn = sqrtf(xx);
            const float wd = LINEAR ? 1e-5f * xn * emax : 2e-5f * xn * emax + 2.4e-7f * (xn + emax) * (xn + emax);
            const float win = fabsf(ntau) * wd;
            float thr_e;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(thr_e) : "f"(-win * LOG2E));
            thr_e *= 0.999f;
            unsigned long long cand = 0ull;
            if (valid && ntau != 0.f && ((s4[0] + s4[1]) + (s4[2] + s4[3])) - 1.f >= thr_e) {
#pragma unroll
                for (int k = 0; k < KP; ++k) cand |= v[k] >= thr_e ? (1ull << k) : 0ull;
            }
            const bool need = __popcll(cand) >= 2;
            // tcgen05.ld is warp-collective: if any row of the warp needs the exact path, the whole warp reloads its scores
            // (rows that do not need it recompute the same softmax)
            if (__any_sync(0xffffffffu, need)) {
                m1 = load_scores();
                if (need) {
                    const float* xrow = p.x + (size_t)(row0 + r) * D;
#pragma unroll
                    for (int k = 0; k < KP; ++k)
                        if ((cand >> k) & 1ull) v[k] = exact_score<D>(xrow, p.table + (size_t)k * D, xx, sBias[k], tau, LINEAR);
                    m1 = -INFINITY;
#pragma unroll
                    for (int k = 0; k < KP; ++k) m1 = fmaxf(m1, v[k]);
                    if (p.stats) atomicAdd(p.stats, 1u);
                }
                exp_scores(m1);
            }
        }
        tcgen05_fence_before();
        // arg-max over p_code = e * inv (monotone in e), first index on ties (:130): the first code whose e is 1
        int best = 0;
#pragma unroll
        for (int k = KP - 1; k >= 0; --k) best = v[k] >= 1.f ? k : best;
        const float inv = live || !valid ? 1.f / ((s4[0] + s4[1]) + (s4[2] + s4[3])) : 0.f;    // pad rows: p_code = 0
        if (valid && !live) best = 0;
        {
            float* prow = sP + r * KO;
#pragma unroll
            for (int k = 0; k < KP; ++k)
                if (k < KP - 15 || k < K) prow[k] = v[k] * inv;    // softmax (:127)
        }
        if (valid) p.idx[row0 + r] = best;
        if (p.logp) {
            // the one probability of the row that can be close to 1 (its e is exactly 1, so p = inv) gets the full-precision
            // logarithm here; every other code has p <= 1/2 and is served by the fast one in the emission loop below
            const int gr_ = min(row0 + r, p.N - 1), b = gr_ / p.S, sf = gr_ - b * p.S;
            const long long off = ((long long)sf * (p.N / p.S) + b) * K;
            sRow[r] = make_int4((int)(off & 0xffffffffll), (int)(off >> 32), best, __float_as_int(logf(inv + p.eps)));
        }
        if (p.hist) {
            // warp-aggregated histogram: one atomic per distinct code per warp
            const unsigned peers = __match_any_sync(0xffffffffu, live ? best : -1);
            if (live && lane == (__ffs(peers) - 1)) atomicAdd(p.hist + best, (unsigned long long)__popc(peers));
        }
        VQB_PTL(6);
        mbar_wait(t_full, ph);                      // the gather table is complete in shared memory

        // ---- gather + straight-through: the row's result replaces its (dead) operand pieces in the x tile -----------
        {
            const float* crow = sTab + best * DP;
            const bool want_se = p.sqerr != nullptr;
#pragma unroll
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 cv = *reinterpret_cast<const float4*>(crow + kb * 32 + 4 * c);
                    const float x0 = xr[kb * 32 + 4 * c], x1 = xr[kb * 32 + 4 * c + 1], x2 = xr[kb * 32 + 4 * c + 2], x3 = xr[kb * 32 + 4 * c + 3];
                    float4 o;
                    if (LINEAR) {
                        o = cv;                                                         // (:194-197)
                    } else {
                        // new_latent = enc_embs + picked_code - enc_embs.detach()  (:145)
                        o.x = __fsub_rn(__fadd_rn(x0, cv.x), x0); o.y = __fsub_rn(__fadd_rn(x1, cv.y), x1);
                        o.z = __fsub_rn(__fadd_rn(x2, cv.z), x2); o.w = __fsub_rn(__fadd_rn(x3, cv.w), x3);
                        if (skip) o = make_float4(x0, x1, x2, x3);                      // (:142)
                    }
                    if (valid && !live) o = make_float4(0.f, 0.f, 0.f, 0.f);               // pad row
                    if (want_se && live) {
                        const float d0 = x0 - cv.x, d1 = x1 - cv.y, d2 = x2 - cv.z, d3 = x3 - cv.w;
                        se_acc = fmaf(d0, d0, se_acc); se_acc = fmaf(d1, d1, se_acc);
                        se_acc = fmaf(d2, d2, se_acc); se_acc = fmaf(d3, d3, se_acc);
                    }
                    *reinterpret_cast<float4*>(sX + kb * PBLK + sw128_offset(r, c)) = o;
                }
            }
        }
        // ---- this warp's 32 rows leave on their own: no CTA-wide barrier between the softmax and the stores -----------
        fence_proxy_async_smem();                   // tile / p_code writes -> bulk stores
        __syncwarp();
        VQB_PTL(7);
        const int rows_w = min(32, rows - 32 * warp);        // rows of this warp's slab inside the tensor (<= 0: none)
        if (rows_w > 0) {
            float* dst = p.pcode + (size_t)(row0 + 32 * warp) * K;
            const float* src = sP + 32 * warp * KO;
            const int n = rows_w * K;
            if (lane == 0) {
                // new_latent slab: TMA store straight from the swizzled tile (rows beyond N are clipped by TMA)
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) tma_store_2d(&tm_q, sX + kb * PBLK + warp * 32 * 128, kb * 32, row0 + 32 * warp);
                if (K & 1) {
                    const uint32_t bytes = (uint32_t)(n * 4) & ~15u;
                    if (bytes) bulk_store_1d(dst, src, bytes);
                }
                tma_store_commit();
            }
            if (K & 1) {
                const int done = (int)(((uint32_t)(n * 4) & ~15u) >> 2);   // < 16 bytes of a ragged last slab
                if (lane < n - done) dst[done + lane] = src[done + lane];
            } else {
                int rr = lane / K, k = lane - rr * K;           // running (row, code) of element i
                const int step_r = 32 / K, step_k = 32 - step_r * K;
                for (int i = lane; i < n; i += 32) {
                    __stcs(dst + i, src[rr * KO + k]);
                    rr += step_r; k += step_k;
                    if (k >= K) { k -= K; ++rr; }
                }
            }
            if (p.logp) {
                // CTC input, emitted from the staged rows: out[s][b][:] = log(p[b][s][:] + eps) (bin/train_vqvae.py:430-432).
                // The warp walks its slab element by element; the row's output offset, its arg-max code and that code's
                // precise logarithm come from sRow (written by the row's thread above, same warp).
                // lane = consecutive codes of a row, so every store instruction is one or two contiguous runs of the
                // [S][B][K] tensor.  log = MUFU.LG2 * ln 2 (|error| < 2 ulp for p <= 1/2, where |log| >= 0.69).
                int rr = lane / K, k = lane - rr * K;
                const int step_r = 32 / K, step_k = 32 - step_r * K;
                const int4* rowinfo = sRow + 32 * warp;
#pragma unroll 4
                for (int i = lane; i < n; i += 32) {
                    const int4 ri = rowinfo[rr];
                    float lg;
                    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(src[rr * KO + k] + p.eps));
                    const long long off = ((long long)ri.y << 32) | (unsigned int)ri.x;
                    __stcs(p.logp + off + k, k == ri.z ? __int_as_float(ri.w) : lg * 0.693147181f);
                    rr += step_r; k += step_k;
                    if (k >= K) { k -= K; ++rr; }
                }
            }
            if (lane == 0) tma_store_wait_read();   // this warp's shared memory may be overwritten from here on
        }
        VQB_PTL(8);
        if (tile + (int)gridDim.x < p.num_tiles) __syncthreads();       // another tile follows: the buffers are free
        ++it;
    }
    if (p.sqerr) {
        se_acc = warp_sum(se_acc);
        if (lane == 0) sRed[warp] = se_acc;
        __syncthreads();
        if (r == 0) atomicAdd(p.sqerr, (double)sRed[0] + (double)sRed[1] + (double)sRed[2] + (double)sRed[3]);
    }
    VQB_PTL(9);

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<64>(tmem_base);
    if (p.dbg && r == 0 && blockIdx.x == 0) p.dbg[61] = globaltimer_ns();
    if (p.dbg && r == 0) p.dbg[129 + 2 * blockIdx.x] = globaltimer_ns();
}

// -----------------------------------------------------------------------------------------------------------
// operand image of a table that was not assembled by vqb_assemble_table (LINEAR score; raw C-ABI calls without a cache)
// -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
build_image_kernel(const float* __restrict__ w, const float* __restrict__ bias, int K, int D, uint8_t* __restrict__ img) {
    __shared__ float s_max[8], s_nrm[8];
    pdl_launch();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float gmax = 0.f, nmax = 0.f;
    for (int k = warp; k < K; k += 8) {
        float sq = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float v = __ldg(w + (size_t)k * D + d);
            gmax = fmaxf(gmax, fabsf(v));
            sq = fmaf(v, v, sq);
        }
        nmax = fmaxf(nmax, warp_sum(sq));
    }
    gmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(gmax)));   // non-negative floats order like uints
    if (lane == 0) { s_max[warp] = gmax; s_nrm[warp] = nmax; }
    __syncthreads();
    gmax = 0.f; nmax = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { gmax = fmaxf(gmax, s_max[i]); nmax = fmaxf(nmax, s_nrm[i]); }
    write_image(w, D, K, D, gmax, sqrtf(nmax), bias, img);
}

int launch_build_image(const float* w, const float* bias, int K, int D, void* img, cudaStream_t s) {
    build_image_kernel<<<1, 256, 0, s>>>(w, bias, K, D, reinterpret_cast<uint8_t*>(img));
    VQB_CHECK_LAUNCH("build_image_kernel");
    return VQB_OK;
}

// -----------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------
unsigned long long* get_debug_timeline();

bool forward_pcode_supported(const vqb_fwd_args* a) {
    return a->p_code != nullptr && a->n_codes <= 64 && (a->dim == 32 || a->dim == 64);
}

size_t forward_pcode_workspace(const vqb_fwd_args* a) { return a->operand_cache ? 0 : (size_t)IMG_BYTES; }

template <int KP, int D, bool LINEAR>
static int launch_pc(const CUtensorMap& tx, const CUtensorMap& tq, PcP p, cudaStream_t s, bool pdl) {
    const int tab_bytes = p.K * (D + 4) * 4, img_bytes = 2 * KP * 128;
    p.se_bytes = ((tab_bytes > img_bytes ? tab_bytes : img_bytes) + 1023) & ~1023;
    const size_t smem = (size_t)2 * PBLK + p.se_bytes + (size_t)PM * (p.K | 1) * 4 + 320 + 4 * 4 + 64 + (p.logp ? PM * 16 : 0) + 1024;
    auto kern = vqb_fwd_pcode_kernel<KP, D, LINEAR>;
    { const int rc_ = ensure_smem(kern, smem, true); if (rc_) return rc_; }
    const int slots = 3 * sm_count();
    const int grid = p.num_tiles < slots ? p.num_tiles : slots;
    kernel_event_begin(s);
    if (pdl) VQB_CUDA(launch_pdl(kern, dim3(grid), dim3(PM), smem, s, tx, tq, p));
    else kern<<<grid, PM, smem, s>>>(tx, tq, p);
    kernel_event_end(s);
    VQB_CHECK_LAUNCH("vqb_fwd_pcode_kernel");
    return VQB_OK;
}

int launch_forward_pcode(const vqb_fwd_args* a, cudaStream_t s) {
    const int64_t N = a->n_rows, K = a->n_codes, D = a->dim;
    if (N == 0) return VQB_OK;
    const bool cached = a->operand_cache != nullptr;
    if (!cached && (!a->workspace || a->workspace_bytes < (size_t)IMG_BYTES)) {
        set_error("vqb_forward: workspace too small (%zu < %d bytes)", a->workspace_bytes, IMG_BYTES);
        return VQB_ERR_WORKSPACE;
    }
    const uint8_t* img = reinterpret_cast<const uint8_t*>(cached ? a->operand_cache : a->workspace);
    if (!cached) {
        int rc = launch_build_image(a->score_w, a->score_b, (int)K, (int)D, a->workspace, s);
        if (rc) return rc;
    }
    CUtensorMap tx, tq;
    int rc = make_tmap_2d_f32(&tx, a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, PM);
    if (rc) return rc;
    if ((rc = make_tmap_2d_f32(&tq, a->new_latent, (uint64_t)N, (uint64_t)D, (uint64_t)D, 32))) return rc;   // one warp's slab per store
    PcP p;
    p.x = a->x; p.table = a->score_w; p.gtab = a->gather_table; p.bias = a->score_b; p.temp = a->temp; p.img = img;
    p.pcode = a->p_code; p.q = a->new_latent; p.idx = (long long*)a->idx; p.hist = (unsigned long long*)a->hist; p.sqerr = a->sq_err_sum;
    p.stats = a->search_stats; p.dbg = get_debug_timeline();
    p.N = (int)N; p.K = (int)K; p.num_tiles = (int)ceil_div(N, PM); p.se_bytes = 0; p.flags = a->flags;
    p.lens = (const long long*)a->row_lengths; p.S = (int)a->frames_per_utt;
    p.logp = a->ctc_logp; p.eps = a->ctc_eps;
    // PDL when the kernel enqueued immediately before is ours: the image build above, or (the caller vouches,
    // VQB_AFTER_ASSEMBLE) the table assembly
    const bool pdl = !cached || (a->flags & VQB_AFTER_ASSEMBLE);
    const int KP = (int)((K + 15) / 16 * 16);
    const bool lin = (a->flags & VQB_SCORE_LINEAR) != 0;
#define VQB_PC_CASE(kp, d) \
    if (KP == kp && D == d) return lin ? launch_pc<kp, d, true>(tx, tq, p, s, pdl) : launch_pc<kp, d, false>(tx, tq, p, s, pdl);
    VQB_PC_CASE(16, 32) VQB_PC_CASE(32, 32) VQB_PC_CASE(48, 32) VQB_PC_CASE(64, 32)
    VQB_PC_CASE(16, 64) VQB_PC_CASE(32, 64) VQB_PC_CASE(48, 64) VQB_PC_CASE(64, 64)
#undef VQB_PC_CASE
    return invalid("vqb_forward: shape not supported by the parity-mode tensor-core kernel");
}

}  // namespace vqb
