// Fused-mode forward of the quantizer (no p_code: the large-codebook search of BASELINE config 3) on the 5th-generation
// tensor cores (tcgen05 + TMEM + TMA): nearest-codeword search + exact re-rank + gather + straight-through; the N x K
// distance matrix never leaves TMEM.  Replaces neg_batch_l2 + argmax + F.embedding + straight-through of
// src/embed.py:208-213, :130, :134, :145.  (The parity-mode forward, where p_code is part of the result, is vqb_fwd_pc.cu.)
//
// Anatomy (one persistent CTA per SM, 6 or 10 warps, warp-specialised):
//   warp 0  TMA producer   x tile [128 rows][D] fp32 (double-buffered when it fits); codebook K-blocks
//                          [BN codes][32 floats]: resident in shared memory when the codebook is one chunk,
//                          otherwise streamed through a ring
//   warp 1  MMA issuer     tcgen05.mma kind::tf32 / kind::f16, M=128, N=BN; two (resident codebook) or four (streamed)
//                          accumulator buffers in TMEM
//   warps 2-5 epilogue     tcgen05.ld 32x32b: thread = row, so the running minimum / candidate list are thread-local
//   warps 6-9 (EW = 2: streamed 1xTF32 search, D <= 128) a second epilogue warpgroup on the other column half of every chunk
//
// Precision.  kind::tf32 reads the top 19 bits of each fp32 operand.  PASSES = 3 ("3xTF32") adds the two
// cross terms with the operands' low parts: x = x_hi + x_lo (x_hi = hardware truncation of the raw tile, x_lo
// written by the epilogue warps into a second tile), e = e_hi + e_lo (split once per call by the prep kernel),
// acc = x.e_hi + x.e_lo + x_lo.e_hi, which restores fp32-level accuracy (error <= 3 * 2^-20 |x||e|).
// PASSES = 1 is used where the MMA is the bound (large codebooks); its error 1.5 * 2^-10 |x||e| is covered by a
// provable candidate window + exact fp32 re-rank (per-row candidate list; full exact scan if it overflows).
// The bias (|e|^2 for L2, b for LINEAR) is folded into the GEMM as one extra K-step: A = [1,1,1,0,..],
// B = the bias split into three tf32-exact words, so the accumulator is directly |e|^2 - 2 x.e (or x.w + b).
#include <cudaTypedefs.h>
#include <math.h>
#include "vqb_common.cuh"
#include "vqb_tc.cuh"
#include "vqb_f16x2.cuh"

namespace vqb {
using namespace tc;

constexpr int BM = 128;                 // rows per tile (UMMA M)
constexpr int XBLK = BM * 128;          // one x K-block: [128 rows][32 fp32] = 16 KB

struct TcP {
    const float* table;        // [K][D] fp32 score table (exact re-rank)
    const float* gtab;         // [K][D] fp32 gather table
    const float* bias;         // [K]    |e|^2 (L2) or b (LINEAR)
    const float* temp;         // [1]
    const float* emax;         // [1]    max_k |e_k|
    const int* gexp;           // [1]    F16 mode: scale exponent of the fp16 codebook copy (copy = table * 2^-gexp)
    long long* idx;
    float* q;
    unsigned long long* hist;
    double* sqerr;
    unsigned int* stats;       // [0] rows re-ranked, [1] rows that needed the full exact scan
    unsigned long long* dbg;   // optional timeline buffer (vqb_debug_set_timeline), NULL in production
    int N, K, D, num_tiles, num_chunks;
    unsigned flags;
};


// hi[k] = [tf32_rn(scale * w_k) (D floats) | bias_k as three tf32-exact words | 0 x 29]   ([Kpad][D+32])
// lo[k] = scale * w_k - hi[k]                                                              ([Kpad][D])
// emax  = max_k |w_k|
// Rows K..Kpad-1 (padding up to a whole chunk) are zero with bias `pad_bias` (+1e30 for the L2 score, -1e30 for
// LINEAR), so a padded code can never be the arg-min / receive probability mass: no masking in the epilogues.
__global__ void __launch_bounds__(128)
build_operands_kernel(const float* __restrict__ w, const float* __restrict__ bias, int K, int D, float scale,
                      float pad_bias, float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ emax) {
    const int k = blockIdx.x;
    float* hrow = hi + (size_t)k * (D + 32);
    if (k >= K) {
        for (int d = threadIdx.x; d < D + 32; d += blockDim.x) hrow[d] = d == D ? pad_bias : 0.f;
        if (lo) for (int d = threadIdx.x; d < D; d += blockDim.x) lo[(size_t)k * D + d] = 0.f;
        return;
    }
    float sq = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float raw = w[(size_t)k * D + d];
        const float v = scale * raw;
        const float h = tf32_rn(v);
        hrow[d] = h;
        if (lo) lo[(size_t)k * D + d] = v - h;
        sq = fmaf(raw, raw, sq);
    }
    if (threadIdx.x < 32) {
        const float b = bias ? bias[k] : 0.f;
        const float b0 = tf32_trunc(b);
        const float r1 = b - b0;
        const float b1 = tf32_trunc(r1);
        const float b2 = r1 - b1;
        const int j = threadIdx.x;
        hrow[D + j] = j == 0 ? b0 : (j == 1 ? b1 : (j == 2 ? b2 : 0.f));
    }
    __shared__ float red[4];
    sq = warp_sum(sq);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0 && emax)
        atomicMax(reinterpret_cast<int*>(emax), __float_as_int(sqrtf(red[0] + red[1] + red[2] + red[3])));
}

void launch_build_operands(const float* w, const float* bias, int K, int Kpad, int D, float scale, float pad_bias,
                           float* hi, float* lo, float* emax, cudaStream_t s) {
    build_operands_kernel<<<(unsigned)Kpad, 128, 0, s>>>(w, bias, K, D, scale, pad_bias, hi, lo, emax);
}

// F16 mode, step 1: table-wide |w|_max (for the common power-of-two scale) and max_k |w_k|_2 (for the candidate window)
__global__ void __launch_bounds__(128)
table_max_kernel(const float* __restrict__ w, int D, unsigned int* __restrict__ gmax_bits, float* __restrict__ emax) {
    const int k = blockIdx.x;
    float sq = 0.f, mx = 0.f;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        const float v = w[(size_t)k * D + d];
        sq = fmaf(v, v, sq);
        mx = fmaxf(mx, fabsf(v));
    }
    __shared__ float red[4];
    __shared__ unsigned int redm[4];
    sq = warp_sum(sq);
    const unsigned int wm = __reduce_max_sync(0xffffffffu, __float_as_uint(mx));      // non-negative floats order like uints
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = sq; redm[threadIdx.x >> 5] = wm; }
    __syncthreads();
    if (threadIdx.x == 0) {
        atomicMax(reinterpret_cast<int*>(emax), __float_as_int(sqrtf(red[0] + red[1] + red[2] + red[3])));
        atomicMax(gmax_bits, max(max(redm[0], redm[1]), max(redm[2], redm[3])));
    }
}
// step 2: e16[k][d] = fp16_rn(w[k][d] * 2^-gE), rows K..Kpad-1 zero (their |e|^2 is +1e30 in the epilogue); gE -> *gexp
__global__ void __launch_bounds__(128)
build_f16_operands_kernel(const float* __restrict__ w, int K, int D, const unsigned int* __restrict__ gmax_bits,
                          __half* __restrict__ e16, int* __restrict__ gexp) {
    const int k = blockIdx.x;
    const int gE = scale_exp(__uint_as_float(*gmax_bits));
    const float sE = pow2i(-gE);
    if (k == 0 && threadIdx.x == 0) *gexp = gE;
    for (int d = threadIdx.x; d < D; d += blockDim.x)
        e16[(size_t)k * D + d] = __float2half_rn(k < K ? w[(size_t)k * D + d] * sE : 0.f);
}

// exact score of code k for the row held (swizzled) in shared memory -- same expression and fmaf order as
// the exact SIMT kernel (vqb_fwd_simt.cu: dot_chunk + score_of)
template <int KB>
__device__ __forceinline__ float exact_score(const uint8_t* sXt, int r, float xx, const float* __restrict__ table,
                                             const float* __restrict__ enorm, int k, float tau) {
    const float* e = table + (size_t)k * (KB * 32);
    float dot = 0.f;
#pragma unroll 1
    for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 xv = *reinterpret_cast<const float4*>(sXt + kb * XBLK + sw128_offset(r, c));
            const float4 w = ldg4(e + kb * 32 + c * 4);
            dot = fmaf(xv.x, w.x, dot); dot = fmaf(xv.y, w.y, dot);
            dot = fmaf(xv.z, w.z, dot); dot = fmaf(xv.w, w.w, dot);
        }
    }
    const float dist = __fsub_rn(__fadd_rn(xx, __ldg(enorm + k)), 2.f * dot);
    return tau * (-dist);
}

// PIPE (see the kernel): |x|^2 in the exact kernel's fmaf order and x_lo = x - trunc_tf32(x), the operand of the third
// MMA pass, for the tile with per-CTA counter `it`; returns |x|^2 of this thread's row `r`.  Same statements as the
// in-loop form of the kernel (streamed 3xTF32 search, one epilogue warpgroup: x_lo slot == x slot).
template <int KB, int XS>
__device__ __forceinline__ float prep_tile(const uint8_t* sX, uint8_t* sXlo, uint64_t* x_full, uint64_t* xlo_full, int r, int lane,
                                           uint32_t it) {
    const uint32_t xs = it % XS, xph = (it / XS) & 1;
    const uint8_t* sXt = sX + (size_t)xs * KB * XBLK;
    uint8_t* sXl = sXlo + (size_t)xs * KB * XBLK;
    mbar_wait(&x_full[xs], xph);
    float xx = 0.f;
#pragma unroll 1
    for (int kb = 0; kb < KB; ++kb) {
        float4 xv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) xv[c] = *reinterpret_cast<const float4*>(sXt + kb * XBLK + sw128_offset(r, c));
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            xx = fmaf(xv[c].x, xv[c].x, xx); xx = fmaf(xv[c].y, xv[c].y, xx);
            xx = fmaf(xv[c].z, xv[c].z, xx); xx = fmaf(xv[c].w, xv[c].w, xx);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float4 lo;
            lo.x = xv[c].x - tf32_trunc(xv[c].x); lo.y = xv[c].y - tf32_trunc(xv[c].y);
            lo.z = xv[c].z - tf32_trunc(xv[c].z); lo.w = xv[c].w - tf32_trunc(xv[c].w);
            *reinterpret_cast<float4*>(sXl + kb * XBLK + sw128_offset(r, c)) = lo;
        }
    }
    fence_proxy_async_smem();                                       // generic writes -> tcgen05.mma operand reads
    __syncwarp();
    if (lane == 0) mbar_arrive(&xlo_full[xs]);
    return xx;
}

#define VQB_TL(tag) do { if (p.dbg && et == 0 && wg == 0 && cg == 0 && blockIdx.x == 0 && tl_n < 120) { p.dbg[tl_n++] = ((unsigned long long)(tag) << 56) | (globaltimer_ns() & 0x00FFFFFFFFFFFFFFull); } } while (0)

// NOAUG (streamed 1xTF32 search at D = 256 only): the |e|^2 term is not folded into the GEMM as an extra K-step but
// added by the epilogue from a per-chunk copy of enorm in shared memory; this frees the 16 KB A block and one piece
// per chunk, which buys a fifth ring slot where shared memory is otherwise full.
// (Measured and dropped in round 2, profiles/r2_search_experiments.txt: a column-split second epilogue warpgroup, clusters
// of two with the codebook pieces multicast, and the one-pass kernel below K = 1024 -- none was faster.)
// F16 (the one-pass search, K > 1024 or D = 256): operands as single fp16 pieces instead of tf32 -- the same 11 significant
// bits, so the same candidate window, but kind::f16 runs at twice the tf32 rate and an operand byte carries twice the
// flops: the x tile is half the size, a 16 KB codebook piece covers 64 dimensions instead of 32, and the ring holds twice
// the work in flight (the streamed search is bound by the bytes in flight from L2, profiles/r1e_ncu_full_search.csv).
// The row threads rescale their row by a power of two and convert it IN PLACE over the raw TMA tile (x16 block j takes
// the place of raw block j, which has been read by then); the codebook copy is converted once per call
// (build_f16_operands_kernel).  Exactness is untouched: the re-rank, the gather and the straight-through need the raw fp32
// tile again, so it is simply fetched a second time (from L2) once the tile's last MMA has retired -- 128 KB against the
// 8 MB of codebook that streamed through meanwhile.
template <int KB, int BN, int XS, int BS, int PASSES, bool RESIDENT, bool NOAUG, bool F16, int EW>
__global__ void __launch_bounds__(64 + 128 * EW, 1)
vqb_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_hi,
                  const __grid_constant__ CUtensorMap tm_lo, const __grid_constant__ CUtensorMap tm_q, TcP p) {
    constexpr int PIECE = BN * 128;                               // one codebook K-block in bytes
    constexpr int KH = KB / 2;                                    // fp16 blocks of 64 dimensions (F16)
    constexpr int PIECES = F16 ? KH : (PASSES == 3 ? 2 * KB : KB) + (NOAUG ? 0 : 1);   // per chunk: hi/lo per K-block + the bias block
    static_assert(!NOAUG || (PASSES == 1 && !RESIDENT), "NOAUG: streamed one-pass search");
    static_assert(!F16 || (NOAUG && KB % 2 == 0), "F16: one-pass streamed search, D a multiple of 64, bias added by the epilogue");
    // EW = 2 (streamed 1xTF32 search, D <= 128): a SECOND epilogue warpgroup.  Both read the same TMEM lanes (thread = row)
    // but different column halves of every chunk, each with its own running minimum and candidate list; the lists are
    // merged under the window of the smaller minimum before the exact re-rank (each list is a superset of what that
    // window needs from its half, because a half's own threshold is never tighter).  At D <= 128 the chunk time is the
    // epilogue's, not the MMA's (profiles/r2_ncu_lines_search_d64.txt), and one warp per scheduler hides nothing.
    static_assert(EW == 1 || (EW == 2 && PASSES == 1 && !RESIDENT && !NOAUG && !F16), "second epilogue warpgroup: streamed 1xTF32 search");
    // accumulator buffers in TMEM: four for the streamed kernels (all 512 columns; one CTA per SM) -- the epilogue's time per
    // chunk varies with the candidate scans, the MMA's does not, and two buffers let each side wait for the other
    constexpr int NB = RESIDENT ? 2 : 4;
    constexpr int TMEM_COLS = NB * BN;
    constexpr int CAP = NOAUG ? 15 : 16;                          // per-row candidate list capacity (SEARCH); 15: fits 227 KB
    static_assert(!RESIDENT || BS == PIECES, "resident codebook needs one slot per piece");
    constexpr int NTHREADS = 64 + 128 * EW;
    constexpr int XLS = XS;                                       // x_lo slots

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (tc::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS)
    uint8_t* sX = smem;                                                        // [XS][KB][16 KB]
    uint8_t* sXlo = sX + (size_t)XS * KB * XBLK;                               // [XS][KB][16 KB]  (PASSES == 3)
    uint8_t* sAug = sXlo + (PASSES == 3 ? (size_t)XLS * KB * XBLK : 0);        // [16 KB] A block [1,1,1,0,...]
    uint8_t* sB = sAug + (NOAUG ? 0 : XBLK);                                   // [BS][PIECE]
    uint2* sCand = reinterpret_cast<uint2*>(sB + (size_t)BS * PIECE);          // [CAP][128] (value, code)
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sCand) + EW * CAP * BM * 8);   // [EW] lists
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + XS;
    uint64_t* xlo_full = x_empty + XS;
    uint64_t* b_full = xlo_full + XS;
    uint64_t* b_empty = b_full + BS;
    uint64_t* t_full = b_empty + BS;
    uint64_t* t_empty = t_full + NB;
    uint64_t* xlo_free = t_empty + NB;
    uint64_t* xr_full = xlo_free + 1;                             // [XS] F16: the raw x tile has been fetched again
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xr_full + XS);
    // per-row index words [128] and (NOAUG) the |e|^2 staging [2][BN] behind the barriers, 16-byte aligned (float4 reads)
    int* sAfterBars = reinterpret_cast<int*>(smem + ((reinterpret_cast<uint8_t*>(tmem_slot + 4) - smem + 15) & ~(size_t)15));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) p.dbg[120] = globaltimer_ns();

    // ---- one-time setup ----------------------------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_hi);
        if (PASSES == 3) tma_prefetch_desc(&tm_lo);
        tma_prefetch_desc(&tm_q);
        for (int i = 0; i < XS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 4); mbar_init(&xlo_full[i], 4); }
        for (int i = 0; i < BS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < NB; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4 * EW); }
        mbar_init(xlo_free, 1);
        for (int i = 0; i < XS; ++i) mbar_init(&xr_full[i], 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
    if (!NOAUG)
        for (int i = threadIdx.x; i < XBLK / 16; i += NTHREADS)
            reinterpret_cast<float4*>(sAug)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) p.dbg[121] = globaltimer_ns();
    if (!NOAUG && threadIdx.x < BM)
        *reinterpret_cast<float4*>(sAug + sw128_offset(threadIdx.x, 0)) = make_float4(1.f, 1.f, 1.f, 0.f);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const int tile_end = p.num_tiles;
    const uint32_t tmem_base = *tmem_slot;
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) p.dbg[122] = globaltimer_ns();
    constexpr uint32_t IDESC = umma_idesc(F16 ? 0u : 2u, BM, BN);
    // PDL: let the next kernel in the stream begin its prologue; everything below that depends on the kernel BEFORE
    // this one (the operand / table preparation) sits behind pdl_wait().  x is older than that kernel, so the
    // producer may fetch its first tile before waiting.
    pdl_launch();
    if (warp != 0) pdl_wait();

    if (warp == 0) {
        // =============================== TMA producer =====================================================
        if (lane == 0) {
            uint32_t x_it = 0, b_it = 0;
            bool resident_loaded = false;
            for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
                const uint32_t xs = x_it % XS, xph = (x_it / XS) & 1;
                mbar_wait(&x_empty[xs], xph ^ 1);
                mbar_arrive_expect_tx(&x_full[xs], KB * XBLK);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(sX + ((size_t)xs * KB + kb) * XBLK, &tm_x, kb * 32, tile * BM, &x_full[xs]);
                ++x_it;
                if (RESIDENT && resident_loaded) continue;
                if (!resident_loaded) pdl_wait();
                resident_loaded = true;
                for (int chunk = 0; chunk < p.num_chunks; ++chunk) {
                    for (int j = 0; j < PIECES; ++j) {
                        const uint32_t bs = b_it % BS, bph = (b_it / BS) & 1;
                        if (!RESIDENT) mbar_wait(&b_empty[bs], bph ^ 1);
                        mbar_arrive_expect_tx(&b_full[bs], PIECE);
                        const bool is_lo = PASSES == 3 && j < 2 * KB && (j & 1);
                        const int kb = PASSES == 3 ? (j >> 1) : j;                 // the bias block has kb == KB
                        tma_load_2d(sB + (size_t)bs * PIECE, is_lo ? &tm_lo : &tm_hi, F16 ? j * 64 : kb * 32, chunk * BN, &b_full[bs]);
                        ++b_it;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer =======================================================
        if (lane == 0) {
            uint32_t x_it = 0, b_it = 0, c_it = 0;
            bool resident_ready = false;
            for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
                const uint32_t xs = x_it % XS, xph = (x_it / XS) & 1;
                const uint8_t* xt = sX + (size_t)xs * KB * XBLK;
                const uint32_t xls = xs, xlph = xph;
                const uint8_t* xlt = sXlo + (size_t)xls * KB * XBLK;
                mbar_wait(&x_full[xs], xph);
                if ((PASSES == 3 && !RESIDENT) || F16) mbar_wait(&xlo_full[xls], xlph);   // x_lo written / x converted to fp16
                tcgen05_fence_after();
                for (int chunk = 0; chunk < p.num_chunks; ++chunk) {
                    const uint32_t buf = c_it % NB, tph = (c_it / NB) & 1;
                    mbar_wait(&t_empty[buf], tph ^ 1);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * BN;
                    if (RESIDENT) {
                        if (!resident_ready) {
                            for (int j = 0; j < PIECES; ++j) mbar_wait(&b_full[j], 0);
                            tcgen05_fence_after();
                            resident_ready = true;
                        }
                        // group A: x . e_hi (+ bias);  group B: x . e_lo;  group C (needs x_lo): x_lo . e_hi
#pragma unroll
                        for (int kb = 0; kb < KB; ++kb) {
                            const uint64_t a = umma_desc_sw128(xt + kb * XBLK);
                            const uint64_t b = umma_desc_sw128(sB + (size_t)(PASSES == 3 ? 2 * kb : kb) * PIECE);
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, a + 2 * k, b + 2 * k, IDESC, (kb | k) != 0);
                        }
                        if constexpr (!NOAUG)
                        umma_tf32(d_tmem, umma_desc_sw128(sAug), umma_desc_sw128(sB + (size_t)(PIECES - 1) * PIECE), IDESC, true);
                        if (PASSES == 3) {
#pragma unroll
                            for (int kb = 0; kb < KB; ++kb) {
                                const uint64_t a = umma_desc_sw128(xt + kb * XBLK);
                                const uint64_t b = umma_desc_sw128(sB + (size_t)(2 * kb + 1) * PIECE);
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, a + 2 * k, b + 2 * k, IDESC, true);
                            }
                            mbar_wait(&xlo_full[xls], xlph);
                            tcgen05_fence_after();
#pragma unroll
                            for (int kb = 0; kb < KB; ++kb) {
                                const uint64_t a = umma_desc_sw128(xlt + kb * XBLK);
                                const uint64_t b = umma_desc_sw128(sB + (size_t)(2 * kb) * PIECE);
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, a + 2 * k, b + 2 * k, IDESC, true);
                            }
                        }
                    } else {
                        for (int j = 0; j < PIECES; ++j) {
                            const uint32_t bs = b_it % BS, bph = (b_it / BS) & 1;
                            mbar_wait(&b_full[bs], bph);
                            tcgen05_fence_after();
                            const uint64_t b = umma_desc_sw128(sB + (size_t)bs * PIECE);
                            const bool is_aug = !NOAUG && j == PIECES - 1;
                            const bool is_lo = PASSES == 3 && !is_aug && (j & 1);
                            const int kb = PASSES == 3 ? (j >> 1) : j;
                            if constexpr (F16) {
                                const uint64_t a = umma_desc_sw128(xt + j * XBLK);          // fp16 block j: 64 dimensions
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_bf16(d_tmem, a + 2 * k, b + 2 * k, IDESC, (j | k) != 0);
                            } else if (is_aug) {
                                umma_tf32(d_tmem, umma_desc_sw128(sAug), b, IDESC, true);
                            } else {
                                const uint64_t a = umma_desc_sw128(xt + kb * XBLK);
#pragma unroll
                                for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, a + 2 * k, b + 2 * k, IDESC, (j | k) != 0);
                                if (PASSES == 3 && !is_lo) {
                                    const uint64_t al = umma_desc_sw128(xlt + kb * XBLK);
#pragma unroll
                                    for (int k = 0; k < 4; ++k) umma_tf32(d_tmem, al + 2 * k, b + 2 * k, IDESC, true);
                                }
                            }
                            umma_commit(&b_empty[bs]);              // frees the ring slot when these MMAs retire
                            ++b_it;
                        }
                    }
                    umma_commit(&t_full[buf]);                      // accumulator of this chunk is complete
                    ++c_it;
                }
                ++x_it;
            }
        }
    } else {
        // =============================== epilogue (thread = row) ==========================================
        const int q4 = warp & 3;                                    // TMEM lane quadrant this warp may read
        const int r = q4 * 32 + lane;                               // row within the tile == TMEM lane
        const int wg = (warp - 2) >> 2, cg = wg;                    // epilogue warpgroup (0: primary; 1: second column half, EW == 2)
        const int et = ((warp - 2) & 3) * 32 + lane;                // 0..127 within the warpgroup
        // (EW == 2) minimum / count / overflow of the second half: the last entry row of the second list (which holds CAP - 1)
        int2* sMerge = reinterpret_cast<int2*>(sCand + (size_t)(2 * CAP - 1) * BM);
        const int cap = CAP - ((EW == 2 && wg == 1) ? 1 : 0);
        const uint32_t bar_id = 1 + wg;                             // named barrier of this warpgroup
        const bool linear = (p.flags & VQB_SCORE_LINEAR) != 0;
        const float tau = linear ? 1.f : fmaxf(__ldg(p.temp), 0.f);
        const float emax = __ldg(p.emax);
        const uint32_t lane_addr = (uint32_t)(q4 * 32) << 16;
        uint32_t x_it = 0, c_it = 0;                                // tile / chunk counters
        float se_acc = 0.f;
        int tl_n = 0;
        VQB_TL(1);
        // PIPE (streamed 3xTF32 search with two x / x_lo slots and at most two chunks per tile, i.e. 128 < K <= 256):
        // x_lo of the NEXT tile is produced before the epilogue of the current one, so the MMA warp -- which cannot start
        // a streamed tile without x_lo -- runs tile t+1 (and the producer streams its codebook pieces) while these warps
        // scan, re-rank, gather and store tile t.  Slot (t+1) % 2 is free by then: its previous user, tile t-1, was
        // drained (x slot: x_empty -> TMA refill -> x_full; x_lo slot: every MMA of t-1 had retired before the last
        // t_full of t-1 was consumed).  Why two chunks at most: the producer issues x(t+1) only after the last codebook
        // piece of tile t, the ring frees slots only as the MMA consumes them, and the MMA can run two chunks ahead of
        // the epilogue (two TMEM buffers) -- with more chunks the wait for x(t+1) at the top of tile t would close a
        // cycle (measured: it does; B200, K = 1024).  Measured at N = 2^20, K = 256, D = 64: 0.719 -> 0.658 ms.
        constexpr bool PIPE_OK = !RESIDENT && PASSES == 3 && XS == 2;
        const bool pipe = PIPE_OK && p.num_chunks <= 2 && (p.flags & 0x40000000u) != 0;
        float xx_next = 0.f;
        if constexpr (PIPE_OK) {
            if (pipe && (int)blockIdx.x < p.num_tiles) xx_next = prep_tile<KB, XS>(sX, sXlo, x_full, xlo_full, r, lane, x_it);
        }
        for (int tile = blockIdx.x; tile < tile_end; tile += gridDim.x) {
            const uint32_t xs = x_it % XS, xph = (x_it / XS) & 1;
            const uint32_t xls = xs;
            uint8_t* sXt = sX + (size_t)xs * KB * XBLK;
            uint8_t* sXl = sXlo + (size_t)xls * KB * XBLK;
            VQB_TL(2);
            float xx_pipe = 0.f;
            bool prepped = false;
            if constexpr (PIPE_OK) {
                if (pipe) {
                    xx_pipe = xx_next;                              // this tile was prepared one iteration ago
                    if (tile + (int)gridDim.x < p.num_tiles) xx_next = prep_tile<KB, XS>(sX, sXlo, x_full, xlo_full, r, lane, x_it + 1);
                    prepped = true;
                }
            }
            if (!PIPE_OK || !prepped) mbar_wait(&x_full[xs], xph);
            VQB_TL(3);
            const int row0 = tile * BM;
            const int rows = min(BM, p.N - row0);
            const bool valid = r < rows;
            // |x|^2 in the exact kernel's fmaf order; x_lo = x - trunc_tf32(x) for the third MMA pass
            float xx = 0.f;
            if (PIPE_OK && prepped) xx = xx_pipe;
            float u_row = 1.f;                                      // F16: accumulator -> -2 x.e
            if constexpr (F16) {
                // pass 1: |x|^2 and the row maximum; pass 2: rescale, convert to fp16 in place (block j <- raw blocks 2j, 2j+1)
                float mx = 0.f;
#pragma unroll 1
                for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 v4 = *reinterpret_cast<const float4*>(sXt + kb * XBLK + sw128_offset(r, c));
                        xx = fmaf(v4.x, v4.x, xx); xx = fmaf(v4.y, v4.y, xx); xx = fmaf(v4.z, v4.z, xx); xx = fmaf(v4.w, v4.w, xx);
                        mx = fmaxf(fmaxf(mx, fmaxf(fabsf(v4.x), fabsf(v4.y))), fmaxf(fabsf(v4.z), fabsf(v4.w)));
                    }
                }
                const int er = scale_exp(mx);
                const float sx = pow2i(-er);
                u_row = -2.f * pow2i(er) * pow2i(__ldg(p.gexp));
#pragma unroll 1
                for (int j = 0; j < KH; ++j) {
                    float4 xv[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c)                    // all loads first: the stores below alias raw block j
                        xv[c] = *reinterpret_cast<const float4*>(sXt + (2 * j + (c >> 3)) * XBLK + sw128_offset(r, c & 7));
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float4 a4 = xv[2 * c], b4 = xv[2 * c + 1];
                        const uint4 h = make_uint4(pack_h2(a4.x * sx, a4.y * sx), pack_h2(a4.z * sx, a4.w * sx),
                                                   pack_h2(b4.x * sx, b4.y * sx), pack_h2(b4.z * sx, b4.w * sx));
                        *reinterpret_cast<uint4*>(sXt + j * XBLK + sw128_offset(r, c)) = h;
                    }
                }
                fence_proxy_async_smem();                           // generic writes -> tcgen05.mma operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&xlo_full[xls]);
            }
#pragma unroll 1
            for (int kb = 0; kb < ((F16 || (PIPE_OK && prepped)) ? 0 : KB); ++kb) {
                float4 xv[8];
#pragma unroll
                for (int c = 0; c < 8; ++c)                         // all loads first: the stores below may alias
                    xv[c] = *reinterpret_cast<const float4*>(sXt + kb * XBLK + sw128_offset(r, c));
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    xx = fmaf(xv[c].x, xv[c].x, xx); xx = fmaf(xv[c].y, xv[c].y, xx);
                    xx = fmaf(xv[c].z, xv[c].z, xx); xx = fmaf(xv[c].w, xv[c].w, xx);
                }
                if (PASSES == 3) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4 lo;
                        lo.x = xv[c].x - tf32_trunc(xv[c].x); lo.y = xv[c].y - tf32_trunc(xv[c].y);
                        lo.z = xv[c].z - tf32_trunc(xv[c].z); lo.w = xv[c].w - tf32_trunc(xv[c].w);
                        *reinterpret_cast<float4*>(sXl + kb * XBLK + sw128_offset(r, c)) = lo;
                    }
                }
            }
            if (PASSES == 3 && (!PIPE_OK || !prepped)) {
                fence_proxy_async_smem();                           // generic writes -> tcgen05.mma operand reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&xlo_full[xls]);
            }
            VQB_TL(4);

            int best = 0;
            {
                // ---------------- running minimum + candidate list over the codebook chunks ----------------
                // |approx - exact| <= eps for every code of this row (see header), so the exact arg-min lies
                // among the codes whose approximate value is within W = 2 eps of the approximate minimum.
                // Fast path per 32 columns: one min-reduction and one compare; only batches that reach the
                // running threshold append (value, code) pairs to the row's list in shared memory.
                const float xn = sqrtf(xx);
                const float rel = PASSES == 3 ? 6.0f * 9.5367431640625e-7f : 3.0f * 9.765625e-4f;   // 6*2^-20 | 3*2^-10
                const float W = 2.f * (1.02f * rel * xn * emax + 2e-6f * (xn + emax) * (xn + emax));
                float mn = INFINITY, thr = INFINITY;
                int cnt = 0;
                bool overflow = false;
                uint2* sCandG = sCand + (size_t)wg * CAP * BM;           // this warpgroup's list
                if (EW == 2 && wg == 1 && x_it > 0)
                    asm volatile("bar.sync 4, 256;" ::: "memory");        // the primary has read the previous tile's list
                // NOAUG: |e|^2 of the chunk's codes, staged by the row threads themselves (thread et <-> code), double-buffered;
                // the value of the next chunk is fetched one iteration ahead.  Codes beyond K get +1e30 (never the minimum).
                float* sEn = reinterpret_cast<float*>(sAfterBars + 2 * BM);     // [2][BN]
                float en_next = 0.f;
                if (NOAUG) en_next = et < p.K ? __ldg(p.bias + et) : 1e30f;
                // one 32-column batch of a chunk's accumulator: minimum, window test, candidate list (both warpgroups)
                auto scan_batch = [&](const int c, const uint32_t buf, const int chunk) {
                        float v[32];
                        tmem_ld_32x32(tmem_base + lane_addr + buf * BN + c * 32, v);
                        const int col0 = chunk * BN + c * 32;
                        if (NOAUG) {
                            const float4* en4 = reinterpret_cast<const float4*>(sEn + (chunk & 1) * BN + c * 32);
#pragma unroll
                            for (int j4 = 0; j4 < 8; ++j4) {
                                const float4 e4 = en4[j4];                  // same address in every lane: a broadcast
                                if constexpr (F16) {
                                    v[4 * j4] = fmaf(v[4 * j4], u_row, e4.x); v[4 * j4 + 1] = fmaf(v[4 * j4 + 1], u_row, e4.y);
                                    v[4 * j4 + 2] = fmaf(v[4 * j4 + 2], u_row, e4.z); v[4 * j4 + 3] = fmaf(v[4 * j4 + 3], u_row, e4.w);
                                } else {
                                    v[4 * j4] += e4.x; v[4 * j4 + 1] += e4.y; v[4 * j4 + 2] += e4.z; v[4 * j4 + 3] += e4.w;
                                }
                            }
                        }
                        // batch minimum as a tree, not a 31-long dependent chain; its second level -- eight minima over the
                        // column groups {g, g + 8, g + 16, g + 24} -- is kept: the candidate scan below only opens the groups
                        // that hold something inside the window
                        float bm, t8[8];
                        {
                            float t16[16];
#pragma unroll
                            for (int j = 0; j < 16; ++j) t16[j] = fminf(v[j], v[j + 16]);
#pragma unroll
                            for (int j = 0; j < 8; ++j) t8[j] = fminf(t16[j], t16[j + 8]);
                            bm = fminf(fminf(fminf(t8[0], t8[4]), fminf(t8[2], t8[6])), fminf(fminf(t8[1], t8[5]), fminf(t8[3], t8[7])));
                        }
                        if (bm <= thr) {
                            mn = fminf(mn, bm);
                            thr = mn + W;
                            if (cnt > 0 && cnt >= cap - 4) {
                                // make room: drop entries that fell out of the (tighter) window
                                int kept = 0;
                                for (int c2 = 0; c2 < cnt; ++c2) {
                                    const uint2 e = sCandG[c2 * BM + r];
                                    if (__uint_as_float(e.x) <= thr) sCandG[(kept++) * BM + r] = e;
                                }
                                cnt = kept;
                            }
                            // (the order of the list does not matter: the re-rank breaks ties by code index)
                            // Which groups hold something: a bit mask first, then one loop turn per set bit -- a jump table
                            // picks the group's four registers and the pushes are predicated, so a batch costs a handful of
                            // branches instead of eight group tests plus four value tests per group.
                            unsigned gm = 0;
#pragma unroll
                            for (int g = 0; g < 8; ++g) gm |= t8[g] <= thr ? (1u << g) : 0u;
                            while (gm) {
                                const int g = __ffs(gm) - 1;
                                gm &= gm - 1;
                                float a0, a1, a2, a3;
                                switch (g) {
                                    case 0: a0 = v[0]; a1 = v[8]; a2 = v[16]; a3 = v[24]; break;
                                    case 1: a0 = v[1]; a1 = v[9]; a2 = v[17]; a3 = v[25]; break;
                                    case 2: a0 = v[2]; a1 = v[10]; a2 = v[18]; a3 = v[26]; break;
                                    case 3: a0 = v[3]; a1 = v[11]; a2 = v[19]; a3 = v[27]; break;
                                    case 4: a0 = v[4]; a1 = v[12]; a2 = v[20]; a3 = v[28]; break;
                                    case 5: a0 = v[5]; a1 = v[13]; a2 = v[21]; a3 = v[29]; break;
                                    case 6: a0 = v[6]; a1 = v[14]; a2 = v[22]; a3 = v[30]; break;
                                    default: a0 = v[7]; a1 = v[15]; a2 = v[23]; a3 = v[31]; break;
                                }
                                auto push = [&](const float val, const int col) {
                                    const bool hit = val <= thr, room = cnt < cap;
                                    if (hit && room) sCandG[cnt * BM + r] = make_uint2(__float_as_uint(val), (unsigned)col);
                                    cnt += (hit && room) ? 1 : 0;
                                    overflow = overflow || (hit && !room);
                                };
                                push(a0, col0 + g); push(a1, col0 + g + 8); push(a2, col0 + g + 16); push(a3, col0 + g + 24);
                            }
                        }
                                    };
                for (int chunk = 0; chunk < p.num_chunks; ++chunk) {
                    const uint32_t buf = c_it % NB, tph = (c_it / NB) & 1;
                    if (NOAUG) {
                        sEn[(chunk & 1) * BN + et] = en_next;
                        const int nk = (chunk + 1) * BN + et;
                        en_next = (chunk + 1 < p.num_chunks && nk < p.K) ? __ldg(p.bias + nk) : 1e30f;
                        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // chunk's enorm visible; the buffer written
                    }                                                                  // now was last read two chunks ago
                    mbar_wait(&t_full[buf], tph);
                    tcgen05_fence_after();
#pragma unroll 1
                    for (int c = 0; c < BN / 32 / EW; ++c) scan_batch(c + wg * (BN / 32 / EW), buf, chunk);
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&t_empty[buf]);
                    ++c_it;
                }
                if constexpr (F16) {
                    // every MMA of this tile has retired (the last t_full was consumed above): the fp16 copy is dead, the raw
                    // fp32 tile comes back for the exact re-rank, the gather and the straight-through
                    if (et == 0) {
                        mbar_arrive_expect_tx(&xr_full[xs], KB * XBLK);
                        for (int kb = 0; kb < KB; ++kb) tma_load_2d(sXt + kb * XBLK, &tm_x, kb * 32, row0, &xr_full[xs]);
                    }
                    mbar_wait(&xr_full[xs], xph);
                }
                int ctot1 = 0;                                              // entries in the second warpgroup's list
                if constexpr (EW == 2) {
                    if (wg == 1) {
                        // second column half: publish minimum, count and overflow; the primary merges, re-ranks and stores
                        sMerge[r] = make_int2(__float_as_int(mn), cnt | (overflow ? 0x10000 : 0));
                        asm volatile("bar.arrive 3, 256;" ::: "memory");
                        ++x_it;
                        continue;
                    }
                    asm volatile("bar.sync 3, 256;" ::: "memory");
                    const int2 o = sMerge[r];
                    mn = fminf(mn, __int_as_float(o.x));
                    thr = mn + W;                                           // the window of the row's overall approximate minimum
                    ctot1 = o.y & 0xffff;
                    overflow = overflow || (o.y & 0x10000) != 0;
                }
                const int ctot = cnt;
                // survivors of the final window -> exact fp32 re-rank (same expression / fmaf order as the SIMT kernel)
                int ncand = 0;
                float bs_ = -INFINITY;
                int first = 0;
                for (int c2 = 0; c2 < ctot + ctot1; ++c2) {
                    const uint2 e = sCand[(c2 < ctot ? c2 : CAP + c2 - ctot) * BM + r];
                    if (__uint_as_float(e.x) <= thr) { if (ncand == 0) first = (int)e.y; ++ncand; }
                }
                const bool full_scan = valid && overflow;
                const bool rerank = valid && !overflow && ncand > 1;
                best = first;
                if (full_scan) {
                    for (int k = 0; k < p.K; ++k) {
                        const float sc = exact_score<KB>(sXt, r, xx, p.table, p.bias, k, tau);
                        if (sc > bs_) { bs_ = sc; best = k; }
                    }
                } else if (rerank) {
                    for (int c2 = 0; c2 < ctot + ctot1; ++c2) {
                        const uint2 e = sCand[(c2 < ctot ? c2 : CAP + c2 - ctot) * BM + r];
                        if (__uint_as_float(e.x) <= thr) {
                            const int k = (int)e.y;
                            const float sc = exact_score<KB>(sXt, r, xx, p.table, p.bias, k, tau);
                            if (sc > bs_ || (sc == bs_ && k < best)) { bs_ = sc; best = k; }
                        }
                    }
                }
                if (EW == 2) asm volatile("bar.arrive 4, 256;" ::: "memory");   // the second list may be overwritten
                if (p.stats) {
                    const unsigned m1 = __ballot_sync(0xffffffffu, rerank), m2 = __ballot_sync(0xffffffffu, full_scan);
                    if (lane == 0) {
                        if (m1) atomicAdd(p.stats, (unsigned)__popc(m1));
                        if (m2) atomicAdd(p.stats + 1, (unsigned)__popc(m2));
                    }
                }
            }

            VQB_TL(6);
            // ---- gather + straight-through, in place over the x tile -------------------------------------------
            // Coalesced mapping: 16 consecutive threads handle the 16 sixteen-byte chunks of one row, so the
            // codeword reads (all chunks of ONE code row) and the tile accesses are free of bank conflicts.
            int* sIdxG = sAfterBars;
            sIdxG[r] = valid ? best : -1;
            if (valid) p.idx[row0 + r] = best;
            VQB_TL(11);
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
            VQB_TL(12);
            {
                const bool skip = (p.flags & VQB_SKIP) != 0;
                // L2 score with the codebook resident in shared memory: e = -(hi + lo) / 2 exactly (hi + lo == -2 e)
                constexpr bool SMEM_GATHER = RESIDENT && PASSES == 3;
                // (VQB_GATHER_LDG=1 in the environment, developer A/B: read the codeword from the L1-resident fp32 table with
                //  one 128-bit load instead of reconstructing it from the two operand pieces in shared memory)
                const bool from_smem = SMEM_GATHER && !linear && !(p.flags & 0x80000000u);
                const bool want_se = p.sqerr != nullptr;
                constexpr int D4 = KB * 8;                          // 16-byte chunks per row
                constexpr int ITER = BM * D4 / 128;                 // chunks per thread
#pragma unroll 1
                for (int j0 = 0; j0 < ITER; j0 += 4) {
                    float4 xv[4], cv[4];
                    int rr[4], cc[4], code[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {                   // all loads first: the stores below may alias
                        const int i = et + 128 * (j0 + u);
                        rr[u] = i / D4; cc[u] = i % D4;
                        code[u] = sIdxG[rr[u]];
                        xv[u] = *reinterpret_cast<const float4*>(sXt + (cc[u] >> 3) * XBLK + sw128_offset(rr[u], cc[u] & 7));
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int k = code[u] < 0 ? 0 : code[u];
                        if (from_smem) {
                            const float4 h = *reinterpret_cast<const float4*>(sB + (size_t)(2 * (cc[u] >> 3)) * PIECE + sw128_offset(k, cc[u] & 7));
                            const float4 l = *reinterpret_cast<const float4*>(sB + (size_t)(2 * (cc[u] >> 3) + 1) * PIECE + sw128_offset(k, cc[u] & 7));
                            cv[u] = make_float4(-0.5f * (h.x + l.x), -0.5f * (h.y + l.y), -0.5f * (h.z + l.z), -0.5f * (h.w + l.w));
                        } else {
                            cv[u] = ldg4(p.gtab + (size_t)k * p.D + 4 * cc[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        float4 o;
                        if (linear) {
                            o = cv[u];                                                      // (:194-197)
                        } else {
                            // new_latent = enc_embs + picked_code - enc_embs.detach()  (:145)
                            o.x = __fsub_rn(__fadd_rn(xv[u].x, cv[u].x), xv[u].x); o.y = __fsub_rn(__fadd_rn(xv[u].y, cv[u].y), xv[u].y);
                            o.z = __fsub_rn(__fadd_rn(xv[u].z, cv[u].z), xv[u].z); o.w = __fsub_rn(__fadd_rn(xv[u].w, cv[u].w), xv[u].w);
                            if (skip) o = xv[u];                                            // (:142)
                        }
                        if (want_se && code[u] >= 0) {
                            const float d0 = xv[u].x - cv[u].x, d1 = xv[u].y - cv[u].y, d2 = xv[u].z - cv[u].z, d3 = xv[u].w - cv[u].w;
                            se_acc = fmaf(d0, d0, se_acc); se_acc = fmaf(d1, d1, se_acc);
                            se_acc = fmaf(d2, d2, se_acc); se_acc = fmaf(d3, d3, se_acc);
                        }
                        *reinterpret_cast<float4*>(sXt + (cc[u] >> 3) * XBLK + sw128_offset(rr[u], cc[u] & 7)) = o;
                    }
                }
            }
            VQB_TL(13);
            if (p.hist) {
                // warp-aggregated histogram: one atomic per distinct code per warp
                const unsigned peers = __match_any_sync(0xffffffffu, valid ? best : -1);
                if (valid && lane == (__ffs(peers) - 1)) atomicAdd(p.hist + best, (unsigned long long)__popc(peers));
            }
            VQB_TL(7);
            fence_proxy_async_smem();                               // this thread's tile / p_code writes -> async proxy
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // this warpgroup only
            if (et == 0) {
                // new_latent tile: TMA store straight from the swizzled tile (rows beyond N are clipped by TMA)
#pragma unroll
                for (int kb = 0; kb < KB; ++kb) tma_store_2d(&tm_q, sXt + kb * XBLK, kb * 32, row0);
                tma_store_commit();
            }
            VQB_TL(8);
            if (et == 0) tma_store_wait_read();                     // shared memory may be overwritten from here on
            VQB_TL(9);
            asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");   // sP / x tile fully drained before reuse
            if (lane == 0) mbar_arrive(&x_empty[xs]);               // the x slot may be refilled by TMA
            ++x_it;
        }
        if (p.sqerr) {
            se_acc = warp_sum(se_acc);
            if (lane == 0) atomicAdd(p.sqerr, (double)se_acc);
        }
        if (et == 0) tma_store_wait_all();
        VQB_TL(10);
    }

    // ---- teardown ----------------------------------------------------------------------------------------
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 64) p.dbg[123] = globaltimer_ns();
}

// -----------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 tmap_encoder() {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            set_error("libvqb200: cuTensorMapEncodeTiled is not available from the driver");
            return nullptr;
        }
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    return encode;
}

int make_tmap_2d_plain_f32(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes,
                           uint32_t box0, uint32_t box1) {
    auto encode = tmap_encoder();
    if (!encode) return VQB_ERR_CUDA;
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("libvqb200: cuTensorMapEncodeTiled (plain) failed with CUresult %d (dims %llu x %llu, box %u x %u)", (int)r,
                  (unsigned long long)dim0, (unsigned long long)dim1, box0, box1);
        return VQB_ERR_CUDA;
    }
    return VQB_OK;
}

int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                     uint32_t box_rows, bool atom32b) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            set_error("libvqb200: cuTensorMapEncodeTiled is not available from the driver");
            return VQB_ERR_CUDA;
        }
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_elems * sizeof(float)};
    cuuint32_t box[2] = {32, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("libvqb200: cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols);
        return VQB_ERR_CUDA;
    }
    return VQB_OK;
}

// 2-D fp16 row-major tensor map [rows][cols], box = box_rows x 64 halves (128 bytes), SWIZZLE_128B
static int make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
            cudaGetLastError();
            set_error("libvqb200: cuTensorMapEncodeTiled is not available from the driver");
            return VQB_ERR_CUDA;
        }
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("libvqb200: cuTensorMapEncodeTiled (fp16) failed with CUresult %d (rows=%llu cols=%llu)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols);
        return VQB_ERR_CUDA;
    }
    return VQB_OK;
}

static int g_search_pipe = -1;                 // -1: default (on unless VQB_SEARCH_NOPIPE); 0 / 1: forced (vqb_debug_set_search_pipe)
void set_debug_search_pipe(int v) { g_search_pipe = v; }
static unsigned long long* g_timeline = nullptr;
void set_debug_timeline(void* p) { g_timeline = reinterpret_cast<unsigned long long*>(p); }
unsigned long long* get_debug_timeline() { return g_timeline; }

static int64_t pad_codes(int64_t K, int64_t bn) { return (K + bn - 1) / bn * bn; }
// tf32 hi / lo operand copies of the codebook (workspace): hi [Kpad][D + 32] (the last block carries the bias), lo [Kpad][D]
static size_t hi_bytes(int64_t K, int64_t D) { return ((size_t)pad_codes(K, 128) * (D + 32) * 4 + 255) & ~(size_t)255; }
static size_t lo_bytes(int64_t K, int64_t D) { return ((size_t)pad_codes(K, 128) * D * 4 + 255) & ~(size_t)255; }

// which kernel configuration serves this call (0 = none: use the exact SIMT path)
enum TcMode { TC_NONE = 0, TC_SEARCH3, TC_SEARCH1, TC_SEARCH16 };
static TcMode tc_mode(const vqb_fwd_args* a) {
    const int64_t K = a->n_codes, D = a->dim;
    if (a->p_code) return TC_NONE;                               // parity mode: vqb_fwd_pc.cu
    if (!(a->flags & VQB_SCORE_L2)) return TC_NONE;
    if (D != 32 && D != 64 && D != 128 && D != 256) return TC_NONE;
    if (K <= 1024 && D <= 128) return TC_SEARCH3;
    // one pass + candidate window + exact re-rank.  fp16 operands at D = 256, where the tf32 kernel is bound by the codebook
    // bytes in flight (K = 8192: 9.6 vs 11.5 ms, 459 vs 381 TFLOP/s); tf32 at D <= 128, where the single epilogue warpgroup is
    // the bound and the in-place conversion + second fetch of the x tile only add to it (K = 8192, D = 64: 7.7 vs 6.6 ms) --
    // profiles/r2_sweep_f16_vs_tf32.txt.  Developer A/B: VQB_SEARCH_TF32=1 / VQB_SEARCH_F16=1 force one or the other.
    static const bool tf32_only = getenv("VQB_SEARCH_TF32") != nullptr, f16_all = getenv("VQB_SEARCH_F16") != nullptr;
    if (tf32_only || D % 64 != 0) return TC_SEARCH1;
    return (D == 256 || f16_all) ? TC_SEARCH16 : TC_SEARCH1;
}

bool forward_tensor_supported(const vqb_fwd_args* a) { return tc_mode(a) != TC_NONE; }

int forward_tensor_workspace(const vqb_fwd_args* a, size_t* bytes) {
    *bytes = tc_mode(a) == TC_NONE ? 0 : hi_bytes(a->n_codes, a->dim) + lo_bytes(a->n_codes, a->dim) + 256;
    return VQB_OK;
}

template <int KB, int BN, int XS, int BS, int PASSES, bool RESIDENT, bool NOAUG = false, bool F16 = false, int EW = 1>
static int launch_tc(const CUtensorMap& tx, const CUtensorMap& th, const CUtensorMap& tl, const CUtensorMap& tq, const TcP& p,
                     cudaStream_t s) {
    const size_t smem = (size_t)XS * KB * XBLK + (PASSES == 3 ? (size_t)XS * KB * XBLK : 0) + (NOAUG ? 0 : XBLK) +
                        (size_t)BS * BN * 128 + (size_t)EW * (NOAUG ? 15 : 16) * BM * 8 + 1024 + 320 + 2 * BM * 4 + (NOAUG ? 2 * BN * 4 : 0);
    if ((int)smem > max_optin_smem()) return invalid("vqb_forward: tensor-core configuration needs %zu B of shared memory", smem);
    auto kern = vqb_fwd_tc_kernel<KB, BN, XS, BS, PASSES, RESIDENT, NOAUG, F16, EW>;
    VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    // PDL: this call has just launched build_operands_kernel; the prologue and the first x tile overlap it
    kernel_event_begin(s);
    VQB_CUDA(launch_pdl(kern, dim3(grid), dim3(64 + 128 * EW), smem, s, tx, th, tl, tq, p));
    kernel_event_end(s);
    VQB_CHECK_LAUNCH("vqb_fwd_tc_kernel");
    return VQB_OK;
}

int launch_forward_tensor(const vqb_fwd_args* a, cudaStream_t s) {
    const int64_t N = a->n_rows, K = a->n_codes, D = a->dim;
    if (N == 0) return VQB_OK;
    const TcMode mode = tc_mode(a);
    if (mode == TC_NONE) return invalid("vqb_forward: shape not supported by the tensor-core search");
    const size_t need = hi_bytes(K, D) + lo_bytes(K, D) + 256;
    if (!a->workspace || a->workspace_bytes < need) {
        set_error("vqb_forward: workspace too small (%zu < %zu bytes)", a->workspace_bytes, need);
        return VQB_ERR_WORKSPACE;
    }
    uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
    float* hi = reinterpret_cast<float*>(ws);
    float* lo = reinterpret_cast<float*>(ws + hi_bytes(K, D));
    uint8_t* tail = ws + hi_bytes(K, D) + lo_bytes(K, D);
    float* emax = reinterpret_cast<float*>(tail);
    unsigned int* stats = a->search_stats ? a->search_stats : reinterpret_cast<unsigned int*>(tail + 16);
    VQB_CUDA(cudaMemsetAsync(tail, 0, 256, s));                  // |e|_max / table maximum (atomicMax) and the re-rank counters
    constexpr int BN = 128;
    const int64_t Kpad = pad_codes(K, BN);
    int* gexp = reinterpret_cast<int*>(tail + 32);
    unsigned int* gmax_bits = reinterpret_cast<unsigned int*>(tail + 36);
    CUtensorMap tx, th, tl, tq;
    int rc;
    if (mode == TC_SEARCH16) {
        table_max_kernel<<<(unsigned)K, 128, 0, s>>>(a->score_w, (int)D, gmax_bits, emax);
        VQB_CHECK_LAUNCH("table_max_kernel");
        build_f16_operands_kernel<<<(unsigned)Kpad, 128, 0, s>>>(a->score_w, (int)K, (int)D, gmax_bits, reinterpret_cast<__half*>(hi), gexp);
        VQB_CHECK_LAUNCH("build_f16_operands_kernel");
        if ((rc = make_tmap_2d_f16(&th, hi, (uint64_t)Kpad, (uint64_t)D, BN))) return rc;
        tl = th;
    } else {
        launch_build_operands(a->score_w, a->score_b, (int)K, (int)Kpad, (int)D, -2.f, 1e30f, hi, mode == TC_SEARCH1 ? nullptr : lo, emax, s);
        VQB_CHECK_LAUNCH("build_operands_kernel");
        if ((rc = make_tmap_2d_f32(&th, hi, (uint64_t)Kpad, (uint64_t)(D + 32), (uint64_t)(D + 32), BN))) return rc;
        if ((rc = make_tmap_2d_f32(&tl, lo, (uint64_t)Kpad, (uint64_t)D, (uint64_t)D, BN))) return rc;
    }
    if ((rc = make_tmap_2d_f32(&tx, a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, BM))) return rc;
    if ((rc = make_tmap_2d_f32(&tq, a->new_latent, (uint64_t)N, (uint64_t)D, (uint64_t)D, BM))) return rc;

    TcP p;
    p.table = a->score_w; p.gtab = a->gather_table; p.bias = a->score_b; p.temp = a->temp; p.emax = emax; p.gexp = gexp;
    p.idx = (long long*)a->idx; p.q = a->new_latent;
    p.hist = (unsigned long long*)a->hist; p.sqerr = a->sq_err_sum; p.stats = stats; p.dbg = g_timeline;
    p.N = (int)N; p.K = (int)K; p.D = (int)D;
    p.num_tiles = (int)ceil_div(N, BM); p.num_chunks = (int)ceil_div(K, BN);
    p.flags = a->flags;
    { static const bool ldg = getenv("VQB_GATHER_LDG") != nullptr; if (ldg) p.flags |= 0x80000000u; }
    // software-pipelined x_lo in the streamed 3xTF32 search (see PIPE in the kernel): on by default,
    // VQB_SEARCH_NOPIPE=1 or vqb_debug_set_search_pipe(0) turns it off (developer A/B)
    { static const bool nopipe_env = getenv("VQB_SEARCH_NOPIPE") != nullptr; if (g_search_pipe < 0 ? !nopipe_env : g_search_pipe > 0) p.flags |= 0x40000000u; }

    //                      KB  BN  XS BS PASSES RESIDENT
    if (mode == TC_SEARCH3) {
        if (K <= 128) {                                            // whole codebook resident in shared memory
            if (D == 32) return launch_tc<1, 128, 2, 3, 3, true>(tx, th, tl, tq, p, s);
            if (D == 64) return launch_tc<2, 128, 1, 5, 3, true>(tx, th, tl, tq, p, s);
        }
        // The codebook ring is as deep as shared memory allows: the streamed search is bound by the bytes in flight
        // from L2 (one 16 KB piece per slot; the tensor pipe consumes ~100 GB/s per SM at the tf32 peak, i.e. needs
        // ~150 KB in flight at ~1.5 us of loaded L2 latency -- profiles/r1e_ncu_full_search.csv: nothing else is busy).
        if (D == 32) return launch_tc<1, 128, 2, 8, 3, false>(tx, th, tl, tq, p, s);
        if (D == 64) return launch_tc<2, 128, 2, 4, 3, false>(tx, th, tl, tq, p, s);
        return launch_tc<4, 128, 1, 4, 3, false>(tx, th, tl, tq, p, s);
    }
    if (mode == TC_SEARCH16) {
        //                                KB  BN  XS BS PASSES RESIDENT NOAUG F16
        if (D == 64)  return launch_tc<2, 128, 2, 8, 1, false, true, true>(tx, th, tl, tq, p, s);
        if (D == 128) return launch_tc<4, 128, 1, 8, 1, false, true, true>(tx, th, tl, tq, p, s);
        return launch_tc<8, 128, 1, 5, 1, false, true, true>(tx, th, tl, tq, p, s);
    }
    // D <= 128: two epilogue warpgroups (EW = 2); the second candidate list costs one ring slot where shared memory is full.
    // VQB_SEARCH_EW1=1 (developer A/B): the single-warpgroup kernels
    static const bool ew1 = getenv("VQB_SEARCH_EW1") != nullptr;
    if (!ew1) {
        //                                KB  BN  XS BS PASSES RESIDENT NOAUG  F16   EW
        if (D == 32)  return launch_tc<1, 128, 2, 8, 1, false, false, false, 2>(tx, th, tl, tq, p, s);
        if (D == 64)  return launch_tc<2, 128, 2, 7, 1, false, false, false, 2>(tx, th, tl, tq, p, s);
        if (D == 128) return launch_tc<4, 128, 1, 7, 1, false, false, false, 2>(tx, th, tl, tq, p, s);
    }
    switch (D) {
        case 32:  return launch_tc<1, 128, 2, 8, 1, false>(tx, th, tl, tq, p, s);
        case 64:  return launch_tc<2, 128, 2, 8, 1, false>(tx, th, tl, tq, p, s);
        case 128: return launch_tc<4, 128, 1, 8, 1, false>(tx, th, tl, tq, p, s);
        default:  return launch_tc<8, 128, 1, 5, 1, false, true>(tx, th, tl, tq, p, s);
    }
}

}  // namespace vqb
