// Fused-mode forward on the 5th-generation tensor cores: nearest-codeword search + exact re-rank +
// gather + straight-through, without ever writing the N x K distance matrix.
//
// Replaces (for the no-p_code mode) neg_batch_l2 + argmax + F.embedding + straight-through of
// src/embed.py:208-213, :130, :134, :145.
//
// Kernel anatomy (one persistent CTA per SM, 6 warps, warp-specialised):
//   warp 0  TMA producer   x tile [128 rows][D] fp32 (resident for the whole codebook sweep) and the
//                          codebook streamed as [128 codes][32 floats] K-blocks through a 4-deep ring
//   warp 1  MMA issuer     tcgen05.mma kind::tf32, M=128 N=128 K=8, operands straight from the fp32
//                          tiles (the tensor core reads the top 19 bits), accumulator in TMEM,
//                          double-buffered (2 x 128 columns) so the epilogue of chunk c overlaps chunk c+1
//   warps 2-5 epilogue     tcgen05.ld 32x32b: thread = row, so the running top-4 over the codebook is
//                          thread-local (no shuffles); then the provable candidate window, the exact
//                          fp32 re-rank in the reference's evaluation order, gather and (x + c) - x.
// The |e|^2 bias is folded into the GEMM: the codebook operand is pre-scaled to -2*e and carries one
// extra K-step whose A side is the constant [1,1,1,0,...] and whose B side is |e|^2 split into three
// tf32-exact words, so the accumulator is directly |e|^2 - 2 x.e and the epilogue is one compare per
// element.
//
// Index exactness: tf32 truncates both operands (relative error < 2^-10 each), so
// |approx - exact| <= 2^-8 |x| |e|_max (+ fp32 accumulation slack) =: eps for every code of the row.
// Every code whose approximate distance is within 2*eps of the approximate minimum is re-evaluated in
// exact fp32 -- (|x|^2 + |e|^2) - 2 x.e with the same fmaf order as the exact SIMT kernel -- and the
// arg-max of relu(temp) * -dist (first index on ties) is taken.  If more than 3 codes fall inside the
// window the row falls back to a full exact scan.  Both counts are reported (search_stats).
#include <cudaTypedefs.h>
#include <math.h>
#include "vqb_common.cuh"
#include "vqb_tc.cuh"

namespace vqb {
using namespace tc;

constexpr int BM = 128;                 // rows per tile  (UMMA M)
constexpr int BN = 128;                 // codes per chunk (UMMA N)
constexpr int KBLK_BYTES = BM * 128;    // one K-block: [128][32 fp32] = 16 KB (BM == BN)
constexpr int TC_THREADS = 192;
constexpr int TMEM_COLS = 2 * BN;

struct SearchP {
    const float* table;        // [K][D] fp32: exact re-rank and gather
    const float* enorm;        // [K]
    const float* temp;         // [1]
    const float* emax;         // [1]  max_k |e_k|
    long long* idx;
    float* q;
    unsigned long long* hist;
    double* sqerr;
    unsigned int* stats;       // [0] rows re-ranked, [1] rows that needed the full exact scan
    int N, K, D, num_tiles, num_chunks;
    unsigned flags;
};

// Eaug[k] = [-2 e_k (D floats) | |e_k|^2 as hi, mid, lo tf32-exact words | 0 x 29];  emax = max |e_k|
__global__ void __launch_bounds__(128)
build_eaug_kernel(const float* __restrict__ table, const float* __restrict__ enorm, int K, int D,
                  float* __restrict__ eaug, float* __restrict__ emax) {
    const int k = blockIdx.x;
    float* row = eaug + (size_t)k * (D + 32);
    for (int d = threadIdx.x; d < D; d += blockDim.x) row[d] = -2.f * table[(size_t)k * D + d];
    if (threadIdx.x < 32) {
        const float ee = enorm[k];
        const float hi = __uint_as_float(__float_as_uint(ee) & 0xFFFFE000u);
        const float r1 = ee - hi;
        const float mid = __uint_as_float(__float_as_uint(r1) & 0xFFFFE000u);
        const float lo = r1 - mid;
        const int j = threadIdx.x;
        row[D + j] = j == 0 ? hi : (j == 1 ? mid : (j == 2 ? lo : 0.f));
        if (j == 0) atomicMax(reinterpret_cast<int*>(emax), __float_as_int(sqrtf(ee)));   // non-negative floats order as ints
    }
}

struct Cand { float v; int i; };

// exact score of code k for the row held (swizzled) in shared memory
template <int KB>
__device__ __forceinline__ float exact_score(const uint8_t* sXt, int r, float xx, const float* __restrict__ table,
                                             const float* __restrict__ enorm, int k, float tau) {
    const float* e = table + (size_t)k * (KB * 32);
    float dot = 0.f;
#pragma unroll 1
    for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float4 xv = *reinterpret_cast<const float4*>(sXt + kb * KBLK_BYTES + sw128_offset(r, c));
            const float4 w = ldg4(e + kb * 32 + c * 4);
            dot = fmaf(xv.x, w.x, dot); dot = fmaf(xv.y, w.y, dot);
            dot = fmaf(xv.z, w.z, dot); dot = fmaf(xv.w, w.w, dot);
        }
    }
    const float dist = __fsub_rn(__fadd_rn(xx, __ldg(enorm + k)), 2.f * dot);
    return tau * (-dist);
}

template <int KB, int XS, int BS>
__global__ void __launch_bounds__(TC_THREADS, 1)
vqb_search_tf32_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_e, SearchP p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* sX = smem;                                            // [XS][KB][16 KB]
    uint8_t* sAug = sX + (size_t)XS * KB * KBLK_BYTES;             // [16 KB] constant A block [1,1,1,0,...]
    uint8_t* sB = sAug + KBLK_BYTES;                               // [BS][16 KB]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)BS * KBLK_BYTES);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + XS;
    uint64_t* b_full = x_empty + XS;
    uint64_t* b_empty = b_full + BS;
    uint64_t* t_full = b_empty + BS;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup ----------------------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_e);
        for (int i = 0; i < XS; ++i) { mbar_init(&x_full[i], 1); mbar_init(&x_empty[i], 4); }
        for (int i = 0; i < BS; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<TMEM_COLS>(tmem_slot);
    for (int i = threadIdx.x; i < KBLK_BYTES / 16; i += TC_THREADS)
        reinterpret_cast<float4*>(sAug)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    if (threadIdx.x < BM)
        *reinterpret_cast<float4*>(sAug + sw128_offset(threadIdx.x, 0)) = make_float4(1.f, 1.f, 1.f, 0.f);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t IDESC = umma_idesc(2u, BM, BN);

    if (warp == 0) {
        // =============================== TMA producer ===============================================
        if (lane == 0) {
            uint32_t x_it = 0, b_it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const uint32_t xs = x_it % XS, xph = (x_it / XS) & 1;
                mbar_wait(&x_empty[xs], xph ^ 1);
                mbar_arrive_expect_tx(&x_full[xs], KB * KBLK_BYTES);
                for (int kb = 0; kb < KB; ++kb)
                    tma_load_2d(sX + ((size_t)xs * KB + kb) * KBLK_BYTES, &tm_x, kb * 32, tile * BM, &x_full[xs]);
                ++x_it;
                for (int chunk = 0; chunk < p.num_chunks; ++chunk) {
                    for (int kb = 0; kb <= KB; ++kb) {              // kb == KB: the |e|^2 block
                        const uint32_t bs = b_it % BS, bph = (b_it / BS) & 1;
                        mbar_wait(&b_empty[bs], bph ^ 1);
                        mbar_arrive_expect_tx(&b_full[bs], KBLK_BYTES);
                        tma_load_2d(sB + (size_t)bs * KBLK_BYTES, &tm_e, kb * 32, chunk * BN, &b_full[bs]);
                        ++b_it;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer =================================================
        if (lane == 0) {
            uint32_t x_it = 0, b_it = 0, c_it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const uint32_t xs = x_it % XS, xph = (x_it / XS) & 1;
                mbar_wait(&x_full[xs], xph);
                tcgen05_fence_after();
                for (int chunk = 0; chunk < p.num_chunks; ++chunk) {
                    const uint32_t buf = c_it & 1, tph = (c_it >> 1) & 1;
                    mbar_wait(&t_empty[buf], tph ^ 1);
                    tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * BN;
                    for (int kb = 0; kb <= KB; ++kb) {
                        const uint32_t bs = b_it % BS, bph = (b_it / BS) & 1;
                        mbar_wait(&b_full[bs], bph);
                        tcgen05_fence_after();
                        const uint64_t bdesc = umma_desc_sw128(sB + (size_t)bs * KBLK_BYTES);
                        if (kb < KB) {
                            const uint64_t adesc = umma_desc_sw128(sX + ((size_t)xs * KB + kb) * KBLK_BYTES);
#pragma unroll
                            for (int k = 0; k < 4; ++k)            // 4 x (K = 8 tf32 = 32 bytes) per 128-byte row
                                umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, IDESC, (kb | k) != 0);
                        } else {
                            umma_tf32(d_tmem, umma_desc_sw128(sAug), bdesc, IDESC, true);
                        }
                        umma_commit(&b_empty[bs]);                  // frees the ring slot when the MMAs retire
                        ++b_it;
                    }
                    umma_commit(&t_full[buf]);                      // accumulator of this chunk is complete
                    ++c_it;
                }
                ++x_it;
            }
        }
    } else {
        // =============================== epilogue (thread = row) ====================================
        const int q4 = warp & 3;                                    // TMEM lane quadrant this warp may read
        const int r = q4 * 32 + lane;                               // row within the tile == TMEM lane
        const int et = (warp - 2) * 32 + lane;                      // 0..127 among epilogue threads
        const float tau = fmaxf(__ldg(p.temp), 0.f);
        const float emax = __ldg(p.emax);
        uint32_t x_it = 0, c_it = 0;
        float se_acc = 0.f;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const uint32_t xs = x_it % XS, xph = (x_it / XS) & 1;
            uint8_t* sXt = sX + (size_t)xs * KB * KBLK_BYTES;
            mbar_wait(&x_full[xs], xph);
            const int row0 = tile * BM;
            const int rows = min(BM, p.N - row0);
            const bool valid = r < rows;
            float xx = 0.f;
#pragma unroll 1
            for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 xv = *reinterpret_cast<const float4*>(sXt + kb * KBLK_BYTES + sw128_offset(r, c));
                    xx = fmaf(xv.x, xv.x, xx); xx = fmaf(xv.y, xv.y, xx);
                    xx = fmaf(xv.z, xv.z, xx); xx = fmaf(xv.w, xv.w, xx);
                }
            }
            Cand t0 = {INFINITY, 0}, t1 = {INFINITY, 0}, t2 = {INFINITY, 0}, t3 = {INFINITY, 0};
            for (int chunk = 0; chunk < p.num_chunks; ++chunk) {
                const uint32_t buf = c_it & 1, tph = (c_it >> 1) & 1;
                mbar_wait(&t_full[buf], tph);
                tcgen05_fence_after();
                const bool full_chunk = (chunk + 1) * BN <= p.K;
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    float v[32];
                    tmem_ld_32x32(tmem_base + ((uint32_t)(q4 * 32) << 16) + buf * BN + c * 32, v);
                    const int col0 = chunk * BN + c * 32;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float val = (full_chunk || col0 + j < p.K) ? v[j] : INFINITY;
                        if (val < t3.v) {
                            const Cand n = {val, col0 + j};
                            if (val < t2.v) {
                                t3 = t2;
                                if (val < t1.v) {
                                    t2 = t1;
                                    if (val < t0.v) { t1 = t0; t0 = n; } else { t1 = n; }
                                } else { t2 = n; }
                            } else { t3 = n; }
                        }
                    }
                }
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&t_empty[buf]);
                ++c_it;
            }
            // ---- provable candidate window, exact fp32 re-rank -------------------------------------
            const float eps = 0.00390625f * 1.02f * sqrtf(xx) * emax + 1e-5f * (fabsf(t0.v) + xx);
            const float window = t0.v + 2.f * eps;
            int best = t0.i;
            const int ncand = 1 + (t1.v <= window) + (t2.v <= window) + (t3.v <= window);
            const bool full_scan = valid && ncand == 4 && p.K > 4;
            const bool rerank = valid && ncand > 1;
            if (full_scan) {
                float bs_ = -INFINITY;
                for (int k = 0; k < p.K; ++k) {
                    const float s = exact_score<KB>(sXt, r, xx, p.table, p.enorm, k, tau);
                    if (s > bs_) { bs_ = s; best = k; }
                }
            } else if (rerank) {
                float bs_ = exact_score<KB>(sXt, r, xx, p.table, p.enorm, t0.i, tau);
                const Cand cs[3] = {t1, t2, t3};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    if (c + 1 < ncand) {
                        const float s = exact_score<KB>(sXt, r, xx, p.table, p.enorm, cs[c].i, tau);
                        if (s > bs_ || (s == bs_ && cs[c].i < best)) { bs_ = s; best = cs[c].i; }
                    }
                }
            }
            if (p.stats) {
                const unsigned m1 = __ballot_sync(0xffffffffu, rerank), m2 = __ballot_sync(0xffffffffu, full_scan);
                if (lane == 0) {
                    if (m1) atomicAdd(p.stats, (unsigned)__popc(m1));
                    if (m2) atomicAdd(p.stats + 1, (unsigned)__popc(m2));
                }
            }
            // ---- gather + straight-through, staged in place over the x tile ------------------------
            if (valid) {
                const float* crow = p.table + (size_t)best * p.D;
#pragma unroll 1
                for (int kb = 0; kb < KB; ++kb) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float4* xp = reinterpret_cast<float4*>(sXt + kb * KBLK_BYTES + sw128_offset(r, c));
                        const float4 xv = *xp;
                        const float4 cv = ldg4(crow + kb * 32 + c * 4);
                        float4 o;
                        o.x = __fsub_rn(__fadd_rn(xv.x, cv.x), xv.x); o.y = __fsub_rn(__fadd_rn(xv.y, cv.y), xv.y);
                        o.z = __fsub_rn(__fadd_rn(xv.z, cv.z), xv.z); o.w = __fsub_rn(__fadd_rn(xv.w, cv.w), xv.w);
                        if (p.flags & VQB_SKIP) o = xv;
                        const float d0 = xv.x - cv.x, d1 = xv.y - cv.y, d2 = xv.z - cv.z, d3 = xv.w - cv.w;
                        se_acc = fmaf(d0, d0, se_acc); se_acc = fmaf(d1, d1, se_acc);
                        se_acc = fmaf(d2, d2, se_acc); se_acc = fmaf(d3, d3, se_acc);
                        *xp = o;
                    }
                }
                p.idx[row0 + r] = best;
            }
            if (p.hist) {
                // warp-aggregated histogram: one atomic per distinct code per warp
                const unsigned peers = __match_any_sync(0xffffffffu, valid ? best : -1);
                if (valid && lane == (__ffs(peers) - 1)) atomicAdd(p.hist + best, (unsigned long long)__popc(peers));
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");          // epilogue warps only
            constexpr int D4 = KB * 8;
            for (int i = et; i < rows * D4; i += 128) {
                const int rr = i / D4, c = i % D4;
                const float4 o = *reinterpret_cast<const float4*>(sXt + (c >> 3) * KBLK_BYTES + sw128_offset(rr, c & 7));
                stg4_stream(p.q + (size_t)(row0 + rr) * p.D + 4 * c, o);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&x_empty[xs]);               // the x slot may be refilled by TMA
            ++x_it;
        }
        if (p.sqerr) {
            se_acc = warp_sum(se_acc);
            if (lane == 0) atomicAdd(p.sqerr, (double)se_acc);
        }
    }

    // ---- teardown ----------------------------------------------------------------------------------
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_stride_elems,
                     uint32_t box_rows) {
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
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("libvqb200: cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols);
        return VQB_ERR_CUDA;
    }
    return VQB_OK;
}

static size_t eaug_bytes(int64_t K, int64_t D) { return ((size_t)K * (D + 32) * 4 + 255) & ~(size_t)255; }

int forward_tensor_workspace(const vqb_fwd_args* a, size_t* bytes) {
    *bytes = eaug_bytes(a->n_codes, a->dim) + 256;
    return VQB_OK;
}

template <int KB, int XS, int BS>
static int launch_search(const CUtensorMap& tx, const CUtensorMap& te, const SearchP& p, cudaStream_t s) {
    const size_t smem = (size_t)(XS * KB + 1 + BS) * KBLK_BYTES + 1024 + 256;
    auto kern = vqb_search_tf32_kernel<KB, XS, BS>;
    VQB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = p.num_tiles < sm_count() ? p.num_tiles : sm_count();
    kern<<<grid, TC_THREADS, smem, s>>>(tx, te, p);
    VQB_CHECK_LAUNCH("vqb_search_tf32_kernel");
    return VQB_OK;
}

int launch_forward_tensor(const vqb_fwd_args* a, cudaStream_t s) {
    const int64_t N = a->n_rows, K = a->n_codes, D = a->dim;
    if (N == 0) return VQB_OK;
    if (D != 32 && D != 64 && D != 128 && D != 256)
        return invalid("vqb_forward: the tensor-core search supports D in {32, 64, 128, 256} (got %lld)", (long long)D);
    const size_t need = eaug_bytes(K, D) + 256;
    if (!a->workspace || a->workspace_bytes < need) {
        set_error("vqb_forward: workspace too small (%zu < %zu bytes)", a->workspace_bytes, need);
        return VQB_ERR_WORKSPACE;
    }
    float* eaug = reinterpret_cast<float*>(a->workspace);
    uint8_t* tail = reinterpret_cast<uint8_t*>(a->workspace) + eaug_bytes(K, D);
    float* emax = reinterpret_cast<float*>(tail);
    unsigned int* stats = a->search_stats ? a->search_stats : reinterpret_cast<unsigned int*>(tail + 16);
    VQB_CUDA(cudaMemsetAsync(tail, 0, 256, s));
    build_eaug_kernel<<<(unsigned)K, 128, 0, s>>>(a->score_w, a->score_b, (int)K, (int)D, eaug, emax);
    VQB_CHECK_LAUNCH("build_eaug_kernel");

    CUtensorMap tx, te;
    int rc = make_tmap_2d_f32(&tx, a->x, (uint64_t)N, (uint64_t)D, (uint64_t)D, BM);
    if (rc) return rc;
    rc = make_tmap_2d_f32(&te, eaug, (uint64_t)K, (uint64_t)(D + 32), (uint64_t)(D + 32), BN);
    if (rc) return rc;

    SearchP p;
    p.table = a->gather_table; p.enorm = a->score_b; p.temp = a->temp; p.emax = emax;
    p.idx = (long long*)a->idx; p.q = a->new_latent; p.hist = (unsigned long long*)a->hist; p.sqerr = a->sq_err_sum;
    p.stats = stats;
    p.N = (int)N; p.K = (int)K; p.D = (int)D;
    p.num_tiles = (int)ceil_div(N, BM); p.num_chunks = (int)ceil_div(K, BN);
    p.flags = a->flags;
    switch (D) {
        case 32:  return launch_search<1, 2, 4>(tx, te, p, s);
        case 64:  return launch_search<2, 2, 4>(tx, te, p, s);
        case 128: return launch_search<4, 1, 4>(tx, te, p, s);
        default:  return launch_search<8, 1, 4>(tx, te, p, s);
    }
}

}  // namespace vqb
