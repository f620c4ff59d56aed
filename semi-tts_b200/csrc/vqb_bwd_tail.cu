// Behind the parity-mode backward kernel (vqb_bwd_pc.cu): the per-CTA partial records -> parameter gradients.
//   reduce_partials_kernel   fixed-order sum of the records into the caller's accumulators (plain vqb_backward semantics)
//   bwd_tail_kernel          the fused tail (vqb_bwd_tail): fixed-order sum + backward of the table assembly
//                            (src/embed.py:109-112) + -- in data-parallel runs -- the sum over GPUs as a one-shot
//                            all-reduce over NVLink peer memory, in ONE launch that is PDL-chained to the main kernel
// No atomics anywhere: gradients are bit-reproducible, and bit-identical on every rank.
#include <cudaTypedefs.h>
#include <math.h>
#include "vqb_common.cuh"
#include "vqb_tc.cuh"

namespace vqb {
using namespace tc;

constexpr int H_KD = 64 * 64;
constexpr int H_PARTIAL_FLOATS = 2 * H_KD + 64;   // record of one CTA: [0] d_score_w part, [1] scatter part (LINEAR) / transposed
                                                  // projected columns (L2 with the fused tail), [2] column sums

// out[i] += sum over the CTAs' partial records, in a fixed order (deterministic).  One block = 32 consecutive
// outputs x 32 slices of the CTA list (independent loads, combined through shared memory in slice order).
__global__ void __launch_bounds__(1024)
reduce_partials_kernel(const float* __restrict__ partial, int n_cta, int K, float* __restrict__ dW,
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
        for (int c0 = ty; c0 < n_cta; c0 += 320) {                  // ten records per round, all loads before the first add
            float va[10];
#pragma unroll
            for (int u = 0; u < 10; ++u) va[u] = c0 + 32 * u < n_cta ? __ldg(src + (size_t)(c0 + 32 * u) * H_PARTIAL_FLOATS) : 0.f;
#pragma unroll
            for (int u = 0; u < 10; ++u) a += va[u];
        }
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
    int defer;                // world > 1: this kernel leaves the LOCAL sums in d_flat; the exchange runs later (exchange_kernel)
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
bwd_tail_kernel(const float* __restrict__ partial, int n_cta, int K, TailP t) {
    __shared__ float red[32][33], red2[32][33], red3[32][33], red4[32][33];
    __shared__ float s_eff[64], s_tab[64];
    extern __shared__ float s_attr[];        // [K][A] (projection blocks)
    __shared__ float s_out[64];              // this block's finished outputs
    __shared__ int s_idx[64];                // their positions in the flat gradient
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * 32 + tx;
    const int Dl = 64 - t.Da;
    const int n_l = K * Dl, n_w = t.Da * t.A, n_flat = n_l + n_w + t.Da;
    const bool exchange = t.world > 1 && !t.defer;               // deferred: local sums only, exchange_kernel does the rest
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
            // ten records per round, every load issued before the first add (the sum keeps its fixed order: missing records
            // add 0): one round trip to L2 instead of one per unrolled group -- <= 320 records, i.e. one round on a B200
            for (int c0 = ty; c0 < n_cta; c0 += 320) {
                float va[10], vc[10];
#pragma unroll
                for (int u = 0; u < 10; ++u) {
                    const int cta = c0 + 32 * u;
                    const float* rec = partial + (size_t)cta * H_PARTIAL_FLOATS;
                    va[u] = cta < n_cta ? __ldg(rec + k * 64 + d) : 0.f;
                    vc[u] = cta < n_cta ? __ldg(rec + 2 * H_KD + k) : 0.f;
                }
#pragma unroll
                for (int u = 0; u < 10; ++u) { a += va[u]; c += vc[u]; }
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
                for (int c0 = ty; c0 < n_cta; c0 += 320) {           // (as above: all loads of a round first)
                    float va[10], vc[10];
#pragma unroll
                    for (int u = 0; u < 10; ++u) {
                        const int cta = c0 + 32 * u;
                        const float* rec = partial + (size_t)cta * H_PARTIAL_FLOATS;
                        va[u] = cta < n_cta ? __ldg(rec + H_KD + (Dl + j) * 64 + k) : 0.f;   // column-major copy: lanes read consecutive codes
                        vc[u] = cta < n_cta ? __ldg(rec + 2 * H_KD + k) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 10; ++u) { a[h] += va[u]; c[h] += vc[u]; }
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

// Deferred exchange (vqb_bwd_tail.reserved bit 0 / vqb_exchange_finish): the whole cross-GPU part as a kernel of its own.
// The tail kernel has left this rank's sums in d_flat; every output is pushed to all ranks' buffers (own included), then
// polls this rank's buffer for all ranks' words of the same epoch and adds them in rank order (identical bits everywhere).
// It runs wherever the caller puts it -- behind the rest of the model's backward in a trainer, on a side stream beside the
// next step's forward in the bench -- so neither the remote stores (a kernel does not retire before they are acknowledged
// over NVLink) nor the skew between ranks sit on the quantizer's critical path (profiles/r2_n2_exchange_ab.txt).  Pushes precede polls and no block waits for a block of its own GPU: no deadlock.
__global__ void __launch_bounds__(256)
exchange_kernel(float* __restrict__ d_flat, void* const* __restrict__ peer_bufs, int rank, unsigned int* counter,
                int n_flat, int world, unsigned long long timeout_ns) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned int epoch = *reinterpret_cast<volatile unsigned int*>(counter + 1) + 1u;
    const int n_pad = (n_flat + 3) & ~3;
    const size_t slot_off = (size_t)(epoch & 1u) * world * n_pad;
    if (i < n_flat) {
        const unsigned long long word = ((unsigned long long)epoch << 32) | __float_as_uint(d_flat[i]);
        for (int r = 0; r < world; ++r)
            st_relaxed_sys_b64(reinterpret_cast<unsigned long long*>(peer_bufs[r]) + slot_off + (size_t)rank * n_pad + i, word);
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peer_bufs[rank]) + slot_off + i;
        float sum = 0.f;
        const unsigned long long t0 = globaltimer_ns();
        for (int r = 0; r < world; ++r) {                           // rank order: identical bits on every GPU
            unsigned long long w = ld_relaxed_sys_b64(mine + (size_t)r * n_pad);
            bool gave_up = false;
            while ((unsigned int)(w >> 32) != epoch) {
                if (globaltimer_ns() - t0 > timeout_ns) { atomicMax(counter + 2, (unsigned int)(r + 1)); gave_up = true; break; }
                w = ld_relaxed_sys_b64(mine + (size_t)r * n_pad);
            }
            if (!gave_up) sum += __uint_as_float((unsigned int)w);
        }
        d_flat[i] = sum;
    }
    // the last block hands the ticket back and publishes the epoch (every block has read it by then)
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(counter, 1u) == gridDim.x - 1) { counter[0] = 0u; counter[1] = epoch; }
    }
}

// -----------------------------------------------------------------------------------------------------------
// host side
// -----------------------------------------------------------------------------------------------------------
int launch_exchange_finish(const vqb_bwd_tail* tl, int64_t n_flat, cudaStream_t s) {
    if (tl->world <= 1) return VQB_OK;
    const unsigned long long timeout_ns = (unsigned long long)(tl->timeout_ms ? tl->timeout_ms : 120000u) * 1000000ull;
    exchange_kernel<<<(unsigned)ceil_div(n_flat, 256), 256, 0, s>>>(tl->d_flat, tl->peer_bufs, tl->rank, tl->counter, (int)n_flat,
                                                                 tl->world, timeout_ns);
    VQB_CHECK_LAUNCH("exchange_kernel");
    return VQB_OK;
}

size_t exchange_bytes(int64_t n_flat, int world) { return exch_words(n_flat, world) * 8; }

// Behind the main backward kernel (vqb_bwd_pcode_kernel): the per-CTA partial records -> gradients.
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
        t.defer = (tl->world > 1 && (tl->reserved & 1u)) ? 1 : 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(t.n_learn_blocks + t.Da)); cfg.blockDim = dim3(32, 32); cfg.stream = s;
        cfg.dynamicSmemBytes = (size_t)K * t.A * 4;                  // <= 64 * 63 * 4 B
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = getenv("VQB_NO_PDL") ? 0 : 1;
        VQB_CUDA(cudaLaunchKernelEx(&cfg, bwd_tail_kernel, partial, grid, (int)K, t));
        VQB_CHECK_LAUNCH("bwd_tail_kernel");
        return VQB_OK;
    }
    float* dG = l2 ? nullptr : a->d_gather;
    const int n_out = (int)((dG ? 2 : 1) * K * 64 + K);
    reduce_partials_kernel<<<(unsigned)ceil_div(n_out, 32), dim3(32, 32), 0, s>>>(partial, grid, (int)K, a->d_score_w, dG,
                                                                                  a->colsum);
    VQB_CHECK_LAUNCH("reduce_partials_kernel");
    return VQB_OK;
}

}  // namespace vqb
