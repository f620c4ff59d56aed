// Run-length collapse after the quantizer (reference: VQVAE.mean_forward, src/vqvae.py:218-257), the step that
// directly follows the bottleneck on the unpaired-speech branch (src/vqvae.py:128).  The reference moves every index
// row to the host (.cpu().tolist()) and walks it in Python; here the whole batch is planned and reduced on the GPU
// and the host reads back B lengths once.
//
// Semantics restated (oracle/vq_oracle.py: mean_forward):
//   * a new segment starts at frame t when idx[t] != idx[t-1], or when the current segment already holds
//     max_frames_per_phn + 1 frames (:231: (t - last_pos) > max_frames_per_phn) -- i.e. inside a run of equal indices
//     that began at s, segments start at s, s + (max+1), s + 2 (max+1), ...
//   * segments of the blank code 0 are dropped (:233, :239)
//   * out[b, j, :] = mean over the frames of the j-th kept segment (:234, :242; a one-frame segment is that frame, :245)
//   * rows j >= lens[b] of the padded output are zero (pad_sequence, :254)
#include <limits.h>
#include <math.h>
#include "vqb_common.cuh"

namespace vqb {

constexpr int SEG_THREADS = 256;

// block-wide inclusive scans over SEG_THREADS threads (warp shuffles + one shared array of 8 warp totals)
__device__ __forceinline__ int block_scan_add(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += n; }
    __syncthreads();
    if (lane == 31) s_warp[w] = v;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int i = 0; i < SEG_THREADS / 32; ++i) { const int x = s_warp[i]; if (i < w) base += x; tot += x; }
    total = tot;
    return v + base;
}
__device__ __forceinline__ int block_scan_max(int v, int* s_warp) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v = max(v, n); }
    __syncthreads();
    if (lane == 31) s_warp[w] = v;
    __syncthreads();
    int base = INT_MIN;
#pragma unroll
    for (int i = 0; i < SEG_THREADS / 32; ++i) if (i < w) base = max(base, s_warp[i]);
    return max(v, base);
}

// One CTA per utterance; frames are visited in chunks of SEG_THREADS with the run start and the slot count carried over.
//   slot_of_row[b,t]  output slot of frame t, or -1 (blank)
//   seg_start[b,j]    first frame of kept segment j          (j < lens[b])
//   seg_count[b,j]    number of frames of kept segment j     (zero-filled by the caller; accumulated here)
__global__ void __launch_bounds__(SEG_THREADS)
segment_plan_kernel(const long long* __restrict__ idx, int T, int max_frames, int* __restrict__ slot_of_row,
                    int* __restrict__ seg_start, int* __restrict__ seg_count, long long* __restrict__ lens) {
    __shared__ int s_warp[SEG_THREADS / 32];
    const int b = blockIdx.x;
    const long long* row = idx + (size_t)b * T;
    int* slot_row = slot_of_row + (size_t)b * T;
    int* st_row = seg_start + (size_t)b * T;
    int* ct_row = seg_count + (size_t)b * T;
    const int period = max_frames + 1;
    int carry_run = 0;        // start of the run that is open at the chunk boundary
    int carry_slots = 0;      // kept segments before this chunk
    for (int t0 = 0; t0 < T; t0 += SEG_THREADS) {
        const int t = t0 + threadIdx.x;
        const bool in = t < T;
        const long long k = in ? row[t] : -1;
        const long long kp = (in && t > 0) ? row[t - 1] : -2;
        const bool change = in && (t == 0 || k != kp);
        int rs = block_scan_max(change ? t : INT_MIN, s_warp);
        if (rs == INT_MIN) rs = carry_run;
        const bool start = in && ((t - rs) % period == 0);
        const bool kept = k != 0;
        int total;
        const int incl = block_scan_add((start && kept) ? 1 : 0, s_warp, total);
        if (in) {
            const int slot = kept ? carry_slots + incl - 1 : -1;
            slot_row[t] = slot;
            if (kept) {
                if (start) st_row[slot] = t;
                atomicAdd(ct_row + slot, 1);
            }
        }
        // the run open at the end of this chunk starts at the last thread's rs
        __syncthreads();
        if (threadIdx.x == SEG_THREADS - 1 || t == T - 1) s_warp[0] = rs;
        __syncthreads();
        carry_run = s_warp[0];
        carry_slots += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) lens[b] = carry_slots;
}

// out[b, j, :] = mean of latent[b, start .. start+count-1, :] (j < lens[b]) or 0 -- one warp per output row
__global__ void __launch_bounds__(256)
segment_mean_kernel(const float* __restrict__ latent, const int* __restrict__ seg_start, const int* __restrict__ seg_count,
                    const long long* __restrict__ lens, int B, int T, int D, int Lmax, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long o = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (o >= (long long)B * Lmax) return;
    const int b = (int)(o / Lmax), j = (int)(o % Lmax);
    float* dst = out + (size_t)o * D;
    const int D4 = D >> 2;
    if (j >= lens[b]) {
        for (int c = lane; c < D4; c += 32) stg4_stream(dst + 4 * c, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    const int s = seg_start[(size_t)b * T + j], n = seg_count[(size_t)b * T + j];
    const float* src = latent + ((size_t)b * T + s) * D;
    const float fn = (float)n;
    for (int c = lane; c < D4; c += 32) {
        float4 a = ldg4_stream(src + 4 * c);
        for (int r = 1; r < n; ++r) {
            const float4 v = ldg4_stream(src + (size_t)r * D + 4 * c);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        if (n > 1) { a.x /= fn; a.y /= fn; a.z /= fn; a.w /= fn; }       // a one-frame segment is the frame itself (:245)
        stg4_stream(dst + 4 * c, a);
    }
}

// dlatent[b, t, :] = g_out[b, slot, :] / count(slot), or 0 for blank frames -- one warp per frame
__global__ void __launch_bounds__(256)
segment_mean_backward_kernel(const float* __restrict__ g_out, const int* __restrict__ slot_of_row,
                             const int* __restrict__ seg_count, int B, int T, int D, int Lmax, float* __restrict__ dlatent) {
    const int lane = threadIdx.x & 31;
    const long long o = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (o >= (long long)B * T) return;
    const int b = (int)(o / T);
    const int slot = slot_of_row[o];
    float* dst = dlatent + (size_t)o * D;
    const int D4 = D >> 2;
    if (slot < 0) {
        for (int c = lane; c < D4; c += 32) stg4_stream(dst + 4 * c, make_float4(0.f, 0.f, 0.f, 0.f));
        return;
    }
    const float fn = (float)seg_count[(size_t)b * T + slot];
    const float* src = g_out + ((size_t)b * Lmax + slot) * D;
    for (int c = lane; c < D4; c += 32) {
        float4 v = ldg4(src + 4 * c);
        v.x /= fn; v.y /= fn; v.z /= fn; v.w /= fn;
        stg4_stream(dst + 4 * c, v);
    }
}

// idx[n] = first index of the row maximum of p[n, :K]  (p_code.argmax(-1), src/vqvae.py:223) -- one warp per row
__global__ void __launch_bounds__(256)
row_argmax_kernel(const float* __restrict__ p, long long n, int K, long long* __restrict__ idx) {
    const int lane = threadIdx.x & 31;
    const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= n) return;
    const float* row = p + (size_t)r * K;
    float bv = -INFINITY;
    int bi = INT_MAX;
    for (int k = lane; k < K; k += 32) {
        const float v = __ldg(row + k);
        if (v > bv) { bv = v; bi = k; }                                  // strict >: the first maximum of this lane
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) idx[r] = bi == INT_MAX ? 0 : bi;
}

}  // namespace vqb

using namespace vqb;

extern "C" int vqb_row_argmax(const float* p, int64_t n_rows, int64_t n_codes, int64_t* idx, void* stream) {
    if (n_rows == 0) return VQB_OK;
    if (!p || !idx || n_codes <= 0) return invalid("vqb_row_argmax: NULL pointer or K <= 0");
    row_argmax_kernel<<<(unsigned)ceil_div(n_rows, 8), 256, 0, (cudaStream_t)stream>>>(p, n_rows, (int)n_codes, (long long*)idx);
    VQB_CHECK_LAUNCH("row_argmax_kernel");
    return VQB_OK;
}

extern "C" int vqb_segment_plan(const int64_t* idx, int64_t n_utts, int64_t n_frames, int64_t max_frames_per_phn,
                                int32_t* slot_of_row, int32_t* seg_start, int32_t* seg_count, int64_t* lens, void* stream) {
    if (n_utts == 0 || n_frames == 0) return VQB_OK;
    if (!idx || !slot_of_row || !seg_start || !seg_count || !lens) return invalid("vqb_segment_plan: NULL pointer");
    if (max_frames_per_phn < 0 || n_frames >= (1ll << 30) || n_utts >= (1ll << 30) || max_frames_per_phn >= (1ll << 30))
        return invalid("vqb_segment_plan: bad shape B=%lld T=%lld max_frames_per_phn=%lld", (long long)n_utts,
                       (long long)n_frames, (long long)max_frames_per_phn);
    cudaStream_t s = (cudaStream_t)stream;
    VQB_CUDA(cudaMemsetAsync(seg_count, 0, (size_t)n_utts * n_frames * sizeof(int32_t), s));
    segment_plan_kernel<<<(unsigned)n_utts, SEG_THREADS, 0, s>>>((const long long*)idx, (int)n_frames, (int)max_frames_per_phn,
                                                                 slot_of_row, seg_start, seg_count, (long long*)lens);
    VQB_CHECK_LAUNCH("segment_plan_kernel");
    return VQB_OK;
}

extern "C" int vqb_segment_mean(const float* latent, const int32_t* seg_start, const int32_t* seg_count, const int64_t* lens,
                                int64_t n_utts, int64_t n_frames, int64_t dim, int64_t max_len, float* out, void* stream) {
    if (n_utts == 0 || max_len == 0) return VQB_OK;
    if (!latent || !seg_start || !seg_count || !lens || !out) return invalid("vqb_segment_mean: NULL pointer");
    if (dim % 4 != 0 || !aligned16(latent) || !aligned16(out))
        return invalid("vqb_segment_mean: D must be a multiple of 4 and pointers 16-byte aligned");
    const int64_t warps = n_utts * max_len;
    segment_mean_kernel<<<(unsigned)ceil_div(warps, 8), 256, 0, (cudaStream_t)stream>>>(
        latent, seg_start, seg_count, (const long long*)lens, (int)n_utts, (int)n_frames, (int)dim, (int)max_len, out);
    VQB_CHECK_LAUNCH("segment_mean_kernel");
    return VQB_OK;
}

extern "C" int vqb_segment_mean_backward(const float* g_out, const int32_t* slot_of_row, const int32_t* seg_count,
                                         int64_t n_utts, int64_t n_frames, int64_t dim, int64_t max_len, float* dlatent,
                                         void* stream) {
    if (n_utts == 0 || n_frames == 0) return VQB_OK;
    if (!g_out || !slot_of_row || !seg_count || !dlatent) return invalid("vqb_segment_mean_backward: NULL pointer");
    if (dim % 4 != 0 || !aligned16(g_out) || !aligned16(dlatent))
        return invalid("vqb_segment_mean_backward: D must be a multiple of 4 and pointers 16-byte aligned");
    const int64_t warps = n_utts * n_frames;
    segment_mean_backward_kernel<<<(unsigned)ceil_div(warps, 8), 256, 0, (cudaStream_t)stream>>>(
        g_out, slot_of_row, seg_count, (int)n_utts, (int)n_frames, (int)dim, (int)max_len, dlatent);
    VQB_CHECK_LAUNCH("segment_mean_backward_kernel");
    return VQB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// CTC input preparation (reference: bin/train_vqvae.py:430-432 and :236): ctc_input = (p_code + EPS).transpose(0,1).log()
// -- the only consumer of p_code in training.  One pass: out[s,b,:] = log(p[b,s,:] + eps), contiguous [S,B,K] as
// nn.CTCLoss wants it; consecutive threads walk consecutive OUTPUT elements, so reads and writes are both whole
// K-float runs.  Backward: g_p[b,s,k] (+)= g_out[s,b,k] / (p[b,s,k] + eps).
// ---------------------------------------------------------------------------------------------------------------
namespace vqb {

__global__ void __launch_bounds__(256)
ctc_logp_kernel(const float* __restrict__ p, int B, int S, int K, float eps, float* __restrict__ out) {
    const long long n = (long long)B * S * K;
    for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < n; o += (long long)gridDim.x * blockDim.x) {
        const long long row = o / K;                       // = s * B + b
        const int k = (int)(o - row * K);
        const int s = (int)(row / B), b = (int)(row - (long long)s * B);
        out[o] = logf(__ldg(p + ((long long)b * S + s) * K + k) + eps);
    }
}

__global__ void __launch_bounds__(256)
ctc_logp_backward_kernel(const float* __restrict__ g_out, const float* __restrict__ p, int B, int S, int K, float eps,
                         float* __restrict__ g_p, int accumulate) {
    const long long n = (long long)B * S * K;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long row = i / K;                       // = b * S + s
        const int k = (int)(i - row * K);
        const int b = (int)(row / S), s = (int)(row - (long long)b * S);
        const float g = __ldg(g_out + ((long long)s * B + b) * K + k) / (__ldg(p + i) + eps);
        g_p[i] = accumulate ? g_p[i] + g : g;
    }
}

}  // namespace vqb

extern "C" int vqb_ctc_logp(const float* p_code, int64_t n_utts, int64_t n_frames, int64_t n_codes, float eps, float* out,
                            void* stream) {
    const int64_t n = n_utts * n_frames * n_codes;
    if (n == 0) return VQB_OK;
    if (!p_code || !out) return invalid("vqb_ctc_logp: NULL pointer");
    if (n_utts >= (1ll << 31) || n_frames >= (1ll << 31) || n_codes >= (1ll << 31)) return invalid("vqb_ctc_logp: shape too large");
    const int64_t blocks = ceil_div(n, 256), cap = (int64_t)sm_count() * 16;
    ctc_logp_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(p_code, (int)n_utts, (int)n_frames,
                                                                                              (int)n_codes, eps, out);
    VQB_CHECK_LAUNCH("ctc_logp_kernel");
    return VQB_OK;
}

extern "C" int vqb_ctc_logp_backward(const float* g_out, const float* p_code, int64_t n_utts, int64_t n_frames, int64_t n_codes,
                                     float eps, float* g_p, int accumulate, void* stream) {
    const int64_t n = n_utts * n_frames * n_codes;
    if (n == 0) return VQB_OK;
    if (!g_out || !p_code || !g_p) return invalid("vqb_ctc_logp_backward: NULL pointer");
    if (n_utts >= (1ll << 31) || n_frames >= (1ll << 31) || n_codes >= (1ll << 31)) return invalid("vqb_ctc_logp_backward: shape too large");
    const int64_t blocks = ceil_div(n, 256), cap = (int64_t)sm_count() * 16;
    ctc_logp_backward_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, (cudaStream_t)stream>>>(
        g_out, p_code, (int)n_utts, (int)n_frames, (int)n_codes, eps, g_p, accumulate);
    VQB_CHECK_LAUNCH("ctc_logp_backward_kernel");
    return VQB_OK;
}
