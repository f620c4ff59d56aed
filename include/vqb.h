/*
 * vqb.h -- C ABI of libvqb200.so: the semi-tts vector-quantisation bottleneck on B200 (sm_100a).
 *
 * The reference (ttaoREtw/semi-tts) has no FFI: its quantizer is the Python nn.Module surface of
 * src/embed.py (L2Embedding :57-147, SeperateEmbedding :150-205, neg_batch_l2 :208-213), bound by
 * name at src/vqvae.py:8 and src/tts.py:5.  This header is what a torch.autograd.Function binds
 * (through ctypes) to replace the ATen op sequence of those methods; INTEGRATION.md shows the
 * reference-side stub.  Each entry point cites the reference lines it replaces.
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer owned by the caller (the PyTorch caching allocator);
 *    the library keeps no persistent device allocations;
 *  - all tensors are contiguous row-major fp32, indices are int64, histogram counts are int64;
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises the host;
 *  - return value 0 = success, otherwise a VQB_ERR_* code and vqb_last_error() (thread-local) holds
 *    the message; CUDA out-of-memory keeps the substring "out of memory" (bin/train_vqvae.py:321);
 *  - there is NO CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Notation: N rows (= B*S encoder frames), D = latent_dim, K = vocab_size (codebook size),
 *           A = phoneme attributes (31), D_a = proj_attr (16), D_l = D - D_a.
 */
#ifndef VQB_H_
#define VQB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQB_ABI_VERSION 3

#if defined(__GNUC__)
#define VQB_API __attribute__((visibility("default")))
#else
#define VQB_API
#endif

/* error codes */
#define VQB_OK              0
#define VQB_ERR_INVALID     1   /* bad argument / unsupported shape */
#define VQB_ERR_CUDA        2   /* a CUDA runtime call failed */
#define VQB_ERR_WORKSPACE   3   /* workspace too small */
#define VQB_ERR_NO_DEVICE   4   /* no sm_100 device visible */

/* flags (vqb_fwd_args.flags / vqb_bwd_args.flags) */
#define VQB_SCORE_L2        0x0001u  /* score = -relu(temp) * ((|x|^2 + |e|^2) - 2 x.e)   src/embed.py:115-124,208-213 */
#define VQB_SCORE_LINEAR    0x0002u  /* score = x.W^T + b                                 src/embed.py:190 */
#define VQB_STOP_GRAD       0x0004u  /* gather route is F.embedding (:134 / :194-197); clear = ST-onehot (:137-138 / :199-203) */
#define VQB_SKIP            0x0008u  /* skip connection drawn this step: new_latent = x   src/embed.py:140-142 */
#define VQB_TEMP_GRAD       0x0010u  /* temp is an nn.Parameter: also produce d temp      src/embed.py:33-34 */
#define VQB_TENSOR_CORES    0x0020u  /* run the forward on the tcgen05 kernel where the shape allows it (p_code mode:
                                        K <= 64, D in {32,64}; fused mode: L2 score, D in {32,64,128,256}); other
                                        shapes, or a cleared flag, take the exact-fp32 CUDA-core kernel */
#define VQB_SEARCH_TENSOR   VQB_TENSOR_CORES
#define VQB_AFTER_ASSEMBLE  0x0040u  /* vqb_forward only: the caller vouches that the operation enqueued on `stream`
                                        immediately before this call is vqb_assemble_table (the usual sequence).  The
                                        forward kernel is then launched with programmatic stream serialization: its
                                        prologue and first x tile overlap the assembly kernel, and it waits
                                        (griddepcontrol.wait) before touching the table.  Never set it after a memcpy
                                        or a foreign kernel that produces x. */

VQB_API int vqb_abi_version(void);
VQB_API const char* vqb_last_error(void);
/* number of sm_100 devices usable by this library (0 if none); never fails */
VQB_API int vqb_device_count(void);

/* ---------------------------------------------------------------------------------------------
 * Codebook table assembly  (replaces src/embed.py:109-112, the same expression in :87-94)
 *   table[K,D] = cat([learnable[K,D_l], phn_attr[K,A] @ proj_w[D_a,A]^T + proj_b[D_a]], -1)
 *   enorm[K]   = sum_d table[k,d]^2                      (the |e|^2 term of src/embed.py:211)
 * phn_attr / proj_w / proj_b may be NULL (n_attr = dim_attr = 0): table = learnable.
 * table_bf16 (optional, may be NULL): bf16 copy [K, D] for the tensor-core search.
 * operand_cache (optional, may be NULL): vqb_operand_cache_bytes() bytes that receive, in the same launch, the
 *   tf32 hi/lo operand copies of the table that the tensor-core forward (L2 score) and backward consume; pass
 *   the buffer on in vqb_fwd_args.operand_cache / vqb_bwd_args.operand_cache to skip their own operand pass.
 * ------------------------------------------------------------------------------------------- */
VQB_API size_t vqb_operand_cache_bytes(int64_t n_codes, int64_t dim);
VQB_API int vqb_assemble_table(const float* learnable, const float* phn_attr, const float* proj_w,
                       const float* proj_b, int64_t n_codes, int64_t dim, int64_t n_attr,
                       int64_t dim_attr, float* table, float* enorm, void* table_bf16,
                       void* operand_cache, void* stream);

/* Backward of the assembly (autograd of src/embed.py:109-112):
 *   dtable[k,:] += 2 * table[k,:] * colsum[k]     if colsum != NULL (the |e|^2 term of the L2 route)
 *   d_learnable[K,D_l] = dtable[:, :D_l]
 *   d_proj_w[D_a,A]    = dtable[:, D_l:]^T @ phn_attr ;  d_proj_b[D_a] = colsum_k dtable[:, D_l:]
 * d_learnable may alias nothing; all outputs are overwritten (not accumulated). */
VQB_API int vqb_table_backward(const float* dtable, const float* table, const float* colsum,
                       const float* phn_attr, int64_t n_codes, int64_t dim, int64_t n_attr,
                       int64_t dim_attr, float* d_learnable, float* d_proj_w, float* d_proj_b,
                       void* stream);

/* ---------------------------------------------------------------------------------------------
 * Forward  (replaces L2Embedding.forward src/embed.py:105-147 / SeperateEmbedding.forward :187-205)
 * ------------------------------------------------------------------------------------------- */
typedef struct vqb_fwd_args {
    uint32_t struct_size;        /* sizeof(vqb_fwd_args) */
    uint32_t flags;
    int64_t n_rows, dim, n_codes;
    const float* x;              /* [N,D]   enc_embs flattened (src/embed.py:209)                      */
    const float* score_w;        /* [K,D]   L2: assembled table E; LINEAR: asr_final_layer.weight      */
    const float* score_b;        /* [K]     L2: enorm |e|^2;       LINEAR: asr_final_layer.bias        */
    const void*  score_w_bf16;   /* [K,D]   optional bf16 copy of score_w (reserved for a bf16 search)  */
    const float* gather_table;   /* [K,D]   codewords gathered into new_latent (L2: == score_w)        */
    const float* temp;           /* [1]     device scalar; tau = relu(temp) (src/embed.py:115)         */
    float*   p_code;             /* [N,K]   softmax over codes (src/embed.py:127); NULL = fused mode   */
    int64_t* idx;                /* [N]     argmax(p_code) (src/embed.py:130)                          */
    float*   new_latent;         /* [N,D]   L2: (x + E[idx]) - x (:145) or x if VQB_SKIP; LINEAR: T[idx] */
    int64_t* hist;               /* [K]     += per-code usage counts of this call, or NULL
                                            (bin/train_vqvae.py:256-261; src/util.py:139)              */
    double*  sq_err_sum;         /* [1]     += sum (x - E[idx])^2 (numerator of the loss extensions), or NULL */
    uint32_t* search_stats;      /* [2]     tensor-core fused mode only, or NULL: += rows re-ranked in exact fp32,
                                            += rows that needed the full exact scan                      */
    const void* operand_cache;   /* from vqb_assemble_table (L2 score only), or NULL                    */
    void*    workspace;          /* vqb_forward_workspace() bytes, or NULL if that is 0                 */
    size_t   workspace_bytes;
    const int64_t* row_lengths;  /* [n_rows / frames_per_utt] or NULL.  Length-aware rows (the padded
                                    batches of src/vqvae.py:106-126,259-271): frame s of utterance b is a PAD row when
                                    s >= row_lengths[b].  Pad rows are not searched: their p_code / new_latent rows are
                                    zero, their index is 0, they are not counted in hist; tiles that hold only pad rows
                                    are neither loaded nor computed.  Valid rows are bit-identical to the dense call.
                                    Parity-mode tensor-core route only (p_code given, K <= 64, D in {32,64}). */
    int64_t  frames_per_utt;     /* S; required (> 0, dividing n_rows) when row_lengths or ctc_logp is given */
    float*   ctc_logp;           /* [S, B, K] or NULL: log(p_code + ctc_eps) transposed to the layout nn.CTCLoss consumes,
                                    written by the forward epilogue itself (replaces `(p + EPS).transpose(0,1).log()`,
                                    bin/train_vqvae.py:430-432 and :236).  Parity-mode tensor-core route. */
    float    ctc_eps;            /* EPS of bin/train_vqvae.py:18 (1e-10) */
} vqb_fwd_args;

VQB_API int vqb_forward_workspace(const vqb_fwd_args* args, size_t* bytes);
VQB_API int vqb_forward(const vqb_fwd_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Backward (autograd of the above, entered from src/solver.py:144; algebra in DESIGN.md)
 *   G   = g_p (+ g_q @ T^T when !STOP_GRAD);  Gs = P * (G - rowsum(G*P))
 *   L2:     Gd = -tau*Gs;  dx = g_q + 2 x rowsum(Gd) - 2 Gd@E
 *           d_score_w = -2 Gd*^T @ x + scatter_add(idx, g_q) ; colsum = colsum(Gd*)   (Gd* = rows < n_real_rows)
 *           d_temp = sum Gs * (-dist) * [temp > 0]
 *   LINEAR: dx = Gs@W;  d_score_w = Gs^T @ x;  colsum = colsum(Gs);  d_gather = scatter_add(idx, g_q)
 * Routes (vqb_backward_kernel_name): K <= 64 with D = 64 and STOP_GRAD runs on the tensor cores; K <= 64 with
 * D % 8 == 0, D <= 128 on the register-tiled exact-fp32 kernel (all flag combinations); every other shape (any K,
 * D % 4 == 0, D <= 512) on the any-K route, which keeps the N x K coefficient matrix in the workspace
 * (vqb_backward_workspace grows by 4 N K bytes) and serves the same flag combinations.
 * All outputs are ACCUMULATED into (+=): the caller zeroes d_score_w, colsum, d_gather, d_temp first.
 * dx is overwritten.  If g_p == NULL, STOP_GRAD and L2, only the scatter route runs and dx may be
 * NULL (the caller aliases dx = g_q: the straight-through identity costs zero bytes).
 * ------------------------------------------------------------------------------------------- */
/* Optional fused tail of vqb_backward (L2 score, tensor-core route): the fixed-order sum of the per-CTA partial
 * gradients, the backward of the table assembly (vqb_table_backward) and -- in data-parallel runs -- the sum of the
 * result over all GPUs run as ONE block-parallel kernel behind the main backward kernel, instead of three launches,
 * a fill and an NCCL call.  Every block finishes a few outputs of the flat gradient on its own, pushes them into every
 * peer's exchange buffer over NVLink as 8-byte (value, epoch) words -- the data carries its own flag, so there is no
 * fence and no separate signal -- polls its own buffer for the peers' words of the same outputs and adds them in rank
 * order (all ranks obtain bit-identical sums).
 *   With a tail, d_score_w and colsum are not used and may be NULL.
 * Exchange buffer of each rank (peer-mapped, e.g. torch symmetric memory), vqb_exchange_bytes(n_flat, world) bytes,
 * zero-filled once before the first call:  [2 slots][world senders][n_flat padded to a multiple of 4] 8-byte words.
 * Words carry a call counter (`epoch`, never 0) and the slots alternate with its parity, so the buffers need no reset
 * between calls or CUDA-graph replays; all ranks must make the same sequence of calls.  A rank whose shard is empty
 * (n_rows == 0) still calls vqb_backward with the tail: only the tail kernel runs, over zero partial records. */
typedef struct vqb_bwd_tail {
    const float* phn_attr;       /* [K,A], A <= 63, or NULL (then n_attr = dim_attr = 0 and d_flat = d_learnable only) */
    int64_t n_attr, dim_attr;
    float* d_flat;               /* [K*D_l + D_a*A + D_a] = d_learnable | d_proj_w | d_proj_b, overwritten */
    uint32_t* counter;           /* [4] device words, zero before the first call: [0] block ticket (left zero),
                                    [1] epoch of the exchange (incremented by every call with world > 1),
                                    [2] error flag: 0, or 1 + the rank an exchange gave up waiting for (sticky; the host
                                    reads it when it synchronises anyway), [3] reserved */
    int32_t world, rank;         /* world <= 1: no exchange */
    void* const* peer_bufs;      /* DEVICE array [world] of each rank's exchange-buffer address as mapped here */
    uint32_t timeout_ms;         /* how long a block polls for a peer's words before it raises counter[2] and returns
                                    with an incomplete sum (no trap: the context survives); 0 = 120 000 ms */
    uint32_t reserved;           /* bit 0 (VQB_TAIL_DEFER), world > 1: this call leaves THIS rank's sums in d_flat and touches no
                                    peer memory; vqb_exchange_finish() runs the exchange later, on a stream of the caller's
                                    choice */
} vqb_bwd_tail;
#define VQB_TAIL_DEFER 1u

#define VQB_MAX_WORLD 16
VQB_API size_t vqb_exchange_bytes(int64_t n_flat, int32_t world);
/* The deferred exchange (tail.reserved & VQB_TAIL_DEFER): push d_flat[0 .. n_flat) to every rank's buffer, poll this rank's
 * buffer for every rank's words of the same exchange, add them in rank order and write the sums back to d_flat.  Enqueue it
 * on any stream ordered after the vqb_backward call that produced d_flat; the module's exchanges must be enqueued in the
 * same order on every rank, one after the other (the exchange slots alternate: a rank may run one exchange ahead of its
 * peers, not two).  On a side stream, the remote stores, the NVLink round trip and the skew of the ranks are hidden behind
 * whatever the caller runs meanwhile (the rest of the model's backward, the next forward). */
VQB_API int vqb_exchange_finish(const vqb_bwd_tail* tail, int64_t n_flat, void* stream);

typedef struct vqb_bwd_args {
    uint32_t struct_size;
    uint32_t flags;
    int64_t n_rows, dim, n_codes;
    int64_t n_real_rows;         /* first_n_real_mel * S; rows >= this do not reach the table through
                                    the distance route (src/embed.py:115-122); <= 0 means all rows do */
    const float* x;              /* [N,D] */
    const float* score_w;        /* [K,D] */
    const float* score_b;        /* [K]   (L2: enorm, needed only with VQB_TEMP_GRAD) */
    const float* gather_table;   /* [K,D] */
    const float* temp;           /* [1]   */
    const float* p_code;         /* [N,K] saved forward output; NULL if g_p == NULL and STOP_GRAD */
    const int64_t* idx;          /* [N]   */
    const float* g_p;            /* [N,K] upstream grad of p_code, or NULL */
    const float* g_q;            /* [N,D] upstream grad of new_latent, or NULL */
    float* dx;                   /* [N,D] */
    float* d_score_w;            /* [K,D] += */
    float* colsum;               /* [K]   += */
    float* d_gather;             /* [K,D] += (LINEAR only; L2 accumulates the scatter into d_score_w) */
    float* d_temp;               /* [1]   += (only with VQB_TEMP_GRAD) */
    const void* operand_cache;   /* from vqb_assemble_table (L2 score only), or NULL */
    void*  workspace;
    size_t workspace_bytes;
    const int64_t* row_lengths;  /* as in vqb_fwd_args: pad rows contribute nothing (their dx rows are zero, whatever g_p / g_q
                                    hold there); vqb_bwd_pcode_kernel route only */
    int64_t  frames_per_utt;
    const float* g_logp;         /* [S, B, K] or NULL: upstream gradient of the forward's ctc_logp output.  Folded into the
                                    softmax backward: G = g_p + g_logp^T / (p_code + ctc_eps), staged row by row by the kernel
                                    itself -- no [N, K] gradient tensor is materialised.  Give EITHER g_p or g_logp (g_p may be
                                    NULL then); vqb_bwd_pcode_kernel route, frames_per_utt required. */
    float    ctc_eps;
    const vqb_bwd_tail* tail;    /* optional fused tail (see above); NULL = plain accumulate-into semantics.  Only
                                    taken on the route vqb_backward_kernel_name() reports as "vqb_bwd_pcode_kernel"
                                    with VQB_SCORE_L2; otherwise vqb_backward fails with VQB_ERR_INVALID */
} vqb_bwd_args;

VQB_API int vqb_backward_workspace(const vqb_bwd_args* args, size_t* bytes);
VQB_API int vqb_backward(const vqb_bwd_args* args, void* stream);

/* Name of the dominant kernel vqb_forward / vqb_backward would launch for these arguments (a static string;
 * no device work).  Measurement aid: bench.py and the tests label timings / assert the tensor-core route with it. */
/* kernels of this library launched (or captured into a CUDA graph) by this process so far; host-side tally */
VQB_API uint64_t vqb_launch_count(void);
VQB_API const char* vqb_forward_kernel_name(const vqb_fwd_args* args);
VQB_API const char* vqb_backward_kernel_name(const vqb_bwd_args* args);

/* ---------------------------------------------------------------------------------------------
 * Gather-only paths
 * ------------------------------------------------------------------------------------------- */
/* out[m,:] = table[txt[m],:]  (replaces L2Embedding.inference src/embed.py:96-103 and
 * SeperateEmbedding.inference :180-185 on the assembled table).  Returns VQB_ERR_INVALID through a
 * device-side flag only in debug builds; indices are clamped to [0,K) like a bounds-checked gather. */
VQB_API int vqb_inference_gather(const int64_t* txt, int64_t n_tokens, const float* table, int64_t n_codes,
                         int64_t dim, float* out, void* stream);

/* dtable[txt[m],:] += g[m,:]; hist[txt[m]] += 1 (hist may be NULL)  (autograd of the gather: F.embedding backward,
 * src/embed.py:134; histogram semantics of bin/train_vqvae.py:256-261).
 * Small tables (8 copies fit in shared memory) use per-warp private accumulators; large tables with at least 64 Ki
 * tokens use a ticket + permutation + per-code gather-sum that needs vqb_scatter_workspace() bytes of scratch (without
 * it, or below that size, rows go to the table by 128-bit global reductions). */
VQB_API int vqb_scatter_workspace(int64_t n_tokens, int64_t n_codes, int64_t dim, size_t* bytes);
VQB_API int vqb_scatter_add(const int64_t* txt, int64_t n_tokens, const float* g, int64_t n_codes,
                    int64_t dim, float* dtable, int64_t* hist, void* workspace, size_t workspace_bytes, void* stream);

/* Loss extensions (NO reference arithmetic; van den Oord et al. 2017): backward of
 *   vq_loss = commit_loss = mean((x - E[idx])^2):
 *   dx[n,:] (+)= g_commit * 2 (x - c) / (N D);  dtable[idx[n],:] += g_vq * 2 (c - x) / (N D)
 * g_vq / g_commit are device scalars (may be NULL = 0). dx_accumulate != 0 adds into dx. */
VQB_API int vqb_loss_backward(const float* x, const float* table, const int64_t* idx, int64_t n_rows,
                      int64_t dim, int64_t n_codes, const float* g_vq, const float* g_commit,
                      float* dx, int dx_accumulate, float* dtable, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Run-length collapse after the quantizer  (replaces VQVAE.mean_forward, src/vqvae.py:218-257: the
 * `.cpu().tolist()` + per-utterance Python loop that follows the bottleneck on the unpaired branch, :128)
 *   A segment starts at frame t when idx[t] != idx[t-1] or the open segment already holds
 *   max_frames_per_phn + 1 frames (:231); segments of code 0 (blank) are dropped (:233,:239); the
 *   output row of a kept segment is the mean of its frames (:234,:242,:245); rows beyond lens[b] are 0.
 * ------------------------------------------------------------------------------------------- */
/* idx[n] = first index of the maximum of p[n,:K]  (p_code.argmax(-1), src/vqvae.py:223) */
VQB_API int vqb_row_argmax(const float* p, int64_t n_rows, int64_t n_codes, int64_t* idx, void* stream);
/* Plan: slot_of_row[B,T] (output slot of each frame, -1 = blank), seg_start[B,T] / seg_count[B,T] (first frame and
 * frame count of kept segment j < lens[b]; the rest of seg_start is unspecified, of seg_count zero), lens[B]. */
VQB_API int vqb_segment_plan(const int64_t* idx, int64_t n_utts, int64_t n_frames, int64_t max_frames_per_phn,
                     int32_t* slot_of_row, int32_t* seg_start, int32_t* seg_count, int64_t* lens, void* stream);
/* out[B,max_len,D]: segment means, zero rows for j >= lens[b] (max_len = max(lens), read back by the host) */
VQB_API int vqb_segment_mean(const float* latent, const int32_t* seg_start, const int32_t* seg_count, const int64_t* lens,
                     int64_t n_utts, int64_t n_frames, int64_t dim, int64_t max_len, float* out, void* stream);
/* dlatent[B,T,D] = g_out[b, slot, :] / count(slot), 0 for blank frames (autograd of the means) */
VQB_API int vqb_segment_mean_backward(const float* g_out, const int32_t* slot_of_row, const int32_t* seg_count,
                     int64_t n_utts, int64_t n_frames, int64_t dim, int64_t max_len, float* dlatent, void* stream);

/* ---------------------------------------------------------------------------------------------
 * CTC input preparation  (replaces `(p_code + EPS).transpose(0,1).log()`, bin/train_vqvae.py:430-432 and :236,
 * EPS = 1e-10: the only consumer of p_code in training)
 *   out[s,b,:] = log(p_code[b,s,:] + eps)            contiguous [S,B,K]
 *   g_p[b,s,k] (+)= g_out[s,b,k] / (p_code[b,s,k] + eps)      (accumulate != 0 adds into g_p)
 * ------------------------------------------------------------------------------------------- */
VQB_API int vqb_ctc_logp(const float* p_code, int64_t n_utts, int64_t n_frames, int64_t n_codes, float eps, float* out,
                 void* stream);
VQB_API int vqb_ctc_logp_backward(const float* g_out, const float* p_code, int64_t n_utts, int64_t n_frames,
                 int64_t n_codes, float eps, float* g_p, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQB_H_ */
