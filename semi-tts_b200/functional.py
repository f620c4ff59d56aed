"""torch.autograd.Function wrappers over the C ABI (include/vqb.h).

Each Function replaces the ATen op sequence of one reference method and its autograd backward:
  vq_l2           L2Embedding.forward         src/embed.py:105-147 (+ neg_batch_l2 :208-213)
  vq_linear       SeperateEmbedding.forward   src/embed.py:187-205
  codebook_lookup *.inference                 src/embed.py:96-103, :180-185
PyTorch only supplies device memory, the current stream and the autograd graph; all arithmetic is in
libvqb200.so.  CPU tensors are rejected: there is no fallback path.
"""
import ctypes

import torch

from . import _lib
from ._lib import ptr


def _require(t, name, dtype=torch.float32):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError("semi-tts_b200: `%s` must be a CUDA tensor -- this package has no CPU path" % name)
    if t.dtype != dtype:
        raise RuntimeError("semi-tts_b200: `%s` must be %s (got %s)" % (name, dtype, t.dtype))


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream(t):
    """the current CUDA stream of the tensor's device as a raw handle (the C ABI takes a void*)"""
    if _raw_stream is not None:
        return _raw_stream(t.device.index)           # no Stream object: a fraction of a microsecond
    return torch.cuda.current_stream(t.device).cuda_stream


def _c(t):
    return None if t is None else t.contiguous()


class _NullCtx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


_NULL = _NullCtx()


def _on(dev):
    """context that makes `dev` current for the C-ABI call; free when it already is (the usual case)"""
    return _NULL if torch.cuda.current_device() == dev.index else torch.cuda.device(dev)


# ------------------------------------------------------------------------------------------------
# raw (non-differentiable) table assembly: src/embed.py:109-112
# ------------------------------------------------------------------------------------------------
def assemble_table(learnable, phn_attr=None, proj_w=None, proj_b=None, want_bf16=False, want_cache=False):
    """Returns (table[K,D], enorm[K], table_bf16 or None[, operand_cache]).  No autograd."""
    lib = _lib.load()
    _require(learnable, "learnable_table")
    learnable = _c(learnable.detach())
    K, Dl = learnable.shape
    A = Da = 0
    if phn_attr is not None:
        _require(phn_attr, "phn_attr"); _require(proj_w, "proj_attr.weight"); _require(proj_b, "proj_attr.bias")
        phn_attr, proj_w, proj_b = _c(phn_attr.detach()), _c(proj_w.detach()), _c(proj_b.detach())
        A, Da = phn_attr.shape[1], proj_w.shape[0]
    D = Dl + Da
    table = torch.empty(K, D, device=learnable.device, dtype=torch.float32)
    enorm = torch.empty(K, device=learnable.device, dtype=torch.float32)
    tbf = torch.empty(K, D, device=learnable.device, dtype=torch.bfloat16) if want_bf16 else None
    cache = None
    if want_cache:
        nb = lib.vqb_operand_cache_bytes(K, D)           # 0: this shape has no tensor-core operand image
        cache = torch.empty(nb, device=learnable.device, dtype=torch.uint8) if nb else None
    with torch.cuda.device(learnable.device):
        _lib.check(lib.vqb_assemble_table(ptr(learnable), ptr(phn_attr), ptr(proj_w), ptr(proj_b), K, D, A, Da,
                                          ptr(table), ptr(enorm), ptr(tbf), ptr(cache), _stream(learnable)))
    if want_cache:
        return table, enorm, tbf, cache
    return table, enorm, tbf


def _table_backward(dtable, table, colsum, phn_attr, Da, tail=None):
    """Returns (d_learnable, d_proj_w, d_proj_b) from the accumulated dtable (+ the |e|^2 term).  With a fused exchange
    attached to the module (`tail.exchange`), the flat result is summed over the data-parallel group here (NCCL), so that
    every gradient autograd receives from this package is already a global sum (dist.reduce_route_output)."""
    lib = _lib.load()
    K, D = dtable.shape
    A = phn_attr.shape[1] if phn_attr is not None else 0
    if phn_attr is None:
        Da = 0
    # the three parameter gradients are views of ONE flat buffer, so the data-parallel layer can all-reduce them
    # in place with a single collective and no pack / unpack kernels (dist.allreduce_codebook_grads)
    n_l, n_w = K * (D - Da), Da * A
    flat = torch.empty(n_l + n_w + Da, device=dtable.device, dtype=torch.float32)
    d_learn = flat[:n_l].view(K, D - Da)
    d_w = flat[n_l:n_l + n_w].view(Da, A) if Da else None
    d_b = flat[n_l + n_w:] if Da else None
    with torch.cuda.device(dtable.device):
        _lib.check(lib.vqb_table_backward(ptr(dtable), ptr(table), ptr(colsum), ptr(phn_attr), K, D, A, Da,
                                          ptr(d_learn), ptr(d_w), ptr(d_b), _stream(dtable)))
    if tail is not None and tail.exchange is not None:
        from .dist import reduce_route_output
        reduce_route_output(flat, tail)
    return d_learn, d_w, d_b


def _lengths_arg(lengths, n_utts, dev):
    """[B] int64 on `dev` (accepts a CPU / int32 tensor or a list), checked against the batch"""
    if lengths is None:
        return None
    t = torch.as_tensor(lengths)
    if t.numel() != n_utts:
        raise RuntimeError("semi-tts_b200: `lengths` must hold one entry per utterance (%d), got %d" % (n_utts, t.numel()))
    return t.to(device=dev, dtype=torch.int64).contiguous()


def _run_forward(flags, x2d, score_w, score_b, gather_table, temp, want_pcode, hist, want_sqerr,
                 score_w_bf16=None, search_stats=None, operand_cache=None, lengths=None, frames=0, ctc_eps=None):
    """`ctc_eps` (with `frames`): also ask the parity-mode kernel for log(p_code + ctc_eps) in nn.CTCLoss's [S, B, K] layout
    (returned as a sixth value; None when the route taken cannot emit it)"""
    lib = _lib.load()
    N, D = x2d.shape
    K = score_w.shape[0]
    dev = x2d.device
    p_code = torch.empty(N, K, device=dev, dtype=torch.float32) if want_pcode else None
    idx = torch.empty(N, device=dev, dtype=torch.int64)
    q = torch.empty(N, D, device=dev, dtype=torch.float32)
    sq = torch.zeros(1, device=dev, dtype=torch.float64) if want_sqerr else None
    a = _lib.FwdArgs()
    a.struct_size = ctypes.sizeof(_lib.FwdArgs)
    a.flags = flags
    a.n_rows, a.dim, a.n_codes = N, D, K
    a.x, a.score_w, a.score_b, a.gather_table = ptr(x2d), ptr(score_w), ptr(score_b), ptr(gather_table)
    a.score_w_bf16 = ptr(score_w_bf16)
    a.temp, a.p_code, a.idx, a.new_latent = ptr(temp), ptr(p_code), ptr(idx), ptr(q)
    a.hist, a.sq_err_sum, a.search_stats = ptr(hist), ptr(sq), ptr(search_stats)
    a.operand_cache = ptr(operand_cache)
    a.row_lengths, a.frames_per_utt = ptr(lengths), (frames if lengths is not None else 0)
    logp = None
    if ctc_eps is not None and N > 0 and frames > 0 and lib.vqb_forward_kernel_name(ctypes.byref(a)) == b"vqb_fwd_pcode_kernel":
        logp = torch.empty(frames, N // frames, K, device=dev, dtype=torch.float32)
        a.frames_per_utt, a.ctc_logp, a.ctc_eps = frames, ptr(logp), float(ctc_eps)
    with torch.cuda.device(dev):
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.vqb_forward_workspace(ctypes.byref(a), ctypes.byref(nbytes)))
        ws = torch.empty(nbytes.value, device=dev, dtype=torch.uint8) if nbytes.value else None
        a.workspace, a.workspace_bytes = ptr(ws), nbytes.value
        _lib.check(lib.vqb_forward(ctypes.byref(a), _stream(x2d)))
    if ctc_eps is not None:
        return p_code, idx, q, sq, logp
    return p_code, idx, q, sq


class NoGradCache:
    """Per-module state of the no-grad fast path (validation and encode loops: bin/train_vqvae.py:343-346, bin/gen_specgram.py:
    95-108): the assembled table, |e|^2 and the operand image are kept until a parameter changes (data pointer or in-place
    version counter), the argument struct is built once, and the call goes straight to the C ABI -- no autograd.Function,
    no per-call table assembly, no workspace query.  What is left per call is three output allocations and one launch."""

    def __init__(self):
        self.key = None
        self.table = self.enorm = self.image = None
        self.args = None
        self.ws = None
        self.fn = self.ref = self.last = None

    def __deepcopy__(self, memo):
        return NoGradCache()                      # a copied module rebuilds its cache (the struct holds raw device pointers)

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()

    @staticmethod
    def _key(learnable, phn_attr, proj_w, proj_b, want_image):
        # optimizer steps and load_state_dict write in place (version counters), .to() / .data = ... move the storage
        if proj_w is None:
            return (learnable.data_ptr(), learnable._version, want_image)
        return (learnable.data_ptr(), learnable._version, proj_w.data_ptr(), proj_w._version, proj_b._version,
                phn_attr.data_ptr(), phn_attr._version, want_image)

    def tables(self, learnable, phn_attr, proj_w, proj_b, want_image):
        key = self._key(learnable, phn_attr, proj_w, proj_b, bool(want_image))
        if key != self.key:
            res = assemble_table(learnable, phn_attr, proj_w, proj_b, want_cache=want_image)
            self.table, self.enorm = res[0], res[1]
            self.image = res[3] if want_image else None
            self.key = key
            self.args = None
        return self.table, self.enorm, self.image


def forward_nograd(cache, x, learnable, phn_attr, proj_w, proj_b, temp, skip, want_pcode, hist, tensor_cores, lengths=None):
    """L2 quantizer forward without autograd (src/embed.py:105-147 under torch.no_grad()).
    Returns (p_code[B,S,K] or None, new_latent[B,S,D], idx[B,S]).
    The per-call host work is kept to: three output allocations in their final shape, six struct fields, one C call."""
    if not x.is_cuda or x.dtype != torch.float32:
        _require(x, "enc_embs")
    if x.dim() != 3:
        raise RuntimeError("semi-tts_b200: enc_embs must be [B, S, D]")
    B, S, D = x.shape
    if not x.is_contiguous():
        x = x.contiguous()
    use_image = bool(tensor_cores and want_pcode)
    table, enorm, image = cache.tables(learnable, phn_attr, proj_w, proj_b, use_image)
    K = table.shape[0]
    if table.shape[1] != D:
        raise RuntimeError("semi-tts_b200: enc_embs has D=%d but the codebook has D=%d" % (D, table.shape[1]))
    dev = x.device
    p_code = torch.empty((B, S, K), device=dev, dtype=torch.float32) if want_pcode else None
    idx = torch.empty((B, S), device=dev, dtype=torch.int64)
    q = torch.empty((B, S, D), device=dev, dtype=torch.float32)
    a = cache.args
    if a is None:
        a = _lib.FwdArgs()
        a.struct_size = ctypes.sizeof(_lib.FwdArgs)
        a.dim, a.n_codes = D, K
        a.score_w, a.score_b, a.gather_table = ptr(table), ptr(enorm), ptr(table)
        a.operand_cache = ptr(image)
        cache.args = a
        cache.ws = None
        cache.fn = _lib.load().vqb_forward
        cache.ref = ctypes.byref(a)
        cache.last = None
    # fields that rarely change between calls are written only when they do
    var = (B * S, skip, tensor_cores, temp.data_ptr(), None if hist is None else hist.data_ptr(), lengths is None)
    if var != cache.last:
        a.flags = _lib.SCORE_L2 | _lib.STOP_GRAD | (_lib.SKIP if skip else 0) | (_lib.TENSOR_CORES if tensor_cores else 0)
        a.n_rows = B * S
        a.temp, a.hist = ptr(temp), ptr(hist)
        a.row_lengths, a.frames_per_utt = None, 0
        cache.last = var
    a.x, a.idx, a.new_latent = x.data_ptr(), idx.data_ptr(), q.data_ptr()
    a.p_code = p_code.data_ptr() if want_pcode else None
    if lengths is not None:
        lens = _lengths_arg(lengths, B, dev)
        a.row_lengths, a.frames_per_utt = ptr(lens), S
    with _on(dev):
        if not (use_image and image is not None):
            # shapes outside the cached-image route (fused search, CUDA-core kernels) may need scratch: sized once per module
            if cache.ws is None:
                nbytes = ctypes.c_size_t(0)
                _lib.check(_lib.load().vqb_forward_workspace(cache.ref, ctypes.byref(nbytes)))
                cache.ws = torch.empty(max(nbytes.value, 1), device=dev, dtype=torch.uint8)
            a.workspace, a.workspace_bytes = ptr(cache.ws), cache.ws.numel()
        rc = cache.fn(cache.ref, _stream(x))
        if rc:
            _lib.check(rc)
    return p_code, q, idx


def lookup_nograd(cache, txt, learnable, phn_attr, proj_w, proj_b):
    """inference(txt) without autograd (src/embed.py:96-103 under torch.no_grad()): a gather from the cached table."""
    if not txt.is_cuda or txt.dtype != torch.int64:
        _require(txt, "txt", torch.int64)
    want_image = cache.key[-1] if cache.key is not None else False
    table, _, _ = cache.tables(learnable, phn_attr, proj_w, proj_b, want_image)
    K, D = table.shape
    t = txt if txt.is_contiguous() else txt.contiguous()
    out = torch.empty(t.shape + (D,), device=t.device, dtype=torch.float32)
    with _on(t.device):
        rc = _lib.load().vqb_inference_gather(t.data_ptr(), t.numel(), table.data_ptr(), K, D, out.data_ptr(), _stream(t))
        if rc:
            _lib.check(rc)
    return out


class FusedTail:
    """Per-module state of the fused backward tail (include/vqb.h: vqb_bwd_tail): the ticket / epoch words and, in
    data-parallel runs, the peer-mapped exchange buffers (dist.enable_fused_allreduce).  `fused` records whether the
    last backward took the fused route (then the parameter gradients are already summed over the group)."""

    def __init__(self):
        import os
        self.counter = None          # uint32 [2] on the module's device
        self.exchange = None         # dist.PeerExchange or None
        self.fused = False
        self.enabled = not os.environ.get("VQB_NO_TAIL_TEST")     # developer switch
        # deferred exchange (data-parallel runs): the backward's tail leaves this rank's sums in the flat gradient; the
        # exchange kernel (push to the peers, poll, rank-ordered sum) runs in finish() -- called by dist.allreduce_codebook_grads / dist.finish_codebook_grads,
        # i.e. behind the rest of the model's backward -- or at the latest before the module's next backward
        self.defer = False
        self.pending = None          # (BwdTail struct, flat tensor, n_flat) of the exchange that still has to be finished

    def finish(self, stream=None):
        """second half of a deferred exchange: after this (stream-ordered) the parameter gradients of the last backward
        hold the sum over all ranks"""
        if self.pending is None:
            return
        tl, flat, n_flat = self.pending
        self.pending = None
        lib = _lib.load()
        st = stream if stream is not None else torch.cuda.current_stream(flat.device)
        if stream is not None:
            flat.record_stream(stream)         # the allocator must not hand the block out again while the side stream uses it
        with _on(flat.device):
            _lib.check(lib.vqb_exchange_finish(ctypes.byref(tl), n_flat, ctypes.c_void_p(st.cuda_stream)))

    def counter_for(self, dev):
        # [0] block ticket, [1] epoch of the exchange, [2] error flag (1 + the rank a timed-out exchange waited for), [3] spare
        if self.counter is None or self.counter.device != dev:
            self.counter = torch.zeros(4, device=dev, dtype=torch.int32)
        return self.counter


def _exchange_timeout_ms():
    """How long the in-kernel exchange waits for a peer before it gives up and raises the module's error flag
    (dist.check_exchange); comparable to NCCL's watchdog rather than to a step time: rank-0 validation, checkpointing or a
    data-loader stall must not kill the job."""
    import os
    return int(os.environ.get("VQB_EXCHANGE_TIMEOUT_MS", "120000"))


def _run_backward(flags, n_real_rows, x2d, score_w, score_b, gather_table, temp, p_code, idx, g_p, g_q,
                  want_dx_buffer, separate_gather, operand_cache=None, tail=None, phn_attr=None, Da=0, lengths=None, frames=0,
                  g_logp=None, ctc_eps=0.0):
    """`g_logp` [S, B, K] (instead of g_p): the upstream gradient of the forward's ctc_logp output, folded into the kernel.
    Returns (dx or None, d_score_w, colsum, d_gather or None, d_temp or None, flat or None).
    `flat` is set when the fused tail ran: [d_learnable | d_proj_w | d_proj_b], already summed over the group."""
    lib = _lib.load()
    N, D = x2d.shape
    K = score_w.shape[0]
    dev = x2d.device
    n_g = K * D if separate_gather else 0
    a = _lib.BwdArgs()
    a.struct_size = ctypes.sizeof(_lib.BwdArgs)
    a.flags = flags
    a.n_rows, a.dim, a.n_codes, a.n_real_rows = N, D, K, n_real_rows
    a.x, a.score_w, a.score_b, a.gather_table, a.temp = ptr(x2d), ptr(score_w), ptr(score_b), ptr(gather_table), ptr(temp)
    a.p_code, a.idx, a.g_p, a.g_q = ptr(p_code), ptr(idx), ptr(g_p), ptr(g_q)
    a.operand_cache = ptr(operand_cache)
    a.row_lengths, a.frames_per_utt = ptr(lengths), (frames if lengths is not None else 0)
    if g_logp is not None:
        a.frames_per_utt, a.g_logp, a.ctc_eps = frames, ptr(g_logp), float(ctc_eps)
        a.g_p = None
        if g_p is not None or lib.vqb_backward_kernel_name(ctypes.byref(a)) != b"vqb_bwd_pcode_kernel":
            # this call is not served by the kernel that folds the division in (or g_p is there as well): one extra pass
            # turns g_logp into (an addition to) g_p
            acc = g_p is not None
            g_p = g_p.clone() if acc else torch.empty(N, K, device=dev, dtype=torch.float32)
            with torch.cuda.device(dev):
                _lib.check(lib.vqb_ctc_logp_backward(ptr(g_logp), ptr(p_code), N // frames, frames, K, float(ctc_eps), ptr(g_p),
                                                     1 if acc else 0, _stream(x2d)))
            a.g_logp, a.ctc_eps, a.frames_per_utt = None, 0.0, (frames if lengths is not None else 0)
        a.g_p = ptr(g_p)
    use_tail = bool(tail is not None and tail.enabled and (flags & _lib.SCORE_L2) and not separate_gather
                    and (lib.vqb_backward_kernel_name(ctypes.byref(a)) == b"vqb_bwd_pcode_kernel"
                         or (N == 0 and tail.exchange is not None)))      # an empty shard still joins its peers' exchange
    flat = tl = None
    if use_tail:
        # the tail overwrites its scratch and outputs: nothing to zero-fill
        A = phn_attr.shape[1] if phn_attr is not None else 0
        Da = Da if phn_attr is not None else 0
        n_flat = K * (D - Da) + Da * A + Da
        flat = torch.empty(n_flat, device=dev, dtype=torch.float32)
        zeros = None
        tl = _lib.BwdTail()
        tl.phn_attr, tl.n_attr, tl.dim_attr, tl.d_flat = ptr(phn_attr), A, Da, ptr(flat)
        tl.counter = ptr(tail.counter_for(dev))
        ex = tail.exchange
        tl.world, tl.rank, tl.peer_bufs = (ex.world, ex.rank, ex.peer_ptrs_dev(n_flat)) if ex is not None else (1, 0, None)
        tl.timeout_ms = _exchange_timeout_ms()
        if ex is not None and tail.defer:
            tail.finish()                                  # (an unfinished exchange of the previous step: finish it first)
            tl.reserved = 1                                # VQB_TAIL_DEFER: local sums now, the exchange in finish()
        a.tail = ctypes.pointer(tl)
    else:
        # one zero-filled buffer (one fill kernel) carved into the accumulation targets
        zeros = torch.zeros(K * D + n_g + K + 4, device=dev, dtype=torch.float32)
    if zeros is not None:
        d_w = zeros[:K * D].view(K, D)
        d_gather = zeros[K * D:K * D + n_g].view(K, D) if separate_gather else None
        colsum = zeros[K * D + n_g:K * D + n_g + K]
        d_temp = zeros[K * D + n_g + K:K * D + n_g + K + 1] if flags & _lib.TEMP_GRAD else None
    else:
        d_w = d_gather = colsum = d_temp = None           # the fused tail writes the parameter gradients directly
    dx = torch.empty(N, D, device=dev, dtype=torch.float32) if want_dx_buffer else None
    a.dx, a.d_score_w, a.colsum, a.d_gather, a.d_temp = ptr(dx), ptr(d_w), ptr(colsum), ptr(d_gather), ptr(d_temp)
    with torch.cuda.device(dev):
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.vqb_backward_workspace(ctypes.byref(a), ctypes.byref(nbytes)))
        ws = torch.empty(nbytes.value, device=dev, dtype=torch.uint8) if nbytes.value else None
        a.workspace, a.workspace_bytes = ptr(ws), nbytes.value
        _lib.check(lib.vqb_backward(ctypes.byref(a), _stream(x2d)))
    if tail is not None:
        tail.fused = use_tail
        if use_tail and tl.reserved:
            tail.pending = (tl, flat, n_flat)
    return dx, d_w, colsum, d_gather, d_temp, flat


def _loss_backward(x2d, table, idx, g_vq, g_commit, dx, dx_accumulate, dtable):
    lib = _lib.load()
    N, D = x2d.shape
    with torch.cuda.device(x2d.device):
        _lib.check(lib.vqb_loss_backward(ptr(x2d), ptr(table), ptr(idx), N, D, table.shape[0], ptr(g_vq),
                                         ptr(g_commit), ptr(dx), 1 if dx_accumulate else 0, ptr(dtable),
                                         _stream(x2d)))


class _Cfg:
    """Per-call options (plain Python, not a tensor)."""
    __slots__ = ("stop_grad", "skip", "n_real_rows", "want_pcode", "hist", "want_losses", "tensor_cores", "tail", "lengths", "ctc_eps")

    def __init__(self, stop_grad=True, skip=False, n_real_rows=0, want_pcode=True, hist=None,
                 want_losses=False, tensor_cores=True, tail=None, lengths=None, ctc_eps=None):
        self.tail = tail
        self.lengths = lengths
        self.ctc_eps = None if ctc_eps is None else float(ctc_eps)
        self.stop_grad, self.skip, self.n_real_rows = bool(stop_grad), bool(skip), int(n_real_rows)
        self.want_pcode, self.hist, self.want_losses = bool(want_pcode), hist, bool(want_losses)
        self.tensor_cores = bool(tensor_cores)


def _g32(t):
    return None if t is None else _c(t.to(torch.float32))


# ------------------------------------------------------------------------------------------------
# L2 quantizer
# ------------------------------------------------------------------------------------------------
def _fwd_flags(score, cfg):
    return score | (_lib.STOP_GRAD if cfg.stop_grad else 0) | (_lib.SKIP if cfg.skip else 0)


class _VQL2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, learnable, phn_attr, proj_w, proj_b, temp, cfg):
        _require(x, "enc_embs"); _require(temp, "temp")
        if x.dim() != 3:
            raise RuntimeError("semi-tts_b200: enc_embs must be [B, S, D]")
        B, S, D = x.shape
        x2d = _c(x.detach()).view(B * S, D)
        want_cache = cfg.tensor_cores and cfg.want_pcode
        res = assemble_table(learnable, phn_attr, proj_w, proj_b, want_cache=want_cache)
        table, enorm, tbf = res[:3]
        cache = res[3] if want_cache else None
        K = table.shape[0]
        if table.shape[1] != D:
            raise RuntimeError("semi-tts_b200: enc_embs has D=%d but the codebook has D=%d" % (D, table.shape[1]))
        if not cfg.want_pcode and not cfg.stop_grad:
            raise RuntimeError("semi-tts_b200: the ST-onehot variant (stop_grad=False) needs p_code")
        flags = _fwd_flags(_lib.SCORE_L2, cfg) | (_lib.TENSOR_CORES if cfg.tensor_cores else 0)
        temp_c = _c(temp.detach())
        if want_cache:
            # only kernels (the assembly above, at most a one-element fill) sit between here and the forward kernel
            flags |= _lib.AFTER_ASSEMBLE
        lens = _lengths_arg(cfg.lengths, B, x2d.device)
        if lens is not None and (cfg.want_losses or not cfg.want_pcode):
            raise RuntimeError("semi-tts_b200: `lengths` is served by the parity-mode route only (p_code on, no loss extensions)")
        logp = None
        if cfg.ctc_eps is not None:
            if not cfg.want_pcode:
                raise RuntimeError("semi-tts_b200: ctc_eps needs p_code (fused_search off)")
            p_code, idx, q, sq, logp = _run_forward(flags, x2d, table, enorm, table, temp_c, True, cfg.hist,
                                                    cfg.want_losses, tbf, None, cache, lens, S, cfg.ctc_eps)
            if logp is None:
                # a route without the fused emission (K > 64, SIMT, ...): the standalone pass over p_code
                logp = torch.empty(S, B, K, device=x2d.device, dtype=torch.float32)
                if B * S:
                    with torch.cuda.device(x2d.device):
                        _lib.check(_lib.load().vqb_ctc_logp(ptr(p_code), B, S, K, cfg.ctc_eps, ptr(logp), _stream(x2d)))
        else:
            p_code, idx, q, sq = _run_forward(flags, x2d, table, enorm, table, temp_c, cfg.want_pcode, cfg.hist,
                                              cfg.want_losses, tbf, None, cache, lens, S)
        ctx.lens = lens
        ctx.op_cache = cache
        ctx.set_materialize_grads(False)                    # an unused output must arrive as None, not zeros
        ctx.cfg, ctx.shape = cfg, (B, S, D, K)
        ctx.Da = proj_w.shape[0] if phn_attr is not None else 0
        ctx.temp_grad = bool(temp.requires_grad)
        ctx.save_for_backward(x2d, table, enorm, temp_c, p_code, idx, phn_attr)
        idx3 = idx.view(B, S)
        ctx.mark_non_differentiable(idx3)
        vq = commit = None
        if cfg.want_losses:
            vq = (sq / float(max(B * S * D, 1))).to(torch.float32).view(())
            commit = vq.clone()
        return (p_code.view(B, S, K) if p_code is not None else None), q.view(B, S, D), idx3, vq, commit, logp

    @staticmethod
    def backward(ctx, g_p, g_q, _g_idx, g_vq, g_commit, g_logp=None):
        x2d, table, enorm, temp, p_code, idx, phn_attr = ctx.saved_tensors
        cfg = ctx.cfg
        B, S, D, K = ctx.shape
        N = B * S
        have_loss = g_vq is not None or g_commit is not None
        if g_p is None and g_q is None and g_logp is None and not have_loss:
            return (None,) * 7
        dev = x2d.device
        exchange_on = cfg.tail is not None and cfg.tail.exchange is not None
        if N == 0 and not exchange_on:
            return (None,) * 7
        g_p2 = _g32(g_p).view(N, K) if g_p is not None else None
        g_q2 = _g32(g_q).view(N, D) if g_q is not None else None
        g_l = _g32(g_logp) if (g_logp is not None and N > 0) else None      # [S, B, K]
        if g_logp is not None and N == 0 and g_p2 is None:
            g_p2 = torch.zeros(0, K, device=dev, dtype=torch.float32)
        flags = _fwd_flags(_lib.SCORE_L2, cfg) | (_lib.TEMP_GRAD if ctx.temp_grad else 0) | \
            (_lib.TENSOR_CORES if cfg.tensor_cores else 0)
        d_temp = colsum = flat = None
        if cfg.tail is not None:
            cfg.tail.fused = False
        if N == 0:
            # an empty shard of a data-parallel run still takes part in the exchange its peers run
            dx = torch.zeros(0, D, device=dev, dtype=torch.float32)
            if g_p2 is not None and not have_loss:
                _, d_w, colsum, _, d_temp, flat = _run_backward(
                    flags, 0, x2d, table, enorm, table, temp, p_code, idx, g_p2, g_q2, False, False, ctx.op_cache,
                    tail=cfg.tail, phn_attr=phn_attr, Da=ctx.Da)
            if flat is None:
                d_w = torch.zeros(K, D, device=dev, dtype=torch.float32)
        elif g_p2 is None and g_q2 is None and g_l is None:
            dx, d_w = None, torch.zeros(K, D, device=dev, dtype=torch.float32)
        elif g_p2 is None and g_l is None and cfg.stop_grad and ctx.lens is None:
            # scatter-only: dx = g_q, the straight-through identity, returned as the same tensor (zero bytes)
            _, d_w, _, _, _, _ = _run_backward(flags & ~_lib.TEMP_GRAD, cfg.n_real_rows, x2d, table, enorm, table, temp,
                                               None, idx, None, g_q2, False, False)
            dx = g_q2
        else:
            if ctx.lens is not None and g_p2 is None and g_l is None:
                # (the tensor-core backward is keyed on g_p; a step that only back-propagates through new_latent gets a zero g_p)
                g_p2 = torch.zeros(N, K, device=dev, dtype=torch.float32)
            dx, d_w, colsum, _, d_temp, flat = _run_backward(
                flags, cfg.n_real_rows, x2d, table, enorm, table, temp, p_code, idx, g_p2, g_q2, True, False,
                ctx.op_cache, tail=None if have_loss else cfg.tail, phn_attr=phn_attr, Da=ctx.Da, lengths=ctx.lens, frames=S,
                g_logp=g_l, ctc_eps=cfg.ctc_eps or 0.0)
        if flat is not None:
            # fused tail: table backward (and the sum over GPUs) already done behind the main kernel
            Da, A = (ctx.Da, phn_attr.shape[1]) if phn_attr is not None else (0, 0)
            n_l, n_w = K * (D - Da), Da * A
            return dx.view(B, S, D), flat[:n_l].view(K, D - Da), None, \
                (flat[n_l:n_l + n_w].view(Da, A) if Da else None), (flat[n_l + n_w:] if Da else None), None, None
        if have_loss and N > 0:
            if dx is None:
                dx, acc = torch.empty(N, D, device=dev, dtype=torch.float32), False
            elif dx is g_q2:
                dx, acc = g_q2.clone(), True
            else:
                acc = True
            _loss_backward(x2d, table, idx, _g32(g_vq), _g32(g_commit), dx, acc, d_w)
        d_learn, d_pw, d_pb = _table_backward(d_w, table, colsum, phn_attr, ctx.Da, cfg.tail)
        if ctx.temp_grad and exchange_on and d_temp is not None:
            from .dist import reduce_route_output
            reduce_route_output(d_temp, cfg.tail)
        if ctx.temp_grad and d_temp is None:
            d_temp = torch.zeros(1, device=dev, dtype=torch.float32)
        return (dx.view(B, S, D) if dx is not None else None), d_learn, None, d_pw, d_pb, \
            (d_temp if ctx.temp_grad else None), None


def vq_l2(x, learnable_table, phn_attr, proj_w, proj_b, temp, stop_grad=True, skip=False, n_real_rows=0,
          want_pcode=True, hist=None, want_losses=False, tensor_cores=True, tail=None, lengths=None, ctc_eps=None):
    """L2 quantizer (src/embed.py:105-147).
    Returns (p_code[B,S,K] or None, new_latent[B,S,D], idx[B,S] int64, vq_loss or None, commit_loss or None) and, with
    `ctc_eps`, a sixth value: log(p_code + ctc_eps) as contiguous [S,B,K] -- the nn.CTCLoss input of
    bin/train_vqvae.py:430-432, written by the forward kernel's epilogue, its gradient folded into the backward kernel."""
    cfg = _Cfg(stop_grad, skip, n_real_rows, want_pcode, hist, want_losses, tensor_cores, tail, lengths, ctc_eps)
    out = _VQL2.apply(x, learnable_table, phn_attr, proj_w, proj_b, temp, cfg)
    return out if ctc_eps is not None else out[:5]


def vq_search(x, table, temp=None, hist=None, search_tensor=True, stats=None):
    """Fused-mode forward on a ready-made table (no p_code, no autograd): nearest-codeword search +
    gather + straight-through.  x[N,D] or [B,S,D], table[K,D] -> (idx int64, new_latent)."""
    _require(x, "x"); _require(table, "table")
    shape = x.shape
    x2d = _c(x.detach()).view(-1, shape[-1])
    table = _c(table.detach())
    if temp is None:
        temp = torch.ones(1, device=x.device, dtype=torch.float32)
    tab, enorm, _ = assemble_table(table)
    flags = _lib.SCORE_L2 | _lib.STOP_GRAD | (_lib.SEARCH_TENSOR if search_tensor else 0)
    if stats is not None and (stats.dtype != torch.int32 or stats.numel() < 2 or not stats.is_cuda):
        raise RuntimeError("semi-tts_b200: `stats` must be a CUDA int32 tensor with 2 elements")
    _, idx, q, _ = _run_forward(flags, x2d, tab, enorm, tab, temp, False, hist, False, None, stats)
    return idx.view(shape[:-1]), q.view(shape)


# ------------------------------------------------------------------------------------------------
# "separate" quantizer
# ------------------------------------------------------------------------------------------------
class _VQLinear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, asr_w, asr_b, emb_w, phn_attr, proj_w, proj_b, cfg):
        _require(x, "enc_embs"); _require(asr_w, "asr_final_layer.weight"); _require(asr_b, "asr_final_layer.bias")
        if x.dim() != 3:
            raise RuntimeError("semi-tts_b200: enc_embs must be [B, S, D]")
        B, S, D = x.shape
        x2d = _c(x.detach()).view(B * S, D)
        table, _, _ = assemble_table(emb_w, phn_attr, proj_w, proj_b)
        K = asr_w.shape[0]
        if table.shape != (K, D) or asr_w.shape[1] != D:
            raise RuntimeError("semi-tts_b200: shape mismatch between enc_embs, asr_final_layer and the embedding table")
        w, b = _c(asr_w.detach()), _c(asr_b.detach())
        flags = _fwd_flags(_lib.SCORE_LINEAR, cfg) | (_lib.TENSOR_CORES if cfg.tensor_cores else 0)
        logp = None
        if cfg.ctc_eps is not None:
            p_code, idx, q, _, logp = _run_forward(flags, x2d, w, b, table, None, True, cfg.hist, False, frames=S,
                                                   ctc_eps=cfg.ctc_eps)
            if logp is None:                            # a route without the fused emission: the standalone pass
                logp = torch.empty(S, B, K, device=x2d.device, dtype=torch.float32)
                if B * S:
                    with torch.cuda.device(x2d.device):
                        _lib.check(_lib.load().vqb_ctc_logp(ptr(p_code), B, S, K, cfg.ctc_eps, ptr(logp), _stream(x2d)))
        else:
            p_code, idx, q, _ = _run_forward(flags, x2d, w, b, table, None, True, cfg.hist, False)
        ctx.set_materialize_grads(False)
        ctx.cfg, ctx.shape = cfg, (B, S, D, K)
        ctx.Da = proj_w.shape[0] if phn_attr is not None else 0
        ctx.save_for_backward(x2d, w, table, p_code, idx, phn_attr)
        idx3 = idx.view(B, S)
        ctx.mark_non_differentiable(idx3)
        return p_code.view(B, S, K), q.view(B, S, D), idx3, logp

    @staticmethod
    def backward(ctx, g_p, g_q, _g_idx, g_logp=None):
        x2d, w, table, p_code, idx, phn_attr = ctx.saved_tensors
        cfg = ctx.cfg
        B, S, D, K = ctx.shape
        N = B * S
        if (g_p is None and g_q is None and g_logp is None) or N == 0:
            return (None,) * 8
        g_p2 = _g32(g_p).view(N, K) if g_p is not None else None
        g_q2 = _g32(g_q).view(N, D) if g_q is not None else None
        g_l = _g32(g_logp) if g_logp is not None else None
        flags = _fwd_flags(_lib.SCORE_LINEAR, cfg) | (_lib.TENSOR_CORES if cfg.tensor_cores else 0)
        dx, d_w, colsum, d_tab, _, _ = _run_backward(flags, 0, x2d, w, None, table, None, p_code, idx, g_p2, g_q2,
                                                     True, True, frames=S, g_logp=g_l, ctc_eps=cfg.ctc_eps or 0.0)
        d_emb, d_pw, d_pb = _table_backward(d_tab, None, None, phn_attr, ctx.Da, cfg.tail)
        if cfg.tail is not None and cfg.tail.exchange is not None:
            from .dist import reduce_route_output
            reduce_route_output(d_w, cfg.tail)
            reduce_route_output(colsum, cfg.tail)
        return dx.view(B, S, D), d_w, colsum, d_emb, None, d_pw, d_pb, None


def vq_linear(x, asr_w, asr_b, emb_w, phn_attr, proj_w, proj_b, stop_grad=True, hist=None, tensor_cores=True, tail=None,
              ctc_eps=None):
    """Separate quantizer (src/embed.py:187-205). Returns (p_code, new_latent, idx) and, with `ctc_eps`, a fourth value:
    log(p_code + ctc_eps) as contiguous [S,B,K] (see vq_l2)."""
    cfg = _Cfg(stop_grad=stop_grad, hist=hist, tensor_cores=tensor_cores, tail=tail, ctc_eps=ctc_eps)
    out = _VQLinear.apply(x, asr_w, asr_b, emb_w, phn_attr, proj_w, proj_b, cfg)
    return out if ctc_eps is not None else out[:3]


# ------------------------------------------------------------------------------------------------
# gather-only inference path
# ------------------------------------------------------------------------------------------------
class _Lookup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, txt, learnable, phn_attr, proj_w, proj_b, tail):
        _require(txt, "txt", torch.int64)
        lib = _lib.load()
        table, _, _ = assemble_table(learnable, phn_attr, proj_w, proj_b)
        K, D = table.shape
        t = _c(txt)
        out = torch.empty(*t.shape, D, device=t.device, dtype=torch.float32)
        with torch.cuda.device(t.device):
            _lib.check(lib.vqb_inference_gather(ptr(t), t.numel(), ptr(table), K, D, ptr(out), _stream(t)))
        ctx.set_materialize_grads(False)
        ctx.Da = proj_w.shape[0] if phn_attr is not None else 0
        ctx.tail = tail
        ctx.save_for_backward(t, table, phn_attr)
        return out

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return (None,) * 6
        t, table, phn_attr = ctx.saved_tensors
        lib = _lib.load()
        K, D = table.shape
        g2 = _g32(g).view(-1, D)
        dtab = torch.zeros(K, D, device=g2.device, dtype=torch.float32)
        with torch.cuda.device(g2.device):
            nb = ctypes.c_size_t(0)
            _lib.check(lib.vqb_scatter_workspace(t.numel(), K, D, ctypes.byref(nb)))
            ws = torch.empty(nb.value, device=g2.device, dtype=torch.uint8) if nb.value else None
            _lib.check(lib.vqb_scatter_add(ptr(t), t.numel(), ptr(g2), K, D, ptr(dtab), None, ptr(ws), nb.value, _stream(g2)))
        if ctx.tail is not None and ctx.tail.exchange is not None and ctx.tail.defer:
            raise RuntimeError("semi-tts_b200: a deferred gradient exchange (fused_tail.defer) cannot be combined with gradients "
                               "through inference(): the forward route's flat gradient is incomplete until finish(); "
                               "set module.fused_tail.defer = False for steps that train the codebook through text")
        d_learn, d_pw, d_pb = _table_backward(dtab, None, None, phn_attr, ctx.Da, ctx.tail)
        return None, d_learn, None, d_pw, d_pb, None


def codebook_lookup(txt, learnable_table, phn_attr=None, proj_w=None, proj_b=None, tail=None):
    """inference(txt): table[txt] on the assembled table (src/embed.py:96-103, :180-185), differentiable
    w.r.t. learnable_table / proj_attr (the text->speech branch trains the codebook through it).  `tail`: the module's
    FusedTail; with a fused exchange attached the lookup's gradient is summed over the data-parallel group in its backward."""
    return _Lookup.apply(txt, learnable_table, phn_attr, proj_w, proj_b, tail)
