"""Multi-GPU plumbing: frames shard by batch, the codebook is replicated, and the only exchange is ONE
all-reduce of a flat fp32 buffer [codebook-side gradients | usage histogram].
The reference has no distributed code at all; this is the data-parallel layer around the drop-in module.
Works with any torch.distributed backend (NCCL over NVLink on the B200 box, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous, balanced [lo, hi) range of `n_items` whole utterances for `rank`."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _flat_view(grads):
    """If the gradients are views of one storage that tile a contiguous span of it (as the backward of this package
    produces them: d_learnable | d_proj_w | d_proj_b, whatever order module.parameters() lists them in), return that span
    as a single 1-D tensor; otherwise None."""
    if not grads:
        return None
    first = grads[0]
    try:
        base_ptr = first.untyped_storage().data_ptr()
        for g in grads:
            if not g.is_contiguous() or g.dtype != first.dtype or g.untyped_storage().data_ptr() != base_ptr:
                return None
    except Exception:                                   # noqa: BLE001
        return None
    order = sorted(grads, key=lambda g: g.storage_offset())
    off = order[0].storage_offset()
    for g in order:
        if g.storage_offset() != off:                   # gap or overlap
            return None
        off += g.numel()
    n = off - order[0].storage_offset()
    return torch.as_strided(order[0], (n,), (1,), order[0].storage_offset())


def reduce_route_output(flat, tail, group=None):
    """Called by the autograd backward of every route that does NOT end in the fused tail kernel (inference() lookups, the
    scatter-only route, the loss extensions, shapes outside the tensor-core kernels) when the module has a fused exchange
    attached: the route's flat parameter gradient is summed over the group right here, with one NCCL all-reduce, before
    autograd accumulates it into p.grad.  With an exchange attached every contribution to p.grad is therefore already a
    global sum -- whatever mixture of routes a step takes -- and allreduce_codebook_grads() has nothing left to add."""
    if tail is None or tail.exchange is None or flat is None or isinstance(tail.exchange, LoopbackExchange):
        return
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(tail.exchange.group) > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=tail.exchange.group)


class PeerExchange:
    """Peer-mapped exchange buffers for the fused backward tail (include/vqb.h: vqb_bwd_tail): one buffer per rank
    in torch symmetric memory (CUDA VMM handles shared at rendezvous, mapped over NVLink), zero-filled once; the device
    array of the world's buffer addresses is what the kernel receives.  No NCCL call is made per step."""

    def __init__(self, group=None, max_floats=1 << 16, device=None):
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        group = group if group is not None else dist.group.WORLD
        self.group = group
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 16:
            raise RuntimeError("semi-tts_b200: the fused exchange supports up to 16 GPUs of one NVLink domain")
        device = torch.device("cuda", torch.cuda.current_device()) if device is None else device
        self.max_floats = int(max_floats)
        nbytes = _lib.load().vqb_exchange_bytes(self.max_floats, self.world)
        self.buf = symm.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm.rendezvous(self.buf, group)
        self.ptrs = torch.tensor([int(p) for p in self.handle.buffer_ptrs], dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group)                      # every rank's zero-fill has landed before anyone raises a flag

    def peer_ptrs_dev(self, n_flat):
        if n_flat > self.max_floats:
            raise RuntimeError("semi-tts_b200: exchange buffer holds %d floats, the gradient has %d" % (self.max_floats, n_flat))
        return self.ptrs.data_ptr()


class LoopbackExchange:
    """Single-GPU emulation of a `world`-rank exchange for tests: every emulated rank's exchange buffer lives in THIS GPU's
    memory and the `peer` pointers are plain device pointers, so the same tail kernel, the same (value, epoch) words and the
    same polling loop run as over NVLink.  One instance per emulated rank, all built from one shared buffer set:
        bufs = LoopbackExchange.make_buffers(world, n_floats, device)
        ex_r = LoopbackExchange(bufs, rank=r)            # attach to module replica r: m_r.fused_tail.exchange = ex_r
    The replicas' backward passes must run CONCURRENTLY (different streams): each tail polls for the other's push."""
    group = None

    @staticmethod
    def make_buffers(world, n_floats, device):
        from . import _lib
        nbytes = _lib.load().vqb_exchange_bytes(int(n_floats), world)
        return [torch.zeros((nbytes + 3) // 4, dtype=torch.float32, device=device) for _ in range(world)]

    def __init__(self, bufs, rank):
        self.bufs, self.world, self.rank = bufs, len(bufs), rank
        self.ptrs = torch.tensor([b.data_ptr() for b in bufs], dtype=torch.int64, device=bufs[0].device)

    def peer_ptrs_dev(self, n_flat):
        return self.ptrs.data_ptr()


def enable_fused_allreduce(module, group=None):
    """Sum the quantizer's parameter gradients over `group` INSIDE the backward's tail kernel (one-shot all-reduce over
    NVLink peer memory) instead of an NCCL call after it.  Collective: call on every rank, once, after the process
    group is up and the module is on its GPU.  From then on EVERY gradient the module's autograd functions hand to
    autograd is already summed over the group: the tensor-core forward route by the tail kernel, every other route
    (inference() lookups, scatter-only, loss extensions) by an NCCL all-reduce of its own output inside its backward
    (reduce_route_output).  allreduce_codebook_grads() then only applies `average`.
    All ranks must run the same sequence of quantizer calls with the same set of upstream gradients (as data-parallel
    replicas of one model do); a rank whose shard is empty still takes part (its tail runs over zero rows)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    n = sum(p.numel() for p in module.parameters() if p.requires_grad)
    ex = PeerExchange(group, max_floats=max(n, 1024))
    module.fused_tail.exchange = ex
    return ex


def finish_codebook_grads(module, stream=None):
    """Deferred exchange (module.fused_tail.defer = True): run the gradient exchange of the last backward -- push this rank's
    sums to the peers, poll for every rank's contribution and add them in rank order -- on `stream` (default: the current
    stream; a side stream keeps the remote stores off the stream the next kernels are queued on).  In a trainer
    call it (or allreduce_codebook_grads, which does it) after loss.backward(): the rest of the model's backward has run in
    between, so no rank waits for another.  It must be enqueued before the module's next backward (the module does so
    itself if the caller forgot)."""
    tail = getattr(module, "fused_tail", None)
    if tail is not None:
        tail.finish(stream)


def check_exchange(module):
    """Raise if an in-kernel exchange of this module has timed out waiting for a peer (one host synchronisation; call it
    where the trainer synchronises anyway, e.g. where it reads the loss).  The kernel itself never traps: it records the
    failure in the module's status words and returns, so the CUDA context survives and the error is recoverable."""
    tail = getattr(module, "fused_tail", None)
    if tail is None or tail.counter is None:
        return
    flag = int(tail.counter[2].item())
    if flag:
        raise RuntimeError("semi-tts_b200: the fused gradient exchange timed out waiting for rank %d "
                           "(VQB_EXCHANGE_TIMEOUT_MS, default 120000); the gradients of that step are incomplete" % (flag - 1))


def allreduce_codebook_grads(module, group=None, average=False, include_usage=False):
    """Sum (or average) the quantizer's parameter gradients and its usage histogram across ranks: the only
    exchange of the data-parallel path.  No host synchronisation (CUDA-graph capturable).
    With a fused exchange attached (enable_fused_allreduce) every contribution to p.grad was already summed over the
    group when autograd received it, so only `average` is applied here.  Otherwise ONE NCCL all-reduce: gradients produced
    by this package's backward are views of one flat buffer and are reduced in place; gradients from elsewhere (e.g.
    after accumulation into pre-existing .grad tensors) are packed into a temporary buffer first.  The int64 usage
    histogram is only consumed at plot time (every 500 steps, bin/train_vqvae.py:305), so by default it is exchanged
    there (`allreduce_usage`, or `include_usage=True` to do it in the same call: a second, 8*K-byte all-reduce of the
    counts accumulated since the previous exchange)."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in module.parameters() if p.requires_grad and p.grad is not None]
    tail = getattr(module, "fused_tail", None)
    if tail is not None:
        tail.finish()                      # a deferred exchange completes here at the latest
    if grads and tail is not None and tail.exchange is not None:
        # already summed over the group, route by route
        if average:
            flat = _flat_view(grads)
            for g in ([flat] if flat is not None else grads):
                g /= world
    elif grads:
        flat = _flat_view(grads)
        if flat is not None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                flat /= world
        else:
            packed = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
            off = 0
            for g in grads:
                n = g.numel()
                chunk = packed[off:off + n].view_as(g)
                g.copy_(chunk / world if average else chunk)
                off += n
    if include_usage:
        allreduce_usage(module, group)


def allreduce_usage(module, group=None):
    """Exchange the usage histogram (counts since the previous exchange are summed over ranks exactly once)."""
    usage = getattr(module, "usage", None)
    if usage is not None:
        usage.all_reduce(group)
