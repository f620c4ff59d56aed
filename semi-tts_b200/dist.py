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


def pack(tensors):
    """Flatten a list of tensors (None skipped) into one fp32 buffer; returns (flat, metas)."""
    live = [t for t in tensors if t is not None]
    flat = torch.cat([t.reshape(-1).to(torch.float32) for t in live]) if live else torch.zeros(0)
    return flat


def unpack_into(flat, tensors):
    off = 0
    for t in tensors:
        if t is None:
            continue
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
        off += n


def allreduce_codebook_grads(module, group=None, average=False, include_usage=True):
    """Sum (or average) the quantizer's parameter gradients and its usage histogram across ranks with a
    single all-reduce and no host synchronisation (CUDA-graph capturable).  The int64 histogram rides
    along as two fp32 words per code (count >> 12 and count & 0xFFF), each exactly representable and
    exactly summable for counts < 2^36 and world <= 4096."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world = dist.get_world_size(group)
    if world == 1:
        return
    grads = [p.grad for p in module.parameters() if p.requires_grad and p.grad is not None]
    usage = getattr(module, "usage", None)
    counts = usage.counts if (include_usage and usage is not None and usage.counts is not None) else None
    parts = [g.reshape(-1) for g in grads]
    if counts is not None:
        parts += [(counts >> 12).to(torch.float32), (counts & 0xFFF).to(torch.float32)]
    if not parts:
        return
    flat = torch.cat(parts)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        chunk = flat[off:off + n].view_as(g)
        g.copy_(chunk / world if average else chunk)
        off += n
    if counts is not None:
        k = counts.numel()
        hi, lo = flat[off:off + k], flat[off + k:off + 2 * k]
        counts.copy_((hi.to(torch.int64) << 12) + lo.to(torch.int64))
