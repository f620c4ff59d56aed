"""Device-side code-usage histogram.

The reference appends `unpair_prob.argmax(-1).cpu().flatten().tolist()` to a Python list every other
step (bin/train_vqvae.py:256-261) and, every 500 steps, computes `data.count(i)/len(data)` per code with
entry 0 forced to zero (src/util.py:139-143) before resetting the list (:310).  Here the counts are
accumulated by the forward kernel itself into an int64 [K] buffer; `bar()` applies the same formula without a
per-step device->host sync.

WHICH ROWS ARE COUNTED differs from the reference unless the trainer says so: the kernel counts every row of every
forward while `module.track_usage` is True (the default) -- paired and unpaired rows, validation passes included --
whereas the reference counts only the unpaired batch's arg-max, on the speech-first steps (:256-261).  To reproduce
the reference's plot exactly, set `codebook.track_usage = False` and switch it on around the unpaired forward only,
or count rows of one's choosing from `module.last_idx` with `usage.add(idx)`.  (Under the reference's own, unmodified
trainer its host list keeps working as before -- it reads p_code -- and this histogram is an extra.)  Rows masked
by `lengths=` are never counted.

Data-parallel runs: `counts` holds this rank's rows that have not been exchanged yet, `reduced` the part
already summed over all ranks.  `all_reduce()` moves `counts` into `reduced` (sum over ranks) and zeroes it,
so it may be called every step, every N steps or only at plot time -- `bar()` / `total()` always report
`reduced + counts`, and nothing is ever counted twice.
"""
import torch


class UsageHistogram:
    def __init__(self, n_codes):
        self.n_codes = n_codes
        self.counts = None            # int64 [K] on the device of the first forward (this rank, not yet exchanged)
        self.reduced = None           # int64 [K] already summed over the data-parallel group

    def buffer_for(self, ref):
        if self.counts is None or self.counts.device != ref.device:
            self.counts = torch.zeros(self.n_codes, dtype=torch.int64, device=ref.device)
        return self.counts

    def add(self, idx):
        """count the codes of an index tensor (any shape, int64, on the histogram's device) -- for callers that pick the
        rows themselves, e.g. `usage.add(codebook.last_idx[first_n_real_mel:])` for the reference's unpaired rows"""
        flat = idx.reshape(-1)
        self.buffer_for(flat).add_(torch.bincount(flat, minlength=self.n_codes)[:self.n_codes])

    def reset(self):
        if self.counts is not None:
            self.counts.zero_()
        if self.reduced is not None:
            self.reduced.zero_()

    def all_counts(self):
        """int64 [K]: everything counted since the last reset (exchanged part + this rank's pending part)."""
        if self.counts is None:
            return None
        return self.counts if self.reduced is None else self.counts + self.reduced

    def all_reduce(self, group=None):
        """Sum the pending per-rank counts over `group` (no host sync; CUDA-graph capturable)."""
        import torch.distributed as dist
        if self.counts is None or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return
        if self.reduced is None or self.reduced.device != self.counts.device:
            self.reduced = torch.zeros_like(self.counts)
        dist.all_reduce(self.counts, op=dist.ReduceOp.SUM, group=group)
        self.reduced += self.counts
        self.counts.zero_()

    def total(self):
        return 0 if self.counts is None else int(self.all_counts().sum().item())

    def bar(self, zero_pad_tok=True):
        """`cnts` of src/util.py:139-143 as a list of K floats (one host sync, at plot time only)."""
        if self.counts is None:
            return [0.0] * self.n_codes
        c = self.all_counts().to(torch.float64)
        tot = c.sum()
        out = (c / tot) if tot > 0 else c
        if zero_pad_tok:
            out = out.clone()
            out[0] = 0
        return out.cpu().tolist()
