"""Run-length collapse of the quantizer's output on the GPU: drop-in for `VQVAE.mean_forward`
(reference: src/vqvae.py:218-257), the step that follows the bottleneck on the unpaired-speech branch (:128).

The reference moves every row of indices to the host and walks it in Python; here one kernel plans all utterances
(segment boundaries, blank removal, output slots), a second one writes the padded means, and the host reads back the B
segment counts once (the output's padded length depends on them).  Differentiable w.r.t. `latent`.
"""
import ctypes

import torch

from . import _lib
from ._lib import ptr
from .functional import _require, _stream, _c, _g32


def row_argmax(p_code):
    """p_code.argmax(-1) with first-index tie-breaking, on the GPU (src/vqvae.py:223)."""
    _require(p_code, "p_code")
    lib = _lib.load()
    p2 = _c(p_code.detach()).view(-1, p_code.shape[-1])
    idx = torch.empty(p2.shape[0], device=p2.device, dtype=torch.int64)
    with torch.cuda.device(p2.device):
        _lib.check(lib.vqb_row_argmax(ptr(p2), p2.shape[0], p2.shape[1], ptr(idx), _stream(p2)))
    return idx.view(p_code.shape[:-1])


class _SegmentMean(torch.autograd.Function):
    @staticmethod
    def forward(ctx, latent, idx, max_frames_per_phn):
        _require(latent, "latent"); _require(idx, "idx", torch.int64)
        lib = _lib.load()
        B, T, D = latent.shape
        lat = _c(latent.detach())
        idx = _c(idx)
        dev = lat.device
        plan = torch.empty(3, B, T, device=dev, dtype=torch.int32)          # slot_of_row | seg_start | seg_count
        lens = torch.empty(B, device=dev, dtype=torch.int64)
        with torch.cuda.device(dev):
            _lib.check(lib.vqb_segment_plan(ptr(idx), B, T, int(max_frames_per_phn), ptr(plan[0]), ptr(plan[1]),
                                            ptr(plan[2]), ptr(lens), _stream(lat)))
        lens_host = lens.cpu()                                              # the one device->host read of this step
        lmin, lmax = (int(lens_host.min()), int(lens_host.max())) if B else (0, 0)
        ctx.mark_non_differentiable(lens)
        ctx.set_materialize_grads(False)
        if lmin == 0:                                                       # an all-blank sample: the caller returns None (:247-248)
            ctx.empty = True
            return torch.zeros(B, 0, D, device=dev), lens
        out = torch.empty(B, lmax, D, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(lib.vqb_segment_mean(ptr(lat), ptr(plan[1]), ptr(plan[2]), ptr(lens), B, T, D, lmax, ptr(out),
                                            _stream(lat)))
        ctx.empty = False
        ctx.dims = (B, T, D, lmax)
        ctx.save_for_backward(plan)
        return out, lens

    @staticmethod
    def backward(ctx, g_out, _g_lens):
        if g_out is None or ctx.empty:
            return None, None, None
        (plan,) = ctx.saved_tensors
        B, T, D, lmax = ctx.dims
        lib = _lib.load()
        g = _g32(g_out)
        dlat = torch.empty(B, T, D, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            _lib.check(lib.vqb_segment_mean_backward(ptr(g), ptr(plan[0]), ptr(plan[2]), B, T, D, lmax, ptr(dlat), _stream(g)))
        return dlat, None, None


def mean_forward(p_code, latent, max_frames_per_phn, idx=None):
    """(batch_latent[B,Lmax,D], trimmed_len[B] int64) or None if any utterance is all blank -- the return contract of
    VQVAE.mean_forward (src/vqvae.py:247-257).  `idx` may pass the indices the quantizer already produced
    (module.last_idx rows of the same utterances) to skip the argmax over p_code."""
    if idx is None:
        idx = row_argmax(p_code)
    out, lens = _SegmentMean.apply(latent, idx, max_frames_per_phn)
    if out.shape[1] == 0:
        return None
    return out, lens


def vqvae_mean_forward(self, p_code, latent):
    """Method form, installed over src.vqvae.VQVAE.mean_forward by patch.install_into_reference()."""
    return mean_forward(p_code, latent, self.max_frames_per_phn)


class _CtcLogp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p_code, eps):
        _require(p_code, "p_code")
        lib = _lib.load()
        B, S, K = p_code.shape
        p = _c(p_code.detach())
        out = torch.empty(S, B, K, device=p.device, dtype=torch.float32)
        with torch.cuda.device(p.device):
            _lib.check(lib.vqb_ctc_logp(ptr(p), B, S, K, float(eps), ptr(out), _stream(p)))
        ctx.eps = float(eps)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(p)
        return out

    @staticmethod
    def backward(ctx, g):
        if g is None:
            return None, None
        (p,) = ctx.saved_tensors
        B, S, K = p.shape
        lib = _lib.load()
        g = _g32(g)
        gp = torch.empty(B, S, K, device=p.device, dtype=torch.float32)
        with torch.cuda.device(p.device):
            _lib.check(lib.vqb_ctc_logp_backward(ptr(g), ptr(p), B, S, K, ctx.eps, ptr(gp), 0, _stream(p)))
        return gp, None


def ctc_log_probs(p_code, eps=1e-10):
    """`(p_code + EPS).transpose(0, 1).log()` of bin/train_vqvae.py:430-432 / :236 in one pass: p_code[B,S,K] ->
    contiguous [S,B,K] log-probabilities for nn.CTCLoss, differentiable w.r.t. p_code."""
    return _CtcLogp.apply(p_code, eps)
