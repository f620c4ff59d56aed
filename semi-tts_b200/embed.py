"""Drop-in quantizer modules: same names, constructor, methods, attributes and state-dict keys as the
reference's src/embed.py, with forward/backward running in libvqb200.so.

Interface contract (reference file:line):
  ctor        Cls(vocab_size, ema, softmax, latent_dim, commit_weight, vq_weight, temp, skip_prob,
                  stop_grad, phn_attr_pth=None, proj_attr=None)                 src/embed.py:59-60, :152-153
              called as Cls(vocab_size, False, **codebook)                      src/vqvae.py:57,59; src/tts.py:66
  forward     (enc_embs[B,S,D], first_n_real_mel=0) -> (p_code, new_latent, vq_loss, commit_loss)
                                                                                src/embed.py:105-147, :187-205
  inference   (txt[B,L] int64) -> [B,L,D]                                       src/embed.py:96-103, :180-185
  attributes  out_dim, latent_dim, vocab_size, temp, embedding(.weight), create_msg(),
              load_pretrained_embedding()                                       src/embed.py:13-55, :87-94
  state dict  L2: learnable_table, temp, onehot.weight, phn_attr.weight, proj_attr.{weight,bias}
              separate: temp, onehot.weight, asr_final_layer.*, phn_attr.weight, proj_attr.*, embedding.weight
Extensions (not in the reference): non-zero commit_weight / vq_weight return loss tensors in slots 3/4
(the reference asserts they are 0 and returns the literal 0, 0 -- that behaviour is kept for weight 0);
`usage` accumulates the per-code histogram on the device; `last_idx` exposes the picked indices.
"""
import csv

import numpy as np
import torch
import torch.nn as nn

from . import functional as VF
from .usage import UsageHistogram

N_SPECIAL_TOKENS = 3      # <pad>, <space>, <eos> precede the phones (src/util.py:15)


def read_phn_attr(path, neg_val=0):
    """[3 + n_phones, n_attr] table: TSV with a header row and a phone-name column, three zero rows
    prepended for the special tokens, zeros replaced by neg_val (behaviour of src/util.py:240-245)."""
    rows = []
    with open(path, newline="") as f:
        reader = csv.reader(f, delimiter="\t")
        next(reader)                                   # header
        for rec in reader:
            if rec:
                rows.append([float(v) for v in rec[1:]])
    attr = np.asarray(rows, dtype=np.float64)
    attr[attr == 0] = neg_val
    return np.concatenate([np.zeros((N_SPECIAL_TOKENS, attr.shape[1])), attr])


class _QuantizerBase(nn.Module):
    """Attributes and helpers shared by both variants (reference: BaseEmbedding, src/embed.py:9-55)."""

    def __init__(self, vocab_size, softmax, latent_dim, commit_weight, vq_weight, temp):
        super().__init__()
        self.vocab_size = vocab_size
        self.softmax = softmax
        self.latent_dim = latent_dim
        self.out_dim = latent_dim
        self.commit_weight = commit_weight
        self.vq_weight = vq_weight
        self.ema = False
        self.phn_attr = None
        self.proj_attr = None
        self.use_phn_attr = False
        # The reference draws (and discards) a [K, D] normal table here before anything else
        # (src/embed.py:28, :62); drawing it keeps same-seed construction identical.
        torch.empty(vocab_size, latent_dim).normal_()
        self.onehot = nn.Embedding.from_pretrained(torch.eye(vocab_size), freeze=True)
        if temp < 0:
            self.temp = nn.Parameter(torch.FloatTensor([1]))
        else:
            self.register_buffer("temp", torch.FloatTensor([temp]))
        # device-side usage histogram (replaces the host list of bin/train_vqvae.py:256-261)
        self.usage = UsageHistogram(vocab_size)
        self.track_usage = True
        # forward on the tcgen05 kernel where the shape allows it (False: exact-fp32 CUDA-core kernel)
        self.tensor_cores = True
        self.last_idx = None
        # fused backward tail (partial sums + table backward [+ cross-GPU sum] in one kernel); see functional.FusedTail
        self.fused_tail = VF.FusedTail()
        # no-grad fast path (validation / encode loops): cached table + operand image, direct C-ABI call
        self._nograd = VF.NoGradCache()
        # extension: set to EPS of bin/train_vqvae.py:18 and every grad-mode forward also leaves
        # `self.ctc_logp` = log(p_code + EPS) as contiguous [S, B, K], the tensor compute_ctc_loss builds at :430-432 with a
        # transpose + add + log -- written by the forward kernel's epilogue, its gradient folded into the backward kernel
        self.ctc_eps = None
        self._ctc_out = _Transient()

    @property
    def ctc_logp(self):
        """log(p_code + ctc_eps) [S, B, K] of the last grad-mode forward (None unless `ctc_eps` is set)"""
        return self._ctc_out.value

    def _init_attr(self, latent_dim, phn_attr_pth, proj_attr):
        self.use_phn_attr = phn_attr_pth is not None and phn_attr_pth != ""
        if self.use_phn_attr:
            assert latent_dim > proj_attr > 0, "Currently, proj attr is necessary"
            table = torch.FloatTensor(read_phn_attr(phn_attr_pth))
            self.phn_attr = nn.Embedding.from_pretrained(table, freeze=True, padding_idx=0)
            self.proj_attr = nn.Linear(table.shape[1], proj_attr)
            return proj_attr
        return 0

    def _attr_params(self):
        if self.use_phn_attr:
            return self.phn_attr.weight, self.proj_attr.weight, self.proj_attr.bias
        return None, None, None

    def _hist(self, ref):
        return self.usage.buffer_for(ref) if self.track_usage else None

    def _nograd_params(self):
        """(learnable_table, phn_attr.weight, proj_attr.weight, proj_attr.bias, temp) for the no-grad fast path, looked up
        once: nn.Module.__getattr__ costs more than the rest of the call's bookkeeping.  Rebuilt when the table Parameter
        or temp object is replaced."""
        c = self.__dict__.get("_np")
        lt = self._parameters.get("learnable_table")
        tp = self._parameters.get("temp")
        if tp is None:
            tp = self._buffers.get("temp")
        if c is None or c[0] is not lt or c[4] is not tp:
            c = (lt,) + tuple(self._attr_params()) + (tp,)
            self.__dict__["_np"] = c
        return c

    def create_msg(self):
        return "           | EMA update = {}\t | Temp. = {}\t| Phn. attributs = {} ( projected = {})".format(
            self.ema, "learnable" if type(self.temp) is nn.Parameter else self.temp.data.item(),
            self.use_phn_attr, self.proj_attr is not None)

    def load_pretrained_embedding(self, old_emb):
        """Same key handling as src/embed.py:41-48."""
        if "emb.embedding.weight" in old_emb.keys():
            self.embedding = nn.Embedding.from_pretrained(old_emb["emb.embedding.weight"].data, freeze=False)
            if "emb.temp" in old_emb.keys():
                self.temp.data = old_emb["emb.temp"].data
            if "emb.running_tok_freq" in old_emb.keys():
                self.running_tok_freq = old_emb["emb.running_tok_freq"]
            if "emb.running_ema" in old_emb.keys():
                self.terunning_emamp = old_emb["emb.running_ema"]
        else:
            self.embedding = nn.Embedding.from_pretrained(old_emb["emb.weight"], freeze=False)

    def _losses(self, vq, commit):
        # reference behaviour for zero weights: the integer literal 0 in both slots (src/embed.py:147)
        return (vq if self.vq_weight > 0 else 0), (commit if self.commit_weight > 0 else 0)


class _Transient:
    """holder of a per-step graph tensor that must not travel with copy.deepcopy / pickle of the module"""
    __slots__ = ("value",)

    def __init__(self):
        self.value = None

    def __deepcopy__(self, memo):
        return _Transient()

    def __reduce__(self):
        return (_Transient, ())


class L2Embedding(_QuantizerBase):
    """Nearest-codeword quantizer with an L2 score (reference: src/embed.py:57-147)."""

    def __init__(self, vocab_size, ema, softmax, latent_dim, commit_weight, vq_weight, temp,
                 skip_prob, stop_grad, phn_attr_pth=None, proj_attr=None):
        super().__init__(vocab_size, softmax, latent_dim, commit_weight, vq_weight, temp)
        assert self.softmax == "normal"
        assert not ema
        assert commit_weight >= 0 and vq_weight >= 0       # the reference requires == 0 (src/embed.py:65-66)
        self.skip_prob = skip_prob
        self.stop_grad = stop_grad
        d_attr = self._init_attr(latent_dim, phn_attr_pth, proj_attr)
        self.learnable_table = nn.Parameter(torch.randn((vocab_size, latent_dim - d_attr)))
        # large-codebook option: do not materialise p_code (slot 1 of the return tuple is None)
        self.fused_search = False

    @property
    def embedding(self):
        table, _, _ = VF.assemble_table(self.learnable_table, *self._attr_params())
        return nn.Embedding.from_pretrained(table)

    def inference(self, txt):
        if not torch.is_grad_enabled():
            lt, attr, pw, pb, _ = self._nograd_params()
            return VF.lookup_nograd(self._nograd, txt, lt, attr, pw, pb)
        return VF.codebook_lookup(txt, self.learnable_table, *self._attr_params(), tail=self.fused_tail)

    def forward(self, enc_embs, first_n_real_mel=0, lengths=None):
        """`lengths` (extension; None = the reference's behaviour): valid encoder frames per utterance of
        the zero-padded batch (src/vqvae.py:106-126,259-271).  Frames beyond them are PAD rows: they are not searched, their
        p_code / new_latent rows are zero, they take no part in the histogram or in any gradient, and tiles that hold
        nothing else are skipped; valid rows are bit-identical to the dense call.  (With `--actual_len` the reference's CTC
        already ignores those frames, bin/train_vqvae.py:436-439; the unpaired branch keeps using every frame.)"""
        # numpy's global RNG is consumed exactly when the reference consumes it (src/embed.py:140)
        skip = bool(self.training and self.skip_prob > 0 and np.random.rand() < self.skip_prob)
        want_losses = self.vq_weight > 0 or self.commit_weight > 0
        if not torch.is_grad_enabled() and not want_losses:
            # bin/train_vqvae.py:343 (validate) and the encode path: nothing to differentiate, nothing to save
            lt, attr, pw, pb, temp = self._nograd_params()
            p_code, new_latent, idx = VF.forward_nograd(self._nograd, enc_embs, lt, attr, pw, pb, temp, skip,
                                                        not self.fused_search, self._hist(enc_embs), self.tensor_cores, lengths)
            self.__dict__["last_idx"] = idx              # (plain attribute: skips nn.Module.__setattr__)
            self._ctc_out.value = None
            return p_code, new_latent, 0, 0
        B, S, _ = enc_embs.shape
        attr, pw, pb = self._attr_params()
        ctc_eps = self.ctc_eps if not self.fused_search else None
        out = VF.vq_l2(
            enc_embs, self.learnable_table, attr, pw, pb, self.temp, stop_grad=self.stop_grad, skip=skip,
            n_real_rows=first_n_real_mel * S if first_n_real_mel > 0 else 0,
            want_pcode=not self.fused_search, hist=self._hist(enc_embs), want_losses=want_losses,
            tensor_cores=self.tensor_cores, tail=self.fused_tail, lengths=lengths, ctc_eps=ctc_eps)
        p_code, new_latent, idx, vq, commit = out[:5]
        self._ctc_out.value = out[5] if ctc_eps is not None else None
        self.last_idx = idx
        return (p_code, new_latent) + self._losses(vq, commit)


class SeperateEmbedding(_QuantizerBase):
    """Linear-score quantizer with separate ASR / TTS tables (reference: src/embed.py:150-205)."""

    def __init__(self, vocab_size, ema, softmax, latent_dim, commit_weight, vq_weight, temp,
                 skip_prob, stop_grad, phn_attr_pth=None, proj_attr=None):
        super().__init__(vocab_size, softmax, latent_dim, commit_weight, vq_weight, temp)
        assert self.softmax == "normal"
        assert not ema
        assert commit_weight == 0
        assert vq_weight == 0
        assert skip_prob == 0
        self.stop_grad = stop_grad
        self.asr_final_layer = nn.Linear(latent_dim, vocab_size)
        d_attr = self._init_attr(latent_dim, phn_attr_pth, proj_attr)
        self.embedding = nn.Embedding(vocab_size, latent_dim - d_attr)

    def inference(self, txt):
        return VF.codebook_lookup(txt, self.embedding.weight, *self._attr_params(), tail=self.fused_tail)

    def forward(self, enc_embs, first_n_real_mel=0):
        # first_n_real_mel is unused here, as in the reference (src/embed.py:188)
        attr, pw, pb = self._attr_params()
        ctc_eps = self.ctc_eps if torch.is_grad_enabled() else None
        out = VF.vq_linear(
            enc_embs, self.asr_final_layer.weight, self.asr_final_layer.bias, self.embedding.weight,
            attr, pw, pb, stop_grad=self.stop_grad, hist=self._hist(enc_embs), tensor_cores=self.tensor_cores,
            tail=self.fused_tail, ctc_eps=ctc_eps)
        p_code, new_latent, idx = out[:3]
        self._ctc_out.value = out[3] if ctc_eps is not None else None
        self.last_idx = idx
        return p_code, new_latent, 0, 0
