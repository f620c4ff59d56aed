#!/usr/bin/env python
"""BASELINE config 4: a full bin/train_vqvae.py training step of the REFERENCE's own VQVAE (ASR encoder -> quantizer ->
Tacotron-2, CTC + spectrogram losses, backward, clip, Adam) on synthetic LJSpeech-shaped batches, stock vs. with the
B200 quantizer dropped in (semi_tts_b200.install_into_reference), on one GPU or with the batch sharded over N GPUs.

    python tools/train_step_c4.py                       # 1 GPU: parity at the quantizer boundary + step times
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/train_step_c4.py     # sharded

The reference modules come from baseline/_ref (baseline/install_reference.py: a verbatim, git-ignored copy that travels to
the GPU box) or from /root/reference in the build container.  The step is a restatement of ONE iteration of
VqvaeTrainer.exec (bin/train_vqvae.py:124-270) on the reference's own model, loss and optimizer objects -- the trainer
class itself cannot run without the audio corpus (src/data.py) -- each block citing the lines it follows.

What is compared (hard part 8): everything outside the quantizer is not shard-invariant (BatchNorm
statistics, dropout streams), so parity is judged AT THE QUANTIZER BOUNDARY: same enc_embs in => same p_code / indices /
new_latent out, and same dx / codebook gradients for the same upstream gradients -- for the stock vs drop-in model from
the same seed, and for the sharded vs unsharded quantizer on the captured boundary tensors.
Prints one JSON line (rank 0).
"""
import copy
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np     # noqa: E402
import torch           # noqa: E402
import yaml            # noqa: E402

EPS = 1e-10            # bin/train_vqvae.py:18
N_MELS, LINEAR_DIM, VOCAB, N_SPKR = 80, 1025, 43, 109      # src/audio feat_dim, src/text.py:91-93, corpus/spkr/lj_vctk.json
SAMPLE_RATE = 22050


def load_reference():
    from oracle import ref_import
    if not ref_import.available():
        raise RuntimeError("no reference tree: run `python baseline/install_reference.py` in the build container")
    ref_import.import_reference()
    V_ref = ref_import.import_reference_vqvae()
    import src.util as ref_util
    import src.optim as ref_optim
    with open(os.path.join(ref_import.REFERENCE_ROOT, "config", "semi-multi-spkr-paired-data.yaml")) as f:
        cfg = yaml.load(f, Loader=yaml.FullLoader)
    return ref_import, V_ref, ref_util, ref_optim, cfg


def build_model(V_ref, ref_import, cfg, device, dropin, seed=0):
    """VqvaeTrainer.set_model (bin/train_vqvae.py:77): VQVAE(n_mels, linear_dim, vocab_size, n_spkr, **config['model'])."""
    import semi_tts_b200 as V
    with ref_import.reference_cwd():                     # phn_attr_pth: 'data/phn_attr.csv' is relative to the reference root
        if dropin:
            V.install_into_reference()
        try:
            torch.manual_seed(seed)
            np.random.seed(seed)
            model = V_ref.VQVAE(N_MELS, LINEAR_DIM, VOCAB, N_SPKR, **copy.deepcopy(cfg["model"]))
        finally:
            if dropin:
                V.uninstall_from_reference()
    model = model.to(device)
    if dropin:
        # VQVAE.mean_forward was bound at construction time only through the class: keep the GPU version on this instance
        from semi_tts_b200.segment import vqvae_mean_forward
        model.mean_forward = vqvae_mean_forward.__get__(model, type(model))
    return model


def synth_batch(B, seed, device, t_lo=300, t_hi=800, r=3):
    """What VqvaeTrainer.fetch_data returns (bin/train_vqvae.py:34-52) for LJSpeech-shaped utterances: mel / aug_mel in [0, 1]
    (src/audio.py:284-285), zero beyond each utterance's length (src/data.py:134-136), T ~ U[300, 800]; mel and linear carry
    the extra pad of :44-46; text = phoneme ids 3..42 with L ~ T / 6 (FRAME_PHN_RATIO, src/vqvae.py:18), zero padded."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(t_lo, t_hi + 1, (B,), generator=g)
    lens[0] = t_hi
    T = int(lens.max())
    mel = torch.rand(B, T, N_MELS, generator=g)
    linear = torch.rand(B, T, LINEAR_DIM, generator=g)
    frame = torch.arange(T)[None, :, None]
    keep = (frame < lens[:, None, None]).float()
    mel, linear = mel * keep, linear * keep
    aug_mel = (mel + 0.01 * torch.rand(B, T, N_MELS, generator=g)) * keep
    pad = r - (T % r)                                     # :44 (at least one frame padded)
    mel = torch.cat([mel, torch.zeros(B, pad, N_MELS)], 1)
    linear = torch.cat([linear, torch.zeros(B, pad, LINEAR_DIM)], 1)
    L = max(2, int(T / 6))
    tl = torch.clamp((lens.float() / 6).long(), 2, L)
    text = torch.randint(3, VOCAB, (B, L), generator=g)
    text = text * (torch.arange(L)[None, :] < tl[:, None]).long()
    sid = torch.randint(0, N_SPKR, (B,), generator=g)
    return [t.to(device) for t in (mel, aug_mel, linear, text, sid)]


class Step:
    """One iteration of VqvaeTrainer.exec on a given model; holds the reference's own loss / optimizer objects."""

    def __init__(self, model, ref_util, ref_optim, cfg, dist_group=None):
        from functools import partial
        hp = cfg["hparas"]
        self.model, self.hp = model, hp
        self.freq_loss = partial(ref_util.freq_loss, sample_rate=SAMPLE_RATE, n_mels=N_MELS, loss=hp["freq_loss_type"],
                                 differential_loss=hp["differential_loss"],
                                 emphasize_linear_low=hp["emphasize_linear_low"])              # :81-88
        self.ctc_loss = torch.nn.CTCLoss()                                                     # :89
        self.optimizer = ref_optim.Optimizer(model.parameters(), **hp)                         # :93
        self.group = dist_group
        self.boundary = {}

    def compute_ctcloss(self, model_output, target):
        """bin/train_vqvae.py:430-444 with paras.actual_len False (the default of main.py)."""
        ctc_input = (model_output + EPS).transpose(0, 1).log()
        ctc_target = target.to_sparse().values()
        ctc_len = torch.LongTensor([model_output.shape[1]] * model_output.shape[0]).to(model_output.device)
        return self.ctc_loss(ctc_input, ctc_target, ctc_len, torch.sum(target != 0, dim=-1))

    def _hook_boundary(self):
        """capture enc_embs, the quantizer's outputs and the gradients that cross the boundary"""
        b = self.boundary
        b.clear()
        cb = self.model.codebook

        def pre(_m, args):
            x = args[0]
            b["x"] = x.detach().clone()
            b["first_n"] = args[1] if len(args) > 1 else 0
            if x.requires_grad:
                x.register_hook(lambda g: b.__setitem__("dx", g.detach().clone()) if g is not None else None)

        def post(_m, _args, out):
            b["p_code"], b["new_latent"] = out[0].detach().clone(), out[1].detach().clone()
            if out[0].requires_grad:
                out[0].register_hook(lambda g: b.__setitem__("g_p", g.detach().clone()) if g is not None else None)
            if out[1].requires_grad:
                out[1].register_hook(lambda g: b.__setitem__("g_q", g.detach().clone()) if g is not None else None)

        return [cb.register_forward_pre_hook(pre), cb.register_forward_hook(post)]

    def run(self, pair, unpair, step, capture=False, do_update=True):
        """speech-first iteration (step % 2 == 0, :137-143, :158-185) with unpaired speech, or text-first (:186-205)."""
        model, hp = self.model, self.hp
        mel, aug_mel, linear, text, sid = pair
        use_unpair_speech = hp["unpair_speech_weight"] > 0 and step > hp["unpair_speech_start_step"]      # :130
        use_unpair_text = hp["unpair_text_weight"] > 0 and step > hp["unpair_text_start_step"]            # :129
        tf_rate = self.optimizer.pre_step(step)                                                          # :132 (zero_grad, lr)
        hooks = self._hook_boundary() if capture else []
        total_loss = 0
        speech_first = step % 2 == 0
        unpair_mel = unpair_aug = unpair_linear = unpair_sid = None
        if speech_first and use_unpair_speech and unpair is not None:
            unpair_mel, unpair_aug, unpair_linear, _, unpair_sid = unpair
        if speech_first:
            pair_prob, _, unpair_prob, unpair_latent, unpair_latent_len, _, _ = \
                model.speech_to_text(paired_mel=aug_mel, unpaired_mel=unpair_aug)                           # :159-160
            ignore_speech_cycle = unpair_latent is None
            unpaired_teacher = None if ignore_speech_cycle else unpair_mel                                  # :163-171
            pair_mel_pred, pair_linear_pred, _, _, unpair_mel_pred, unpair_linear_pred, _, _ = \
                model.text_to_speech(paired_text=text, paired_sid=sid, unpaired_sid=unpair_sid,
                                     unpaired_latent=unpair_latent, unpaired_text=None,
                                     unpaired_latent_len=unpair_latent_len, paired_teacher=mel,
                                     unpaired_teacher=unpaired_teacher, tf_rate=tf_rate)                     # :174-185
        else:
            pair_mel_pred, pair_linear_pred, _, _, unpair_mel_pred, unpair_linear_pred, _, _ = \
                model.text_to_speech(paired_text=text, paired_sid=sid, unpaired_sid=None, unpaired_latent=None,
                                     unpaired_text=None, unpaired_latent_len=None, paired_teacher=mel,
                                     unpaired_teacher=None, tf_rate=tf_rate)                                 # :188-199
            pair_prob, _, unpair_prob, unpair_latent, unpair_latent_len, _, _ = \
                model.speech_to_text(paired_mel=aug_mel, unpaired_mel=None, using_fake_mel=use_unpair_text)  # :202-205
            ignore_speech_cycle = True
        asr_loss = self.compute_ctcloss(pair_prob, text)                                                    # :208
        total_loss = total_loss + hp["asr_weight"] * asr_loss                                               # :214
        tts_loss = self.freq_loss(pair_mel_pred, mel) + self.freq_loss(pair_linear_pred, linear)            # :220-222
        total_loss = total_loss + hp["tts_weight"] * tts_loss                                               # :223
        if speech_first and not ignore_speech_cycle:
            unpair_speech_loss = self.freq_loss(unpair_mel_pred, unpair_mel) + \
                self.freq_loss(unpair_linear_pred, unpair_linear)                                           # :229-230
            if step > hp["unpair_speech_start_step"]:
                total_loss = total_loss + hp["unpair_speech_weight"] * unpair_speech_loss                   # :232-233
        # BaseSolver.backward (src/solver.py:138-151)
        try:
            total_loss.backward()
        finally:
            for h in hooks:
                h.remove()
        if self.group is not None:
            self._allreduce_grads()
        grad_norm = torch.nn.utils.clip_grad_norm_(model.parameters(), 5.0)
        if do_update and not torch.isnan(grad_norm):
            self.optimizer.step()
        return total_loss.detach(), grad_norm.detach()

    def _allreduce_grads(self):
        """data-parallel training: the quantizer's gradients by this package (fused NVLink exchange or one NCCL all-reduce of
        the flat buffer), every other parameter by one flat NCCL all-reduce; averaged, as the losses are batch means."""
        import torch.distributed as dist
        import semi_tts_b200 as V
        world = dist.get_world_size(self.group)
        cb = self.model.codebook
        cb_ids = {id(p) for p in cb.parameters()}
        if isinstance(cb, (V.L2Embedding, V.SeperateEmbedding)):
            V.dist.allreduce_codebook_grads(cb, self.group, average=True)
        else:
            cb_ids = set()
        rest = [p.grad for p in self.model.parameters() if p.grad is not None and id(p) not in cb_ids]
        flat = torch.cat([g.reshape(-1) for g in rest])
        dist.all_reduce(flat, group=self.group)
        flat /= world
        off = 0
        for g in rest:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()


def rel(a, b):
    a, b = a.double(), b.double()
    d = b.norm()
    return float((a - b).norm() / d) if d > 0 else float((a - b).norm())


def boundary_parity(stock, drop):
    """same seed, same state => same enc_embs; compare what the two quantizers did with it"""
    out = {"enc_embs_identical": bool(torch.equal(stock["x"], drop["x"]))}
    idx_s, idx_d = stock["p_code"].argmax(-1), drop["p_code"].argmax(-1)
    out["rows"] = int(idx_s.numel())
    out["index_mismatches"] = int((idx_s != idx_d).sum())
    out["p_code_rel_err"] = rel(drop["p_code"], stock["p_code"])
    same = (idx_s == idx_d)
    out["new_latent_bit_identical_where_index_agrees"] = bool(torch.equal(stock["new_latent"][same], drop["new_latent"][same]))
    for k in ("g_p", "g_q", "dx"):
        if k in stock and k in drop:
            out[k + "_rel_diff_stock_vs_dropin"] = rel(drop[k], stock[k])
    return out


def main():
    # libraries (NCCL prints its version banner) must not pollute stdout: ONE JSON line there
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    out = os.fdopen(real_stdout, "w")
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    group = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    import semi_tts_b200 as V
    from oracle import vq_oracle as O
    ref_import, V_ref, ref_util, ref_optim, cfg = load_reference()
    B = int(os.environ.get("C4_BATCH", cfg["data"]["corpus"]["batch_size"]))       # 8 in the shipped config
    n_timed = int(os.environ.get("C4_STEPS", 6))
    res = {"config": "BASELINE configs[3]: bin/train_vqvae.py step, config/semi-multi-spkr-paired-data.yaml model, synthetic "
                     "LJSpeech-shaped batches (T ~ U[300,800] mel frames)", "world": world, "batch_per_rank": B}

    # ---------------- 1 GPU: stock vs drop-in from the same seed, compared at the quantizer boundary -------------------
    if rank == 0:
        pair, unpair = synth_batch(B, 1, dev), synth_batch(B, 2, dev)
        stock = Step(build_model(V_ref, ref_import, cfg, dev, dropin=False), ref_util, ref_optim, cfg)
        drop = Step(build_model(V_ref, ref_import, cfg, dev, dropin=True), ref_util, ref_optim, cfg)
        assert type(drop.model.codebook) is V.L2Embedding and type(stock.model.codebook) is not V.L2Embedding
        drop.model.load_state_dict(stock.model.state_dict(), strict=True)                   # bin/train_vqvae.py:106
        par = {}
        for name, step_idx in (("speech_first", 2), ("text_first", 3)):
            losses = []
            for st in (stock, drop):
                torch.manual_seed(1234); np.random.seed(1234)                              # same dropout / skip streams
                loss, gn = st.run(pair, unpair, step_idx, capture=True, do_update=False)
                losses.append((float(loss), float(gn)))
            p = boundary_parity(stock.boundary, drop.boundary)
            p["loss_stock"], p["loss_dropin"] = losses[0][0], losses[1][0]
            p["grad_norm_stock"], p["grad_norm_dropin"] = losses[0][1], losses[1][1]
            # the drop-in's backward against the fp64 oracle on ITS OWN captured boundary tensors
            b = drop.boundary
            cbm = drop.model.codebook
            E = O.assemble_table(cbm.learnable_table.detach().cpu().numpy(), cbm.phn_attr.weight.cpu().numpy(),
                                 cbm.proj_attr.weight.detach().cpu().numpy(), cbm.proj_attr.bias.detach().cpu().numpy())
            x = b["x"].cpu().numpy()
            Bx, Sx, D = x.shape
            f = O.l2_forward(x, E, 1.0)
            g_p = b["g_p"].cpu().numpy() if "g_p" in b else None
            g_q = b["g_q"].cpu().numpy() if "g_q" in b else None
            n_real = int(b["first_n"]) * Sx if b["first_n"] else 0
            ob = O.l2_backward(x, E, 1.0, f["p_code"], b["p_code"].argmax(-1).cpu().numpy(), g_p, g_q,
                               first_n_real_rows=n_real)
            p["p_code_rel_err_vs_fp64_oracle"] = rel(b["p_code"].cpu(), torch.from_numpy(f["p_code"]))
            if "dx" in b:
                p["dx_rel_err_vs_fp64_oracle"] = rel(b["dx"].cpu(), torch.from_numpy(ob["dx"]))
            par[name] = p
            if name == "speech_first":                          # both upstream gradients cross the boundary in this step
                boundary = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in drop.boundary.items()}
        res["boundary_parity"] = par
        # codebook gradients of the whole step, stock vs drop-in (they include the inference() route of text_to_speech)
        g_s = stock.model.codebook.learnable_table.grad
        g_d = drop.model.codebook.learnable_table.grad
        res["learnable_table_grad_rel_diff_text_first_step"] = rel(g_d, g_s)

        # ---------------- step time, stock vs drop-in (optimizer updates on, eager, as the trainer runs) ----------------
        def timed(st):
            for i in range(2):
                st.run(pair, unpair, 2 + i)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for i in range(n_timed):
                st.run(pair, unpair, 4 + i)
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / n_timed * 1e3
        res["ms_per_step_stock_1gpu"] = timed(stock)
        res["ms_per_step_dropin_1gpu"] = timed(drop)
        # the quantizer's share: its own calls inside one speech-first step, timed with events around the module
        cb = drop.model.codebook
        ev = []
        h1 = cb.register_forward_pre_hook(lambda m, a: ev.append(_rec()))
        h2 = cb.register_forward_hook(lambda m, a, o: ev.append(_rec()))
        drop.run(pair, unpair, 20, do_update=False)
        torch.cuda.synchronize()
        h1.remove(); h2.remove()
        res["quantizer_forward_ms_in_step_dropin"] = ev[0].elapsed_time(ev[1])
        ev.clear()
        cbs = stock.model.codebook
        h1 = cbs.register_forward_pre_hook(lambda m, a: ev.append(_rec()))
        h2 = cbs.register_forward_hook(lambda m, a, o: ev.append(_rec()))
        stock.run(pair, unpair, 20, do_update=False)
        torch.cuda.synchronize()
        h1.remove(); h2.remove()
        res["quantizer_forward_ms_in_step_stock"] = ev[0].elapsed_time(ev[1])
        del stock
    # ---------------- N GPUs: the batch sharded by utterance, codebook replicated ----------------------------------------
    if world > 1:
        import torch.distributed as dist
        model = build_model(V_ref, ref_import, cfg, dev, dropin=True)
        V.dist.enable_fused_allreduce(model.codebook)
        st = Step(model, ref_util, ref_optim, cfg, dist_group=group)
        pair, unpair = synth_batch(B, 100 + rank, dev), synth_batch(B, 200 + rank, dev)      # weak scaling: B utterances per rank
        for i in range(2):
            st.run(pair, unpair, 2 + i)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for i in range(n_timed):
            st.run(pair, unpair, 4 + i)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / n_timed * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["ms_per_step_dropin_sharded"] = float(t.item())
        V.dist.check_exchange(model.codebook)
        # replicas stay bit-identical: same codebook on every rank after the updates
        w = model.codebook.learnable_table.detach().clone()
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        res["codebook_bit_identical_across_ranks_after_updates"] = all(torch.equal(ws[0], z) for z in ws)
        # sharded vs unsharded quantizer on rank 0's captured boundary tensors (same enc_embs and upstream gradients)
        names = ["x", "g_p", "g_q"]
        shapes = [None] * 3
        if rank == 0:
            shapes = [tuple(boundary[k].shape) for k in names]
        dist.broadcast_object_list(shapes, 0)
        tens = []
        for k, shp in zip(names, shapes):
            tt = boundary[k].contiguous() if rank == 0 else torch.empty(shp, device=dev)
            dist.broadcast(tt, 0)
            tens.append(tt)
        x, g_p, g_q = tens
        cbm = model.codebook
        sd = [cbm.learnable_table.detach().clone(), cbm.proj_attr.weight.detach().clone(), cbm.proj_attr.bias.detach().clone()]
        for t_ in sd:
            dist.broadcast(t_, 0)

        def cb_grads(xs, gps, gqs, reduce):
            for p_ in cbm.parameters():
                p_.grad = None
            if xs.shape[0] == 0:
                xs = xs.clone()
            xs = xs.detach().requires_grad_(True)
            p, q, _, _ = cbm(xs)
            torch.autograd.backward([p, q], [gps, gqs])
            if reduce:
                V.dist.allreduce_codebook_grads(cbm, group)
            return torch.cat([p_.grad.reshape(-1) for p_ in cbm.parameters() if p_.requires_grad]).clone()
        ex = cbm.fused_tail.exchange
        cbm.fused_tail.exchange = None
        full = cb_grads(x, g_p, g_q, reduce=False)                      # unsharded, on every rank
        cbm.fused_tail.exchange = ex
        lo, hi = V.dist.shard_bounds(x.shape[0], rank, world)
        part = cb_grads(x[lo:hi], g_p[lo:hi], g_q[lo:hi], reduce=True)  # this rank's utterances, summed over ranks
        res["sharded_vs_unsharded_codebook_grad_rel_err"] = rel(part, full)
        V.dist.check_exchange(cbm)
    if rank == 0:
        print(json.dumps(res), file=out, flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.barrier(); torch.cuda.synchronize()
        out.flush()
        os._exit(0)


def _rec():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


if __name__ == "__main__":
    main()
