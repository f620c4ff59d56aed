"""Generate tests/golden/*.npz from the UNMODIFIED reference modules -- TEST INFRASTRUCTURE.

Run in the build container only (needs /root/reference):

    python -m oracle.gen_golden            # from the repo root

Each case instantiates the reference quantizer (src/embed.py) exactly as src/vqvae.py:41-59
does from the YAML block, feeds it seeded tensors, runs forward and
torch.autograd.backward([p_code, new_latent], [g_p, g_q]) and stores every input, parameter,
output and gradient.  The committed vectors are what pins oracle/vq_oracle.py,
oracle/torch_port.py and (on the GPU) the CUDA path.
"""
import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _save(name, **arrays):
    arrays = {k: v for k, v in arrays.items() if v is not None}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print("wrote %-28s %s" % (name, {k: getattr(v, "shape", v) for k, v in arrays.items()}))


def _state(mod):
    return {"sd." + k: _np(v) for k, v in mod.state_dict().items()}


def _run_case(name, mod, B, S, seed, first_n_real_mel=0, train=False, want_gp=True, want_gq=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, S, mod.latent_dim, generator=g)
    _run_case_x(name, mod, x, seed, first_n_real_mel, train, want_gp, want_gq, g)


def _run_case_x(name, mod, x, seed, first_n_real_mel=0, train=False, want_gp=True, want_gq=True, g=None):
    """Runs the reference module on a given enc_embs tensor and saves inputs, outputs and gradients."""
    g = g if g is not None else torch.Generator().manual_seed(seed)
    B, S, D = x.shape
    K = mod.vocab_size
    x = x.requires_grad_(True)
    g_p = torch.randn(B, S, K, generator=g) if want_gp else None
    g_q = torch.randn(B, S, D, generator=g) if want_gq else None
    mod.train(train)
    for p in mod.parameters():
        p.grad = None
    p_code, new_latent, vq, commit = mod(x, first_n_real_mel)
    assert vq == 0 and commit == 0
    outs, grads = [], []
    if want_gp:
        outs.append(p_code); grads.append(g_p)
    if want_gq and new_latent.requires_grad:
        outs.append(new_latent); grads.append(g_q)
    torch.autograd.backward(outs, grads)
    arrays = dict(x=_np(x), g_p=_np(g_p), g_q=_np(g_q), p_code=_np(p_code), new_latent=_np(new_latent),
                  idx=_np(p_code.argmax(-1)), dx=_np(x.grad),
                  first_n_real_mel=np.int64(first_n_real_mel), train=np.int64(train))
    for n, p in mod.named_parameters():
        arrays["grad." + n] = _np(p.grad) if p.grad is not None else None
    arrays.update(_state(mod))
    _save(name, **arrays)


def main():
    os.makedirs(OUT, exist_ok=True)
    E = ref_import.import_reference()
    cb_l2 = ref_import.load_codebook_cfg("semi-multi-spkr-paired-data.yaml")
    cb_sep = ref_import.load_codebook_cfg("supervised.yaml")
    assert cb_l2.pop("bone") == "l2" and cb_sep.pop("bone") == "seperate"
    K = 43                                                     # src/text.py:56,91-93

    with ref_import.reference_cwd():
        from src.util import read_phn_attr
        np.save(os.path.join(OUT, "phn_attr_table.npy"),
                read_phn_attr(cb_l2["phn_attr_pth"]).astype(np.float32))

        # ctor RNG-order pin: same seed => same initial parameters (SURVEY hard part 6)
        torch.manual_seed(0)
        m = E.L2Embedding(K, False, **cb_l2)
        _save("init_l2_seed0", **_state(m))
        torch.manual_seed(0)
        m = E.SeperateEmbedding(K, False, **cb_sep)
        _save("init_sep_seed0", **_state(m))

        torch.manual_seed(1)
        _run_case("l2_attr_stopgrad", E.L2Embedding(K, False, **cb_l2), 4, 50, 11)
        _run_case("l2_attr_stopgrad_gq_only", E.L2Embedding(K, False, **cb_l2), 4, 50, 12, want_gp=False)
        _run_case("l2_attr_first_n", E.L2Embedding(K, False, **cb_l2), 4, 50, 13, first_n_real_mel=1)
        _run_case("l2_attr_st_onehot", E.L2Embedding(K, False, **dict(cb_l2, stop_grad=False)), 4, 50, 14)
        _run_case("l2_attr_st_onehot_first_n", E.L2Embedding(K, False, **dict(cb_l2, stop_grad=False)),
                  4, 50, 15, first_n_real_mel=3)
        m = E.L2Embedding(K, False, **dict(cb_l2, temp=-1))
        m.temp.data.fill_(0.6)
        _run_case("l2_attr_learn_temp", m, 4, 50, 16)
        m = E.L2Embedding(K, False, **dict(cb_l2, temp=0.25))
        _run_case("l2_attr_temp_quarter", m, 4, 50, 17)
        m = E.L2Embedding(37, False, **dict(cb_l2, phn_attr_pth=None, proj_attr=None, latent_dim=32))
        _run_case("l2_noattr_k37_d32", m, 3, 40, 18)
        m = E.L2Embedding(300, False, **dict(cb_l2, phn_attr_pth=None, proj_attr=None, latent_dim=128))
        _run_case("l2_noattr_k300_d128", m, 2, 64, 19)
        m = E.L2Embedding(K, False, **dict(cb_l2, skip_prob=1.0))
        _run_case("l2_attr_skip_train", m, 4, 50, 20, train=True)
        _run_case("l2_attr_ragged", E.L2Embedding(K, False, **cb_l2), 1, 1, 21)
        _run_case("l2_attr_ragged_b3_s37", E.L2Embedding(K, False, **cb_l2), 3, 37, 22)

        _run_case("sep_attr_stopgrad", E.SeperateEmbedding(K, False, **cb_sep), 4, 50, 31)
        _run_case("sep_attr_st_onehot", E.SeperateEmbedding(K, False, **dict(cb_sep, stop_grad=False)),
                  4, 50, 32)
        m = E.SeperateEmbedding(29, False, **dict(cb_sep, phn_attr_pth=None, proj_attr=None, latent_dim=48))
        _run_case("sep_noattr_k29_d48", m, 3, 40, 33)
        m = E.SeperateEmbedding(29, False, **dict(cb_sep, phn_attr_pth=None, proj_attr=None, latent_dim=48,
                                                    stop_grad=False))
        _run_case("sep_noattr_st_onehot", m, 3, 40, 34)

        # full BASELINE config-1 size (16 x 200) for index exactness, outputs only
        torch.manual_seed(2)
        m = E.L2Embedding(K, False, **cb_l2)
        _run_case("l2_config1_16x200", m, 16, 200, 40)
        m = E.SeperateEmbedding(K, False, **cb_sep)
        _run_case("sep_config1_16x200", m, 16, 200, 41)

        # inference (gather only), src/embed.py:96-103 / :180-185
        g = torch.Generator().manual_seed(50)
        txt = torch.randint(0, K, (5, 23), generator=g)
        m = E.L2Embedding(K, False, **cb_l2)
        _save("inference_l2", txt=_np(txt), out=_np(m.inference(txt)),
              table=_np(m.embedding.weight), **_state(m))
        m = E.SeperateEmbedding(K, False, **cb_sep)
        _save("inference_sep", txt=_np(txt), out=_np(m.inference(txt)), **_state(m))

        # run-length collapse after the quantizer (src/vqvae.py:218-257), SURVEY 8(f) rank 1
        V = ref_import.import_reference_vqvae()
        for tag, seed, mfp, B_, T_, Kk in (("a", 60, 8, 4, 60, 6), ("b", 61, 3, 3, 41, 4), ("c", 62, 100, 2, 17, 3)):
            g = torch.Generator().manual_seed(seed)
            # sticky random walk over codes so that runs, blanks and long runs all occur
            steps = torch.rand(B_, T_, generator=g) < 0.35
            vals = torch.randint(0, Kk, (B_, T_), generator=g)
            idx = torch.zeros(B_, T_, dtype=torch.long)
            for b in range(B_):
                cur = int(vals[b, 0]) if tag != "c" else 1
                for t in range(T_):
                    if steps[b, t]:
                        cur = int(vals[b, t])
                    idx[b, t] = cur
            if tag == "c":
                idx[0, -1] = 2          # last token non-blank and unique (:243-245)
                idx[1, -3:] = 0         # trailing blanks
            p = torch.nn.functional.one_hot(idx, Kk).float() * 0.9 + 0.1 / Kk
            lat = torch.randn(B_, T_, 16, generator=g)
            fake_self = types.SimpleNamespace(max_frames_per_phn=mfp)
            res = V.VQVAE.mean_forward(fake_self, p, lat)
            assert res is not None
            _save("mean_forward_" + tag, idx=_np(idx), latent=_np(lat), max_frames_per_phn=np.int64(mfp),
                  out=_np(res[0]), lens=_np(res[1]))
        idx = torch.zeros(2, 9, dtype=torch.long); idx[0, 3] = 1
        p = torch.nn.functional.one_hot(idx, 3).float()
        assert V.VQVAE.mean_forward(types.SimpleNamespace(max_frames_per_phn=8), p, torch.randn(2, 9, 4)) is None


if __name__ == "__main__":
    main()
