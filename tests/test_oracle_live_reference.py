"""The oracle against the UNMODIFIED reference run live in the build container (skipped where /root/reference is absent,
e.g. on the GPU box): randomised shapes and options beyond the committed golden vectors.  The reference's functions are
called exactly as its own callers do (src/vqvae.py:57-59, :119, :128); the oracle (oracle/vq_oracle.py) must reproduce
them -- fp64 restatement vs the reference's fp32 within the fp32 noise floor, integer results exactly."""
import types

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import ref_import, vq_oracle as O

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference tree not present")


def _sticky(rng, B, T, K, p_move, p_blank):
    idx = np.zeros((B, T), np.int64)
    for b in range(B):
        cur = int(rng.integers(0, K))
        for t in range(T):
            if rng.random() < p_move:
                cur = 0 if rng.random() < p_blank else int(rng.integers(0, K))
            idx[b, t] = cur
    return idx


@pytest.mark.parametrize("seed", range(12))
def test_mean_forward_oracle_vs_live_reference(seed):
    """VQVAE.mean_forward (src/vqvae.py:218-257) on random run-length patterns: runs longer than max_frames_per_phn,
    leading / trailing blanks, single-frame last tokens (:243-245), T = 1, and utterances that are all blank (-> None)."""
    V = ref_import.import_reference_vqvae()
    rng = np.random.default_rng(1000 + seed)
    for _ in range(12):
        B, T, K = int(rng.integers(1, 5)), int(rng.integers(1, 41)), int(rng.integers(2, 7))
        mfp = int(rng.choice([0, 1, 2, 3, 8, 100]))
        idx = _sticky(rng, B, T, K, p_move=float(rng.choice([0.1, 0.35, 0.8])), p_blank=float(rng.choice([0.0, 0.3, 0.7])))
        lat = rng.standard_normal((B, T, 5)).astype(np.float32)
        p = torch.nn.functional.one_hot(torch.from_numpy(idx), K).float() * 0.9 + 0.1 / K
        ref = V.VQVAE.mean_forward(types.SimpleNamespace(max_frames_per_phn=mfp), p, torch.from_numpy(lat))
        got = O.mean_forward(idx, lat.astype(np.float64), mfp)
        if ref is None:
            assert got is None
            continue
        assert got is not None
        assert np.array_equal(got[1], ref[1].numpy())
        assert got[0].shape == tuple(ref[0].shape)
        assert np.allclose(got[0], ref[0].numpy(), rtol=0, atol=2e-6)


@pytest.mark.parametrize("seed", range(10))
def test_l2_quantizer_oracle_vs_live_reference(seed):
    """L2Embedding.forward + autograd (src/embed.py:105-147) at random K, D, temperature, stop_grad / ST-onehot and
    first_n_real_mel, without phoneme attributes (the attribute path is pinned by the golden cases)."""
    ref_embed = ref_import.import_reference()
    rng = np.random.default_rng(2000 + seed)
    K, D = int(rng.integers(2, 90)), int(rng.choice([4, 8, 20, 32, 48]))
    B, S = int(rng.integers(1, 5)), int(rng.integers(1, 30))
    stop_grad = bool(rng.integers(0, 2))
    temp = float(rng.choice([0.25, 1.0, 3.0]))
    # first_n_real_mel == B is not a reference input: the empty fake part fails in neg_batch_l2's reshape (src/embed.py:209);
    # its callers pass len(paired_mel) < batch or 0 (src/vqvae.py:118)
    first_n = int(rng.integers(0, B))
    torch.manual_seed(seed)
    m = ref_embed.L2Embedding(K, False, softmax="normal", latent_dim=D, commit_weight=0, vq_weight=0, temp=temp,
                              skip_prob=0, stop_grad=stop_grad)
    m.eval()
    x = torch.from_numpy(rng.standard_normal((B, S, D)).astype(np.float32)).requires_grad_(True)
    g_p = torch.from_numpy(rng.standard_normal((B, S, K)).astype(np.float32))
    g_q = torch.from_numpy(rng.standard_normal((B, S, D)).astype(np.float32))
    p, q, vq, commit = m(x, first_n)
    assert vq == 0 and commit == 0
    torch.autograd.backward([p, q], [g_p, g_q])
    table = m.learnable_table.detach().numpy().astype(np.float64)
    f = O.l2_forward(x.detach().numpy(), table, temp, stop_grad=stop_grad)
    rep = O.index_mismatch_report(p.argmax(-1).numpy(), f["idx"], f["dist"])
    assert rep["hard_mismatches"] == 0, rep
    assert rel_err(p.detach().numpy(), f["p_code"]) < 5e-6
    if rep["mismatched"] == 0:
        assert rel_err(q.detach().numpy(), f["new_latent"]) < 1e-6
        b = O.l2_backward(x.detach().numpy(), table, temp, f["p_code"], f["idx"], g_p.numpy(), g_q.numpy(), stop_grad=stop_grad,
                          first_n_real_rows=first_n * S)
        assert rel_err(x.grad.numpy(), b["dx"]) < 1e-5
        assert rel_err(m.learnable_table.grad.numpy(), b["dtable"]) < 1e-5


@pytest.mark.parametrize("seed", range(6))
def test_separate_quantizer_oracle_vs_live_reference(seed):
    """SeperateEmbedding.forward + autograd (src/embed.py:187-205) at random K, D, stop_grad / ST-onehot."""
    ref_embed = ref_import.import_reference()
    rng = np.random.default_rng(3000 + seed)
    K, D = int(rng.integers(2, 90)), int(rng.choice([4, 8, 20, 32, 48]))
    B, S = int(rng.integers(1, 5)), int(rng.integers(1, 30))
    stop_grad = bool(rng.integers(0, 2))
    torch.manual_seed(seed)
    m = ref_embed.SeperateEmbedding(K, False, softmax="normal", latent_dim=D, commit_weight=0, vq_weight=0, temp=1,
                                    skip_prob=0, stop_grad=stop_grad)
    x = torch.from_numpy(rng.standard_normal((B, S, D)).astype(np.float32)).requires_grad_(True)
    g_p = torch.from_numpy(rng.standard_normal((B, S, K)).astype(np.float32))
    g_q = torch.from_numpy(rng.standard_normal((B, S, D)).astype(np.float32))
    p, q, _, _ = m(x)
    torch.autograd.backward([p, q], [g_p, g_q])
    w = m.asr_final_layer.weight.detach().numpy().astype(np.float64)
    bias = m.asr_final_layer.bias.detach().numpy().astype(np.float64)
    emb = m.embedding.weight.detach().numpy().astype(np.float64)
    f = O.separate_forward(x.detach().numpy(), emb, w, bias, stop_grad=stop_grad, emb_weight=emb)
    if not np.array_equal(p.argmax(-1).numpy(), f["idx"]):
        pytest.skip("near-tie in the fp32 argmax of this draw")
    assert rel_err(p.detach().numpy(), f["p_code"]) < 5e-6
    assert rel_err(q.detach().numpy(), f["new_latent"]) < 1e-6
    b = O.separate_backward(x.detach().numpy(), emb, w, f["p_code"], f["idx"], g_p.numpy(), g_q.numpy(), stop_grad=stop_grad)
    assert rel_err(x.grad.numpy(), b["dx"]) < 1e-5
    assert rel_err(m.asr_final_layer.weight.grad.numpy(), b["d_asr_w"]) < 1e-5
    assert rel_err(m.asr_final_layer.bias.grad.numpy(), b["d_asr_b"]) < 1e-5
    assert rel_err(m.embedding.weight.grad.numpy(), b["dtable"]) < 1e-5
