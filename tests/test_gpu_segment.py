"""GPU parity of the run-length collapse (VQVAE.mean_forward, src/vqvae.py:218-257) against the golden vectors the
unmodified reference produced and against the oracle on larger seeded cases."""
import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu


def _sticky_indices(rng, B, T, K, p_move=0.35, p_blank=0.3):
    idx = np.zeros((B, T), np.int64)
    for b in range(B):
        cur = int(rng.integers(0, K))
        for t in range(T):
            if rng.random() < p_move:
                cur = 0 if rng.random() < p_blank else int(rng.integers(0, K))
            idx[b, t] = cur
    return idx


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_mean_forward_vs_reference_golden(tag):
    import semi_tts_b200 as V
    g = load_golden("mean_forward_" + tag)
    idx, lat, mfp = g["idx"], g["latent"], int(g["max_frames_per_phn"])
    K = int(idx.max()) + 1
    p = torch.nn.functional.one_hot(torch.from_numpy(idx), K).float() * 0.9 + 0.1 / K      # as oracle/gen_golden.py
    lt = torch.from_numpy(lat).cuda().requires_grad_(True)
    res = V.mean_forward(p.cuda(), lt, mfp)
    assert res is not None
    out, lens = res
    assert lens.dtype == torch.int64 and lens.is_cuda                   # trimmed_len: LongTensor on the device (:255)
    assert np.array_equal(lens.cpu().numpy(), g["lens"])
    assert out.shape == g["out"].shape
    assert rel_err(out.detach().cpu().numpy(), g["out"]) < 1e-6
    go = torch.from_numpy(np.random.default_rng(3).standard_normal(out.shape).astype(np.float32)).cuda()
    out.backward(go)
    want = O.mean_forward_backward(idx, go.cpu().numpy().astype(np.float64), mfp)
    assert rel_err(lt.grad.cpu().numpy(), want) < 1e-6


def test_mean_forward_all_blank_sample_returns_none():
    import semi_tts_b200 as V
    idx = torch.zeros(2, 9, dtype=torch.long)
    idx[0, 3] = 1
    p = torch.nn.functional.one_hot(idx, 3).float().cuda()
    assert V.mean_forward(p, torch.randn(2, 9, 4).cuda(), 8) is None


@pytest.mark.parametrize("B,T,D,K,mfp", [(64, 400, 64, 43, 3), (5, 1000, 128, 7, 0), (3, 257, 32, 4, 300), (1, 1, 16, 3, 3),
                                         (7, 256, 64, 43, 1)])
def test_mean_forward_vs_oracle_large(B, T, D, K, mfp):
    """Config-2-shaped and boundary cases (T not a multiple of the scan chunk, runs crossing chunk boundaries,
    max_frames_per_phn = 0 and larger than T), indices passed directly (the quantizer's last_idx)."""
    import semi_tts_b200 as V
    rng = np.random.default_rng(B * 1000 + T)
    idx = _sticky_indices(rng, B, T, K)
    idx[:, 0] = np.maximum(idx[:, 0], 1)                                # no all-blank utterance
    lat = rng.standard_normal((B, T, D)).astype(np.float32)
    want = O.mean_forward(idx, lat.astype(np.float64), mfp)
    lt = torch.from_numpy(lat).cuda().requires_grad_(True)
    out, lens = V.mean_forward(None, lt, mfp, idx=torch.from_numpy(idx).cuda())
    assert np.array_equal(lens.cpu().numpy(), want[1])
    assert out.shape == want[0].shape
    assert rel_err(out.detach().cpu().numpy(), want[0]) < 1e-6
    go = rng.standard_normal(out.shape).astype(np.float32)
    out.backward(torch.from_numpy(go).cuda())
    assert rel_err(lt.grad.cpu().numpy(), O.mean_forward_backward(idx, go.astype(np.float64), mfp)) < 1e-6


def test_row_argmax_first_index_on_ties():
    import semi_tts_b200 as V
    g = torch.Generator().manual_seed(5)
    p = torch.rand(1000, 43, generator=g)
    p[::7, 5] = 2.0
    p[::7, 20] = 2.0                                                    # tied maxima: the first index wins (:223 on CPU)
    got = V.row_argmax(p.cuda().view(10, 100, 43)).cpu()
    assert torch.equal(got.view(-1), p.argmax(-1))


@pytest.mark.parametrize("B,S,K", [(4, 50, 43), (64, 800, 43), (3, 7, 300), (1, 1, 5)])
def test_ctc_log_probs_vs_oracle(B, S, K):
    """ctc_log_probs == (p_code + 1e-10).transpose(0,1).log() (bin/train_vqvae.py:430-432) and its gradient."""
    import semi_tts_b200 as V
    rng = np.random.default_rng(B * S + K)
    logits = rng.standard_normal((B, S, K)) * 6.0                         # peaky rows: some p underflow towards EPS
    p = np.exp(logits - logits.max(-1, keepdims=True)); p /= p.sum(-1, keepdims=True)
    p32 = p.astype(np.float32)
    pt = torch.from_numpy(p32).cuda().requires_grad_(True)
    out = V.ctc_log_probs(pt)
    assert out.shape == (S, B, K) and out.is_contiguous()
    assert rel_err(out.detach().cpu().numpy(), O.ctc_input(p32.astype(np.float64))) < 1e-6
    go = rng.standard_normal((S, B, K)).astype(np.float32)
    out.backward(torch.from_numpy(go).cuda())
    assert rel_err(pt.grad.cpu().numpy(), O.ctc_input_backward(p32.astype(np.float64), go.astype(np.float64))) < 1e-6
