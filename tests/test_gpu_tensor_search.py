"""tcgen05 fused-mode search (VQB_SEARCH_TENSOR) against the exact-fp32 SIMT search on the same inputs:
indices must be IDENTICAL (the re-rank evaluates the same fp32 expression in the same fmaf order), the
straight-through output bit-identical, the usage histogram equal to bincount; against the fp64 oracle the
only admissible mismatches are rows whose top-2 distance gap is below 1e-6 relative."""
import numpy as np
import pytest
import torch

from oracle import vq_oracle as O

pytestmark = pytest.mark.gpu


def _case(N, K, D, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(N, D, generator=g) * scale).cuda()
    e = torch.randn(K, D, generator=g).cuda()
    return x, e


@pytest.mark.parametrize("N,K,D", [(1, 1, 32), (200, 43, 64), (128, 128, 64), (1000, 129, 32), (3000, 300, 128),
                                   (4096, 1024, 64), (2500, 4096, 256), (777, 8192, 64),
                                   (1500, 1000, 256), (300, 130, 256),
                                   # several tiles per CTA on the streamed 3xTF32 kernels (x / x_lo slots alternate)
                                   (60000, 256, 64), (45000, 200, 32),
                                   # several tiles per CTA on the 1xTF32 kernels with two epilogue warpgroups (D <= 128, K > 1024):
                                   # the second warpgroup's list is handed over and released once per tile
                                   (40000, 2048, 64), (40000, 1500, 32), (20000, 2048, 128)])
def test_tensor_search_equals_exact_simt_search(N, K, D):
    import semi_tts_b200 as V
    x, e = _case(N, K, D, seed=N + K + D)
    hist_t = torch.zeros(K, dtype=torch.int64, device="cuda")
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    idx_t, q_t = V.vq_search(x, e, hist=hist_t, search_tensor=True, stats=stats)
    idx_s, q_s = V.vq_search(x, e, search_tensor=False)
    torch.cuda.synchronize()
    st = stats.cpu().tolist()
    print("N=%d K=%d D=%d: re-ranked rows %d, full-scan rows %d" % (N, K, D, st[0], st[1]))
    assert torch.equal(idx_t, idx_s)
    assert torch.equal(q_t, q_s)
    assert torch.equal(hist_t.cpu(), torch.bincount(idx_s.cpu(), minlength=K))
    if N * K <= 4096 * 1024:
        d64 = O.l2_distance(x.cpu().numpy(), e.cpu().numpy())
        rep = O.index_mismatch_report(idx_t.cpu().numpy(), np.argmin(d64, -1), d64)
        assert rep["hard_mismatches"] == 0, rep


def test_tensor_search_adversarial_near_ties_fall_back_to_exact_scan():
    """Many codewords at (almost) the same distance overflow the top-4 window -> full exact scan path."""
    import semi_tts_b200 as V
    g = torch.Generator().manual_seed(5)
    base = torch.randn(1, 64, generator=g)
    e = (base + 1e-4 * torch.randn(512, 64, generator=g)).cuda()
    x = (base + 0.5 * torch.randn(300, 64, generator=g)).cuda()
    stats = torch.zeros(2, dtype=torch.int32, device="cuda")
    idx_t, q_t = V.vq_search(x, e, search_tensor=True, stats=stats)
    idx_s, q_s = V.vq_search(x, e, search_tensor=False)
    assert stats.cpu()[1].item() > 0
    assert torch.equal(idx_t, idx_s) and torch.equal(q_t, q_s)


def test_tensor_search_temperature_and_scaled_inputs():
    import semi_tts_b200 as V
    x, e = _case(2048, 600, 64, seed=9, scale=7.5)
    temp = torch.tensor([0.37], device="cuda")
    idx_t, q_t = V.vq_search(x, e, temp=temp, search_tensor=True)
    idx_s, q_s = V.vq_search(x, e, temp=temp, search_tensor=False)
    assert torch.equal(idx_t, idx_s) and torch.equal(q_t, q_s)


def test_module_fused_search_roundtrip_properties_large():
    """Size-independent properties at a C3-like size: quantising codewords returns themselves; histogram sums to N."""
    import semi_tts_b200 as V
    K, D, N = 2048, 64, 1 << 18
    g = torch.Generator().manual_seed(11)
    e = torch.randn(K, D, generator=g).cuda()
    pick = torch.randint(0, K, (N,), generator=g).cuda()
    hist = torch.zeros(K, dtype=torch.int64, device="cuda")
    idx, q = V.vq_search(e[pick], e, hist=hist, search_tensor=True)
    assert torch.equal(idx, pick)                              # codewords are their own nearest neighbour
    assert torch.equal(q, (e[pick] + e[pick]) - e[pick])
    assert int(hist.sum().item()) == N and torch.equal(hist.cpu(), torch.bincount(pick.cpu(), minlength=K))
