"""BASELINE config 4 on the GPU: the drop-in quantizer INSIDE the reference's own VQVAE (src/vqvae.py), one full training
step of bin/train_vqvae.py (speech-first with unpaired speech, and text-first) stock vs. patched from the same seed and
state, compared at the quantizer boundary (hard part 8).

The reference tree is the installed copy baseline/_ref (baseline/install_reference.py; it travels to the GPU box) or
/root/reference in the build container; the tests skip when neither exists.
"""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import ref_import                      # noqa: E402
from oracle import vq_oracle as O                  # noqa: E402

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_import.available(), reason="no reference tree (baseline/_ref)")]


@pytest.fixture(scope="module")
def c4():
    import train_step_c4 as C
    ref_imp, V_ref, ref_util, ref_optim, cfg = C.load_reference()
    dev = torch.device("cuda", 0)
    stock = C.Step(C.build_model(V_ref, ref_imp, cfg, dev, dropin=False), ref_util, ref_optim, cfg)
    drop = C.Step(C.build_model(V_ref, ref_imp, cfg, dev, dropin=True), ref_util, ref_optim, cfg)
    return C, stock, drop, dev


def test_dropin_is_installed_and_state_compatible(c4):
    import semi_tts_b200 as V
    C, stock, drop, dev = c4
    assert type(drop.model.codebook) is V.L2Embedding
    assert type(stock.model.codebook).__module__ == "src.embed"
    # same seed => same initial weights everywhere; the stock checkpoint loads strictly (bin/train_vqvae.py:106)
    a, b = stock.model.state_dict(), drop.model.state_dict()
    assert list(a.keys()) == list(b.keys())
    assert all(torch.equal(a[k], b[k]) for k in a)
    drop.model.load_state_dict(a, strict=True)


@pytest.mark.parametrize("step_idx,name", [(2, "speech_first"), (3, "text_first")])
def test_full_training_step_parity_at_the_quantizer_boundary(c4, step_idx, name):
    """bin/train_vqvae.py:124-270 on synthetic LJSpeech-shaped batches (shortened: T ~ U[120, 240]); src/vqvae.py:106-141;
    src/solver.py:138-151."""
    C, stock, drop, dev = c4
    drop.model.load_state_dict(stock.model.state_dict(), strict=True)
    pair, unpair = C.synth_batch(4, 1, dev, 120, 240), C.synth_batch(4, 2, dev, 120, 240)
    out = []
    for st in (stock, drop):
        torch.manual_seed(77); np.random.seed(77)
        loss, gn = st.run(pair, unpair, step_idx, capture=True, do_update=False)
        out.append((float(loss), float(gn)))
    s, d = stock.boundary, drop.boundary
    assert torch.equal(s["x"], d["x"])                                        # same encoder output reached both quantizers
    idx_s, idx_d = s["p_code"].argmax(-1), d["p_code"].argmax(-1)
    E = O.assemble_table(*(t.detach().cpu().numpy() for t in (
        drop.model.codebook.learnable_table, drop.model.codebook.phn_attr.weight, drop.model.codebook.proj_attr.weight,
        drop.model.codebook.proj_attr.bias)))
    x = d["x"].cpu().numpy()
    f = O.l2_forward(x, E, 1.0)
    rep = O.index_mismatch_report(idx_d.cpu().numpy().ravel(), idx_s.cpu().numpy().ravel(), f["dist"])
    assert rep["hard_mismatches"] == 0, rep
    same = idx_s == idx_d
    assert torch.equal(s["new_latent"][same], d["new_latent"][same])           # (x + c) - x is deterministic
    assert C.rel(d["p_code"], s["p_code"]) < 2e-5
    assert C.rel(d["p_code"].cpu(), torch.from_numpy(f["p_code"])) < 1e-5
    # backward: the drop-in against the fp64 oracle on the gradients that actually crossed ITS boundary
    g_p = d["g_p"].cpu().numpy() if "g_p" in d else None
    g_q = d["g_q"].cpu().numpy() if "g_q" in d else None
    ob = O.l2_backward(x, E, 1.0, f["p_code"], idx_d.cpu().numpy(), g_p, g_q)
    assert C.rel(d["dx"].cpu(), torch.from_numpy(ob["dx"])) < 1e-5
    # ... and against the stock model's own autograd.  With every index equal the two steps see the same graph, so the
    # losses and the gradients arriving at / leaving the boundary agree to fp32 noise of the downstream network.
    if bool(same.all()):
        assert abs(out[0][0] - out[1][0]) <= 1e-4 * abs(out[0][0])
        for k in ("g_p", "g_q", "dx"):
            if k in s and k in d:
                assert C.rel(d[k], s[k]) < 1e-3, (k, C.rel(d[k], s[k]))
        gs, gd = stock.model.codebook.learnable_table.grad, drop.model.codebook.learnable_table.grad
        assert C.rel(gd, gs) < 1e-3
        gs, gd = stock.model.codebook.proj_attr.weight.grad, drop.model.codebook.proj_attr.weight.grad
        assert C.rel(gd, gs) < 1e-3


def test_optimizer_steps_keep_the_two_models_together(c4):
    """three full steps with updates (alternating speech-first / text-first as the trainer does, :137): the drop-in model's
    codebook tracks the stock model's."""
    C, stock, drop, dev = c4
    drop.model.load_state_dict(stock.model.state_dict(), strict=True)
    pair, unpair = C.synth_batch(4, 3, dev, 120, 200), C.synth_batch(4, 4, dev, 120, 200)
    for st in (stock, drop):
        st.optimizer = type(st.optimizer)(st.model.parameters(), **st.hp)      # fresh Adam state
        for i in range(3):
            torch.manual_seed(100 + i); np.random.seed(100 + i)
            st.run(pair, unpair, 2 + i)
    a, b = stock.model.codebook.learnable_table.detach(), drop.model.codebook.learnable_table.detach()
    assert C.rel(b, a) < 1e-3
