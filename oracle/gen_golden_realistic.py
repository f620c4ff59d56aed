"""Extra golden case: the quantizer fed by the reference's OWN speech encoder ("realistic variant").

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden_realistic

mel = rand(B, 2S, 80) (reference mels are clamped to [0, 1], src/audio.py:284-285) goes once through the unmodified
reference CTC encoder (src/asr.py:5-64, built from config/semi-multi-spkr-paired-data.yaml, .eval()) to produce
enc_embs; the unmodified reference L2Embedding then produces p_code / new_latent / gradients exactly as in
oracle/gen_golden.py.  Encoder outputs are far from randn: small norms and neighbouring frames that are almost equal.
TEST INFRASTRUCTURE: runs in the build container only (needs /root/reference); commits tests/golden/l2_realistic_*.npz.
"""
import numpy as np
import torch
import yaml

from oracle import ref_import as R
from oracle.gen_golden import _run_case_x


def main():
    E = R.import_reference()
    import src.asr as ref_asr
    import os
    with R.reference_cwd():
        with open(os.path.join(R.REFERENCE_ROOT, "config", "semi-multi-spkr-paired-data.yaml")) as f:
            cfg = yaml.load(f, Loader=yaml.FullLoader)["model"]
        cb = dict(cfg["codebook"]); cb.pop("bone")
        D = cb["latent_dim"]
        torch.manual_seed(3)
        enc = ref_asr.CTC(80, D, **cfg["encoder"]).eval()
        quant = E.L2Embedding(43, False, **cb)
        g = torch.Generator().manual_seed(70)
        for tag, B, T in (("a", 4, 100), ("b", 8, 256)):
            mel = torch.rand(B, T, 80, generator=g)
            with torch.no_grad():
                x = enc(mel)
            _run_case_x("l2_realistic_" + tag, quant, x.detach().clone(), seed=71 + B, first_n_real_mel=B // 2 if tag == "b" else 0)
            print(tag, "enc_embs", tuple(x.shape), "mean |x| %.4f, row norm %.4f" % (float(x.abs().mean()), float(x.norm(dim=-1).mean())))


if __name__ == "__main__":
    main()
