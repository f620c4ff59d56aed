"""Shared helpers for the parity tests: build the drop-in modules from golden state dicts."""
import os
import tempfile

import numpy as np
import torch

from conftest import GOLDEN

_TSV = None


def phn_attr_tsv():
    """A TSV in the reference's data/phn_attr.csv format, rebuilt from the committed 43x31 golden table
    (tests/golden/phn_attr_table.npy = output of the reference's read_phn_attr)."""
    global _TSV
    if _TSV is None or not os.path.isfile(_TSV):
        tab = np.load(os.path.join(GOLDEN, "phn_attr_table.npy"))
        fd, path = tempfile.mkstemp(suffix=".tsv", prefix="phn_attr_")
        with os.fdopen(fd, "w") as f:
            f.write("\t" + "\t".join("a%d" % i for i in range(tab.shape[1])) + "\n")
            for r, row in enumerate(tab[3:]):
                f.write("p%d\t" % r + "\t".join(str(int(v)) for v in row) + "\n")
        _TSV = path
    return _TSV


def codebook_kwargs(g, stop_grad=True, temp=None, skip_prob=0, bone="l2"):
    has_attr = "sd.phn_attr.weight" in g
    if bone == "l2":
        D = g["sd.learnable_table"].shape[1] + (g["sd.proj_attr.weight"].shape[0] if has_attr else 0)
    else:
        D = g["sd.asr_final_layer.weight"].shape[1]
    return dict(softmax="normal", latent_dim=D, commit_weight=0, vq_weight=0,
                temp=float(g["sd.temp"][0]) if temp is None else temp, skip_prob=skip_prob,
                stop_grad=stop_grad, phn_attr_pth=phn_attr_tsv() if has_attr else None,
                proj_attr=g["sd.proj_attr.weight"].shape[0] if has_attr else None)


def state_dict_of(g):
    return {k[3:]: torch.from_numpy(np.asarray(v).copy()) for k, v in g.items() if k.startswith("sd.")}


def build_module(g, bone="l2", stop_grad=True, device="cuda", learn_temp=False, skip_prob=0):
    import semi_tts_b200 as V
    K = (g["sd.learnable_table"] if bone == "l2" else g["sd.asr_final_layer.weight"]).shape[0]
    kw = codebook_kwargs(g, stop_grad=stop_grad, temp=-1 if learn_temp else None, skip_prob=skip_prob, bone=bone)
    cls = V.L2Embedding if bone == "l2" else V.SeperateEmbedding
    m = cls(K, False, **kw)
    m.load_state_dict(state_dict_of(g), strict=True)       # strict, as bin/train_vqvae.py:106
    return m.to(device)
