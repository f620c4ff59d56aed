import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from conftest import load_golden
from helpers import build_module
name = sys.argv[1] if len(sys.argv) > 1 else "l2_attr_skip_train"
g = load_golden(name)
skip = name == "l2_attr_skip_train"
m = build_module(g, "l2", skip_prob=1.0 if skip else 0); m.train(True)
x0 = torch.from_numpy(g["x"]).cuda()
gp, gq = torch.from_numpy(g["g_p"]).cuda(), torch.from_numpy(g["g_q"]).cuda()
res = {}
for fused in (False, True, True, False, True):
    m.fused_tail.enabled = fused
    for p_ in m.parameters(): p_.grad = None
    x = x0.clone().requires_grad_(True)
    p, q, _, _ = m(x, int(g["first_n_real_mel"]))
    torch.autograd.backward([p, q], [gp, gq])
    torch.cuda.synchronize()
    dx = x.grad.clone()
    if not fused:
        res["dx"] = dx; res["lt"] = m.learnable_table.grad.clone()
    else:
        d = (dx - res["dx"]).abs()
        bad = (d.view(-1, 64).max(dim=1).values > 1e-5).nonzero().flatten().tolist()
        print("fused used:", m.fused_tail.fused, "max |dx diff|", float(d.max()), "bad rows", bad[:10], "...", len(bad),
              "| lt diff", float((m.learnable_table.grad - res["lt"]).abs().max()))
        if bad:
            r = bad[0]
            print(" row", r, "got", dx.view(-1, 64)[r, :6].tolist(), "want", res["dx"].view(-1, 64)[r, :6].tolist(), "g_q", gq.view(-1, 64)[r, :6].tolist())
