import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden, rel_err
from helpers import build_module
from oracle import vq_oracle as O
g = load_golden("l2_attr_stopgrad")
B, S = 2, 64
gen = torch.Generator().manual_seed(1)
x_cpu = torch.randn(B, S, 64, generator=gen); gp_cpu = torch.randn(B, S, 43, generator=gen); gq_cpu = torch.randn(B, S, 64, generator=gen)
res = {}
for tc in (True, False):
    m = build_module(g, "l2"); m.tensor_cores = tc
    x = x_cpu.cuda().requires_grad_(True)
    p, q, _, _ = m(x)
    torch.autograd.backward([p, q], [gp_cpu.cuda(), gq_cpu.cuda()])
    res[tc] = (x.grad.cpu().numpy().reshape(-1, 64), m.learnable_table.grad.cpu().numpy(), m.proj_attr.weight.grad.cpu().numpy(), p.detach().cpu().numpy().reshape(-1, 43))
E = O.assemble_table(g["sd.learnable_table"], g["sd.phn_attr.weight"], g["sd.proj_attr.weight"], g["sd.proj_attr.bias"])
P = res[False][3].astype(np.float64); G = gp_cpu.numpy().reshape(-1, 43).astype(np.float64)
s = (G * P).sum(-1, keepdims=True); C = -(P * (G - s))
xf = x_cpu.numpy().reshape(-1, 64).astype(np.float64); gq = gq_cpu.numpy().reshape(-1, 64).astype(np.float64)
base = gq + 2 * xf * C.sum(-1, keepdims=True)
gem_ref = C @ E
gem_tc = (base - res[True][0]) / 2
gem_simt = (base - res[False][0]) / 2
print("simt gemm err", rel_err(gem_simt, gem_ref), " tc gemm err", rel_err(gem_tc, gem_ref))
for lo, hi in ((0, 32), (32, 64)):
    print("cols", lo, hi, rel_err(gem_tc[:, lo:hi], gem_ref[:, lo:hi]))
for r0 in range(0, 128, 32):
    print("rows", r0, rel_err(gem_tc[r0:r0+32], gem_ref[r0:r0+32]))
# hypothesis tests
Ehi = E.astype(np.float32).view(np.uint32) & 0xFFFFE000; Ehi = Ehi.view(np.float32).astype(np.float64)
print("vs C@E_hi(trunc)", rel_err(gem_tc, C @ Ehi))
print("vs C[:, :32]@E[:32]", rel_err(gem_tc, C[:, :32] @ E[:32]), " vs C[:,32:]@E[32:]", rel_err(gem_tc, C[:, 32:] @ E[32:]))
# least squares: find M such that gem_tc = C @ M
M, *_ = np.linalg.lstsq(C, gem_tc, rcond=None)
print("lstsq residual", rel_err(C @ M, gem_tc), " M vs E", rel_err(M, E))
d = np.abs(M - E); print("rows of M differing most:", np.argsort(-d.sum(1))[:10], " cols:", np.argsort(-d.sum(0))[:10])
np.set_printoptions(precision=3, suppress=True, linewidth=200)
print("M[:6,:8]\n", M[:6, :8], "\nE[:6,:8]\n", E[:6, :8])
print("dW: tc vs simt", rel_err(res[True][1], res[False][1]), rel_err(res[True][2], res[False][2]))
print((res[True][1] / (res[False][1] + 1e-30))[:4, :8])
