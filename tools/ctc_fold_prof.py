"""a few eager steps of config 2 with the CTC input fused in (codebook.ctc_eps) -- the target of an ncu capture"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                  # noqa: E402
import bench                  # noqa: E402
import semi_tts_b200 as V     # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = V.L2Embedding(bench.K, False, **bench._codebook_kwargs()).to(dev).train()
m.ctc_eps = None if os.environ.get("NO_FUSE") else 1e-10
B, S, K, D = 64, 800, bench.K, bench.D
for i in range(6):
    x = torch.randn(B, S, D, device=dev, requires_grad=True)
    p, q, _, _ = m(x)
    logp = m.ctc_logp if m.ctc_eps else V.ctc_log_probs(p)
    torch.autograd.backward([logp, q], [torch.randn(S, B, K, device=dev), torch.randn(B, S, D, device=dev)])
torch.cuda.synchronize()
