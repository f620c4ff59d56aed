"""Small driver for ncu: runs the fused tensor-core search a few times on one shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import semi_tts_b200 as V
K, D = int(sys.argv[1]), int(sys.argv[2])
N = int(sys.argv[3]) if len(sys.argv) > 3 else 148 * 128 * 4
g = torch.Generator().manual_seed(0)
x = torch.randn(N, D, generator=g).cuda(); e = torch.randn(K, D, generator=g).cuda()
for _ in range(3):
    idx, q = V.vq_search(x, e, search_tensor=True)
torch.cuda.synchronize()
print("ok", idx[:4].tolist())
