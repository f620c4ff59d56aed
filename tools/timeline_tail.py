"""Timeline of the fused backward tail (bwd_tail_kernel), one process per GPU under torchrun (or alone):
%globaltimer marks of block 0 / the elected last block, relative to the end of the main backward kernel's CTA 0."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import torch.distributed as dist
import semi_tts_b200 as V
from helpers import phn_attr_tsv
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
m = V.L2Embedding(43, False, softmax="normal", latent_dim=64, commit_weight=0, vq_weight=0, temp=1, skip_prob=0,
                  stop_grad=True, phn_attr_pth=phn_attr_tsv(), proj_attr=16).to(dev)
m.train()
if world > 1:
    V.dist.enable_fused_allreduce(m)
g = torch.Generator().manual_seed(rank)
x = torch.randn(64, 800, 64, generator=g).to(dev).requires_grad_(True)
gp, gq = torch.randn(64, 800, 43, generator=g).to(dev), torch.randn(64, 800, 64, generator=g).to(dev)
lib = V._lib.load()
buf = torch.zeros(2048, dtype=torch.int64, device=dev)     # vqb_debug_set_timeline: 2048 u64 (phase marks + per-CTA start/end)
for it in range(6):
    for p_ in m.parameters(): p_.grad = None
    p, q, _, _ = m(x)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    if it == 5:
        lib.vqb_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
    torch.autograd.backward([p, q], [gp, gq])
    torch.cuda.synchronize()
lib.vqb_debug_set_timeline(None)
raw = buf.cpu().tolist()
names = ["tail block entry", "past pdl_wait (main kernel complete)", "own outputs finished", "pushed to all peers",
         "peers' words gathered, sum stored"]
t0 = raw[100]
for r in range(world):
    if world > 1:
        dist.barrier()
    if r == rank:
        print("rank %d (last block of the tail kernel; t = 0 at its entry)" % rank)
        prev = t0
        for i, nme in enumerate(names):
            t = raw[100 + i]
            if t:
                print("  %-36s +%7.2f us  (d %6.2f)" % (nme, (t - t0) / 1e3, (t - prev) / 1e3)); prev = t
        sys.stdout.flush()
if world > 1:
    dist.barrier(); torch.cuda.synchronize()
os._exit(0)
