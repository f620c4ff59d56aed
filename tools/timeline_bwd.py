import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden
from helpers import build_module
from semi_tts_b200 import _lib
g = load_golden("l2_attr_stopgrad")
m = build_module(g, "l2"); m.train()
x = torch.randn(64, 800, 64, device="cuda", requires_grad=True)
gp = torch.randn(64, 800, 43, device="cuda"); gq = torch.randn(64, 800, 64, device="cuda")
lib = _lib.load()
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
names = {1: "start", 2: "tile_begin", 3: "pg_full", 4: "softmax_bwd_done", 5: "operands_written", 6: "colsum_done", 7: "d1_done", 8: "g_and_scatter_ready", 9: "dx_written", 10: "mma_done", 11: "tile_end", 12: "loop_end", 13: "flushed"}
for _ in range(3):
    p, q, _, _ = m(x); torch.autograd.backward([p, q], [gp, gq])
torch.cuda.synchronize()
p, q, _, _ = m(x)
torch.cuda.synchronize()
lib.vqb_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
torch.autograd.backward([p, q], [gp, gq])
torch.cuda.synchronize()
lib.vqb_debug_set_timeline(None)
names.update({20: "acc_tile_begin", 21: "sorted_list_ready", 22: "g_full", 23: "acc_done", 30: "sort_tile_begin", 31: "sort_buffer_free", 32: "codes_loaded", 33: "sorted"})
raw = buf.cpu().tolist()
v = [int(t) for t in raw[:120] if t != 0]
v = sorted(v, key=lambda e: e & ((1 << 56) - 1))
t0 = v[0] & ((1 << 56) - 1); prev = t0
for e in v:
    tag, t = (e >> 56) & 0xFF, e & ((1 << 56) - 1)
    print("%-18s +%7.2f us  (d %6.2f)" % (names.get(tag, tag), (t - t0) / 1e3, (t - prev) / 1e3)); prev = t
