import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from conftest import load_golden
from helpers import build_module
from semi_tts_b200 import _lib
g = load_golden("l2_attr_stopgrad")
m = build_module(g, "l2"); m.eval()
m.track_usage = os.environ.get("NOHIST") is None
x = torch.randn(64, 800, 64, device="cuda")
lib = _lib.load()
buf = torch.zeros(128, dtype=torch.int64, device="cuda")
names = {1: "start", 2: "tile_begin", 3: "x_full", 4: "xlo_done", 5: "t_full", 6: "softmax_done", 7: "gather_done", 8: "stores_issued", 9: "store_read_done", 10: "end", 11: "idx_written", 12: "after_bar", 13: "gather_loop_done"}
with torch.no_grad():
    for _ in range(3): m(x)
    torch.cuda.synchronize()
    lib.vqb_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(); m(x); ev1.record()
    torch.cuda.synchronize()
    lib.vqb_debug_set_timeline(None)
print("module fwd call (all kernels) %.1f us" % (ev0.elapsed_time(ev1) * 1e3))
raw = buf.cpu().tolist()
v = [int(t) for t in raw[:120] if t != 0]
print("entry->sync1 %.2f us, sync1->setup_done %.2f us, setup_done->exit %.2f us; first epilogue mark at +%.2f us after entry" % ((raw[121]-raw[120])/1e3, (raw[122]-raw[121])/1e3, (raw[123]-raw[122])/1e3, ((v[0] & ((1<<56)-1)) - raw[120])/1e3))
t0 = v[0] & ((1 << 56) - 1)
prev = t0
for e in v:
    tag, t = (e >> 56) & 0xFF, e & ((1 << 56) - 1)
    print("%-16s +%7.2f us  (d %6.2f)" % (names.get(tag, tag), (t - t0) / 1e3, (t - prev) / 1e3)); prev = t
