"""In-kernel %globaltimer marks of CTA 0 of the parity-mode kernels (vqb_fwd_pcode_kernel, vqb_bwd_pcode_kernel) at
BASELINE config 2 (64 x 800 frames): per-tile phases of thread 0.  Developer tool (vqb_debug_set_timeline)."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                                    # noqa: E402
from conftest import load_golden                # noqa: E402
from helpers import build_module                # noqa: E402
from semi_tts_b200 import _lib                  # noqa: E402

FWD = {1: "setup_done", 2: "loads_issued", 3: "x_full", 4: "split_done", 5: "mma_done", 6: "softmax_done", 7: "gather_done",
       8: "store_read_done", 9: "end"}
BWD = {1: "start", 2: "in_full", 3: "staging_consumed", 4: "operands_done", 5: "d1_done", 6: "mma_done", 7: "dx_staged",
       8: "tile_end", 9: "loop_end", 10: "record_written"}


def dump(title, raw, names):
    v = [int(t) for t in raw[:60] if t != 0]
    print("== %s (%d marks)" % (title, len(v)))
    if not v:
        return
    mask = (1 << 56) - 1
    t0 = (raw[60] & mask) if raw[60] else (v[0] & mask)
    prev = t0
    for e in v:
        tag, t = (e >> 56) & 0xFF, e & mask
        print("%-18s +%7.2f us  (d %6.2f)" % (names.get(tag, tag), (t - t0) / 1e3, (t - prev) / 1e3))
        prev = t
    if raw[61]:
        print("%-18s +%7.2f us" % ("kernel_exit", ((raw[61] & mask) - t0) / 1e3))


def spans(title, raw):
    """per-CTA (start, end) marks at raw[128 + 2 b], raw[129 + 2 b]: when do the CTAs start, how long do they live"""
    st = [raw[128 + 2 * b] for b in range(960) if raw[128 + 2 * b] and raw[129 + 2 * b]]
    en = [raw[129 + 2 * b] for b in range(960) if raw[128 + 2 * b] and raw[129 + 2 * b]]
    if not st:
        return
    t0 = min(st)
    starts = sorted((s - t0) / 1e3 for s in st)
    life = sorted((e - s) / 1e3 for s, e in zip(st, en))
    q = lambda v, f: v[min(len(v) - 1, int(f * len(v)))]
    print("== %s: %d CTAs; start after the first CTA: median %.2f us, p90 %.2f, max %.2f; lifetime: min %.2f, median %.2f, p90 %.2f, "
          "max %.2f us; last exit at +%.2f us" % (title, len(st), q(starts, .5), q(starts, .9), starts[-1], life[0], q(life, .5),
                                                  q(life, .9), life[-1], (max(en) - t0) / 1e3))


def main():
    g = load_golden("l2_attr_stopgrad")
    m = build_module(g, "l2")
    m.train()
    lib = _lib.load()
    x = torch.randn(64, 800, 64, device="cuda", requires_grad=True)
    gp = torch.randn(64, 800, 43, device="cuda")
    gq = torch.randn(64, 800, 64, device="cuda")
    buf = torch.zeros(2048, dtype=torch.int64, device="cuda")
    for _ in range(3):
        p, q, _, _ = m(x)
        torch.autograd.backward([p, q], [gp, gq])
    torch.cuda.synchronize()
    lib.vqb_debug_set_timeline(ctypes.c_void_p(buf.data_ptr()))
    e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
    e0.record()
    p, q, _, _ = m(x)
    e1.record()
    torch.cuda.synchronize()
    fwd = buf.cpu().tolist()
    buf.zero_()
    torch.cuda.synchronize()
    e2.record()
    torch.autograd.backward([p, q], [gp, gq])
    e3.record()
    torch.cuda.synchronize()
    bwd = buf.cpu().tolist()
    lib.vqb_debug_set_timeline(None)
    print("module forward call %.1f us, backward call %.1f us (eager, all kernels + host)" % (e0.elapsed_time(e1) * 1e3,
                                                                                         e2.elapsed_time(e3) * 1e3))
    dump("forward, CTA 0 thread 0", fwd, FWD)
    spans("forward", fwd)
    dump("backward, CTA 0 thread 0", bwd, BWD)
    spans("backward", bwd)


if __name__ == "__main__":
    main()
