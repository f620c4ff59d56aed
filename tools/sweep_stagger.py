"""Developer sweep: start stagger of the co-resident CTAs of the parity-mode kernels (vqb_debug_set_stagger) at BASELINE
config 2.  For each setting: CUDA-event time of each kernel alone (bench._time_kernels) and of the graph-replayed step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                 # noqa: E402
import bench                 # noqa: E402


def main():
    import semi_tts_b200 as V
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = V.L2Embedding(bench.K, False, **bench._codebook_kwargs()).to(dev)
    m.train()
    sets = [bench._inputs(i, dev) for i in range(bench.RING)]
    for s in sets:
        s[0].requires_grad_(True)
    lib = V._lib.load()
    pts = [(int(a), int(b)) for a, b in (x.split(":") for x in os.environ.get(
        "VQB_STAGGER_POINTS", "0:0,1000:0,2000:0,3000:0,4000:0,0:1500,0:2500,0:3500,2000:2500").split(","))]

    def step(s):
        p, q, _, _ = m(s[0])
        torch.autograd.backward([p, q], [s[1], s[2]])

    for fns, bns in pts:
        lib.vqb_debug_set_stagger(fns, bns)
        fwd_ms, bwd_ms, _, _ = bench._time_kernels(V, m, sets, iters=16)
        for p_ in m.parameters():
            p_.grad = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(sets[0])
        torch.cuda.current_stream().wait_stream(side)
        graphs, pool = [], None
        for s in sets:
            for p_ in m.parameters():
                p_.grad = None
            s[0].grad = None
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                step(s)
            pool = g.pool()
            graphs.append(g)
        for i in range(16):
            graphs[i % len(graphs)].replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(200):
            graphs[i % len(graphs)].replay()
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"fwd_stagger_ns": fns, "bwd_stagger_ns": bns, "fwd_us": fwd_ms * 1e3, "bwd_us": bwd_ms * 1e3,
                          "step_us": e0.elapsed_time(e1) * 1e3 / 200}), flush=True)
        del graphs
    lib.vqb_debug_set_stagger(0, 0)


if __name__ == "__main__":
    main()
