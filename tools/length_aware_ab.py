"""SURVEY 8f rank 4, measured: the quantizer's forward + backward on a zero-padded batch with config 4's length distribution
(64 utterances, T ~ U[300, 800] mel frames -> 150..400 encoder frames padded to 400), dense vs `lengths=` (pad rows masked,
pad-only tiles skipped).  CUDA-graph replays, events; prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch                  # noqa: E402
import bench                  # noqa: E402


def main():
    import semi_tts_b200 as V
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = V.L2Embedding(bench.K, False, **bench._codebook_kwargs()).to(dev).train()
    B, S, K, D = 64, 400, bench.K, bench.D
    g = torch.Generator().manual_seed(7)
    lens = (torch.randint(300, 801, (B,), generator=g) // 2).clamp(max=S)
    lens[0] = S
    sets = [[torch.randn(B, S, D, generator=g).to(dev).requires_grad_(True), torch.randn(B, S, K, generator=g).to(dev),
             torch.randn(B, S, D, generator=g).to(dev)] for _ in range(8)]
    lens_d = lens.to(dev)
    out = {"workload": "64 x 400 encoder frames, lengths ~ U[150, 400] (config 4's distribution), K=43 D=64, fwd+bwd",
           "pad_row_fraction": float(1.0 - lens.sum().item() / (B * S))}
    for name, ln in (("dense", None), ("length_aware", lens_d)):
        def step(s):
            p, q, _, _ = m(s[0], 0, lengths=ln) if ln is not None else m(s[0])
            torch.autograd.backward([p, q], [s[1], s[2]])
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
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, pool=pool):
                step(s)
            pool = gr.pool()
            graphs.append(gr)
        for i in range(16):
            graphs[i % 8].replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(200):
            graphs[i % 8].replay()
        e1.record()
        torch.cuda.synchronize()
        out[name + "_us_per_step"] = e0.elapsed_time(e1) * 1e3 / 200
        del graphs
    print(json.dumps(out))


if __name__ == "__main__":
    main()
