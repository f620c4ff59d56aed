"""SURVEY 8f rank 3, measured: config 2's quantizer step when the CTC input is part of it -- log(p_code + EPS) in [S, B, K]
and the gradient that comes back for it (bin/train_vqvae.py:430-432) -- as (a) the standalone pass ctc_log_probs + its
backward pass around the quantizer, (b) emitted by the forward kernel's epilogue with the division folded into the backward
kernel (`codebook.ctc_eps`).  CUDA-graph replays over rotating input sets, events; prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch                  # noqa: E402
import bench                  # noqa: E402


def main():
    import semi_tts_b200 as V
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = V.L2Embedding(bench.K, False, **bench._codebook_kwargs()).to(dev).train()
    B, S, K, D = 64, 800, bench.K, bench.D
    g = torch.Generator().manual_seed(7)
    nset = 12                                                   # 12 x (13 + 8.8 + 13) MB: larger than L2 between reuses
    sets = [[torch.randn(B, S, D, generator=g).to(dev).requires_grad_(True), torch.randn(S, B, K, generator=g).to(dev),
             torch.randn(B, S, D, generator=g).to(dev)] for _ in range(nset)]
    out = {"workload": "config 2 (64 x 800 frames, K=43, D=64): forward + CTC input + backward of both outputs"}
    for name in ("standalone_pass", "fused"):
        m.ctc_eps = 1e-10 if name == "fused" else None

        def step(s):
            p, q, _, _ = m(s[0])
            logp = m.ctc_logp if name == "fused" else V.ctc_log_probs(p)
            torch.autograd.backward([logp, q], [s[1], s[2]])
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
        for i in range(2 * nset):
            graphs[i % nset].replay()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for i in range(240):
            graphs[i % nset].replay()
        e1.record()
        torch.cuda.synchronize()
        out[name + "_us_per_step"] = e0.elapsed_time(e1) * 1e3 / 240
        del graphs
    for name in ("standalone_pass", "fused"):
        m.ctc_eps = 1e-10 if name == "fused" else None
        # the two dominant kernels alone (events recorded by the library right around each launch; eager mode)
        import ctypes
        from semi_tts_b200 import _lib
        lib = _lib.load()
        lib.vqb_debug_set_kernel_events.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        lib.vqb_debug_set_kernel_events.restype = None
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        for e in ev:
            e.record()
        fw, bw = [], []
        for i in range(3 * nset):
            s = sets[i % nset]
            lib.vqb_debug_set_kernel_events(ctypes.c_void_p(ev[0].cuda_event), ctypes.c_void_p(ev[1].cuda_event))
            p, q, _, _ = m(s[0])
            logp = m.ctc_logp if name == "fused" else V.ctc_log_probs(p)
            lib.vqb_debug_set_kernel_events(ctypes.c_void_p(ev[2].cuda_event), ctypes.c_void_p(ev[3].cuda_event))
            torch.autograd.backward([logp, q], [s[1], s[2]])
            lib.vqb_debug_set_kernel_events(None, None)
            torch.cuda.synchronize()
            if i >= nset:
                fw.append(ev[0].elapsed_time(ev[1]) * 1e3); bw.append(ev[2].elapsed_time(ev[3]) * 1e3)
        out[name + "_fwd_kernel_us"] = sum(fw) / len(fw)
        out[name + "_bwd_kernel_us"] = sum(bw) / len(bw)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
