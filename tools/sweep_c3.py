"""BASELINE config 3: large-codebook sweep K x D at N = 1M frames on one B200 (fused mode, no p_code).
Times the tcgen05 search forward and the scatter-add backward through the C ABI with CUDA events and
prints one JSON line per point (tensor-pipe and HBM roofline fractions)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import semi_tts_b200 as V  # noqa: E402
from semi_tts_b200 import _lib, functional as VF  # noqa: E402


def main(quiet=False):
    N = int(os.environ.get("VQB_SWEEP_N", 1 << 20))
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.isfile(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    tf32_peak = peaks["bf16_tflops"] / 2.0          # kind::tf32 runs at half the dense bf16 rate
    lib = _lib.load()
    out = []
    only = os.environ.get("VQB_SWEEP_POINTS")                  # e.g. "256x64,8192x256" (K x D)
    only = {tuple(int(v) for v in t.split("x")) for t in only.split(",")} if only else None
    for D in (64, 256):
        if only and not any(d == D for _, d in only):
            continue
        g = torch.Generator().manual_seed(D)
        xs = [torch.randn(N, D, generator=g).cuda() for _ in range(2)]          # 2 x (N*D*4) >= 512 MB > L2
        gq = torch.randn(N, D, generator=g).cuda()
        for K in (256, 1024, 4096, 8192):
            if only and (K, D) not in only:
                continue
            e = torch.randn(K, D, generator=g).cuda()
            tab, enorm, _ = VF.assemble_table(e)
            temp = torch.ones(1, device="cuda")
            idx = torch.empty(N, dtype=torch.int64, device="cuda")
            q = torch.empty(N, D, device="cuda")
            stats = torch.zeros(2, dtype=torch.int32, device="cuda")
            a = _lib.FwdArgs(); a.struct_size = ctypes.sizeof(_lib.FwdArgs)
            a.flags = _lib.SCORE_L2 | _lib.STOP_GRAD | _lib.SEARCH_TENSOR
            a.n_rows, a.dim, a.n_codes = N, D, K
            a.score_w, a.score_b, a.gather_table, a.temp = tab.data_ptr(), enorm.data_ptr(), tab.data_ptr(), temp.data_ptr()
            a.idx, a.new_latent, a.search_stats = idx.data_ptr(), q.data_ptr(), stats.data_ptr()
            nb = ctypes.c_size_t(0)
            a.x = xs[0].data_ptr()
            _lib.check(lib.vqb_forward_workspace(ctypes.byref(a), ctypes.byref(nb)))
            ws = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
            a.workspace, a.workspace_bytes = ws.data_ptr(), nb.value
            sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 6 if K * D >= 4096 * 256 else 10
            for i in range(2):
                a.x = xs[i % 2].data_ptr(); _lib.check(lib.vqb_forward(ctypes.byref(a), sp))
            stats.zero_()
            torch.cuda.synchronize()
            ev0.record()
            for i in range(iters):
                a.x = xs[i % 2].data_ptr(); _lib.check(lib.vqb_forward(ctypes.byref(a), sp))
            ev1.record(); torch.cuda.synchronize()
            fwd_ms = ev0.elapsed_time(ev1) / iters
            st = (stats.cpu().float() / iters).tolist()
            def timed_fwd():
                for i in range(2):
                    a.x = xs[i % 2].data_ptr(); _lib.check(lib.vqb_forward(ctypes.byref(a), sp))
                torch.cuda.synchronize(); ev0.record()
                for i in range(iters):
                    a.x = xs[i % 2].data_ptr(); _lib.check(lib.vqb_forward(ctypes.byref(a), sp))
                ev1.record(); torch.cuda.synchronize()
                return ev0.elapsed_time(ev1) / iters
            ab = {}
            idx_ref = idx.clone()
            if os.environ.get("VQB_SWEEP_PIPE_AB"):              # software-pipelined x_lo (streamed 3xTF32 search) forced off
                lib.vqb_debug_set_search_pipe(0)
                ab["fwd_ms_search_nopipe"] = timed_fwd()
                lib.vqb_debug_set_search_pipe(-1)
                assert torch.equal(idx, idx_ref), "the pipelined search and the plain one disagree"
            # scatter-add backward (codebook gradient + histogram), idx from the last forward
            dtab = torch.zeros(K, D, device="cuda"); hist = torch.zeros(K, dtype=torch.int64, device="cuda")
            nbs = ctypes.c_size_t(0)
            _lib.check(lib.vqb_scatter_workspace(N, K, D, ctypes.byref(nbs)))
            wss = torch.empty(max(nbs.value, 1), dtype=torch.uint8, device="cuda")
            for _ in range(2):
                _lib.check(lib.vqb_scatter_add(idx.data_ptr(), N, gq.data_ptr(), K, D, dtab.data_ptr(), hist.data_ptr(), wss.data_ptr(), nbs.value, sp))
            torch.cuda.synchronize(); ev0.record()
            for _ in range(iters):
                _lib.check(lib.vqb_scatter_add(idx.data_ptr(), N, gq.data_ptr(), K, D, dtab.data_ptr(), hist.data_ptr(), wss.data_ptr(), nbs.value, sp))
            ev1.record(); torch.cuda.synchronize()
            bwd_ms = ev0.elapsed_time(ev1) / iters
            flops = 2.0 * N * K * D
            fwd_bytes = N * (8 * D + 8)
            bwd_bytes = N * (4 * D + 8)
            rec = {"N": N, "K": K, "D": D, "fwd_ms": fwd_ms, "search_tflops": flops / fwd_ms / 1e9,
                   "tensor_frac_of_tf32_peak": flops / fwd_ms / 1e9 / tf32_peak, "tf32_peak_tflops": tf32_peak,
                   "tensor_frac_of_dense_bf16_peak": flops / fwd_ms / 1e9 / peaks["bf16_tflops"],
                   "fwd_gbs": fwd_bytes / fwd_ms / 1e6, "fwd_hbm_frac": fwd_bytes / fwd_ms / 1e6 / peaks["hbm_gbs"],
                   "reranked_rows": st[0], "full_scan_rows": st[1],
                   "scatter_ms": bwd_ms, "scatter_gbs": bwd_bytes / bwd_ms / 1e6,
                   "scatter_hbm_frac": bwd_bytes / bwd_ms / 1e6 / peaks["hbm_gbs"],
                   "frames_per_s_fwd_bwd": N / ((fwd_ms + bwd_ms) * 1e-3)}
            rec.update(ab)
            if not quiet:
                print(json.dumps(rec), flush=True)
            out.append(rec)
    return out


if __name__ == "__main__":
    main()
