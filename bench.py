#!/usr/bin/env python
"""bench.py -- VQ bottleneck fwd+bwd frames/s (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus 1 ...            # the reference's CPU op sequence (oracle port)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (N > 1)

A "step" is one forward + backward of the quantizer over one batch of synthetic encoder frames
(BASELINE.json configs[1]: 64 x 800 frames, K=43, D=64, phoneme-attribute codebook of
config/semi-multi-spkr-paired-data.yaml), with upstream gradients for BOTH outputs (g_p for p_code, g_q for
new_latent), as bin/train_vqvae.py drives it.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np   # noqa: E402
import torch         # noqa: E402

B, S, K, D, DA, A = 64, 800, 43, 64, 16, 31
N_ROWS = B * S
RING = 8             # distinct input/output sets cycled through so that no step finds its inputs in L2
WORKLOAD = "semi-tts L2 quantizer fwd+bwd, batch 64 x 800 frames, K=43 D=64 (config/semi-multi-spkr-paired-data.yaml codebook)"


def _config(world):
    """The `config` object of the JSON line: built from the workload constants and the world size only, so that the GPU arm
    and the reference arm print the SAME dict."""
    fwd_bytes = N_ROWS * (8 * D + 8 + 4 * K)
    bwd_bytes = N_ROWS * (12 * D + 8 * K + 8)
    return {"workload": WORKLOAD, "rows_per_step_per_gpu": N_ROWS, "grads": "g_p+g_q",
            "launch": "GPU arm: one CUDA graph per step (assemble, forward, backward, tail); reference arm: eager torch CPU ops",
            "l2_policy": "ring of %d distinct input/output sets (%.0f MB touched per ring pass) > 126 MB L2" % (
                RING, RING * (fwd_bytes + bwd_bytes) / 1e6),
            "parallelism": "dp%d (frames sharded by batch, codebook replicated; codebook-gradient sum over GPUs by this library's "
                           "own exchange kernel over NVLink peer memory (push, poll, rank-ordered sum): step i's exchange runs on "
                           "a side stream beside step i+1 and is joined at that step's end, the last one inside the timed "
                           "region; the usage histogram is exchanged every 500 steps, the cadence at which the trainer reads it "
                           "(bin/train_vqvae.py:305), so a timed window shorter than that holds no histogram exchange)" % world}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel` at this workload, from the committed
    `ncu --set full` summary (profiles/ncu_traffic.json, written by tools/ncu_summary.py); None if not captured."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(p):
        d = json.load(open(p)).get(kernel)
        if d:
            return d.get("dram_bytes_per_launch")
    return None


def _phn_attr_tsv():
    from helpers import phn_attr_tsv
    return phn_attr_tsv()


def _codebook_kwargs():
    return dict(softmax="normal", latent_dim=D, commit_weight=0, vq_weight=0, temp=1, skip_prob=0,
                stop_grad=True, phn_attr_pth=_phn_attr_tsv(), proj_attr=DA)


def _inputs(seed, device="cpu", pin=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, S, D, generator=g)
    gp = torch.randn(B, S, K, generator=g)
    gq = torch.randn(B, S, D, generator=g)
    if pin:
        x, gp, gq = x.pin_memory(), gp.pin_memory(), gq.pin_memory()
    return [t.to(device) for t in (x, gp, gq)] if device != "cpu" else [x, gp, gq]


class ClockSampler:
    """nvidia-smi sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower() == "active"})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's op sequence on the host cores (oracle/torch_port.py)
# --------------------------------------------------------------------------------------------------
def _cpu_state(seed=0):
    g = torch.Generator().manual_seed(seed)
    tab = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "phn_attr_table.npy")))
    lt = torch.randn(K, D - DA, generator=g).requires_grad_(True)
    bound = 1.0 / np.sqrt(A)
    pw = ((torch.rand(DA, A, generator=g) * 2 - 1) * bound).requires_grad_(True)
    pb = ((torch.rand(DA, generator=g) * 2 - 1) * bound).requires_grad_(True)
    return lt, tab.float(), pw, pb, torch.ones(1)


def _reference_module():
    """The UNMODIFIED reference quantizer (src/embed.py: L2Embedding) from the installed tree baseline/_ref (or
    /root/reference in the build container), built exactly as src/vqvae.py:57 builds it; None if no tree is present."""
    try:
        from oracle import ref_import
        if not ref_import.available():
            return None
        ref_embed = ref_import.import_reference()
        cb = ref_import.load_codebook_cfg("semi-multi-spkr-paired-data.yaml")
        cb.pop("bone")
        with ref_import.reference_cwd():
            torch.manual_seed(0)
            return ref_embed.L2Embedding(K, False, **cb)
    except Exception as e:                               # noqa: BLE001 -- fall back to the port, say why
        print("[bench] reference modules unavailable (%s); timing the port" % str(e).splitlines()[0], file=sys.stderr)
        return None


def cpu_fwd_bwd_rate(steps, warmup, budget_s=None):
    """frames/s of the reference quantizer (fwd + autograd bwd, both upstream grads) on all host threads: the reference's
    own module when it is installed (kind "reference"), otherwise the op-sequence port (kind "port")."""
    torch.set_num_threads(os.cpu_count() or 1)
    x, gp, gq = _inputs(1)
    x.requires_grad_(True)
    mod = _reference_module()
    if mod is not None:
        mod.train()
        kind = "reference"
        what = "the reference's own src/embed.py L2Embedding (installed copy baseline/_ref), torch autograd backward"

        def one():
            x.grad = None
            for p_ in mod.parameters():
                p_.grad = None
            p, q, _, _ = mod(x)                           # src/vqvae.py:119
            torch.autograd.backward([p, q], [gp, gq])     # src/solver.py:144
    else:
        from oracle import torch_port as TP
        lt, tab, pw, pb, temp = _cpu_state()
        kind = "port"
        what = "oracle/torch_port.py: the reference's ATen op sequence, src/embed.py:105-147, torch autograd backward"

        def one():
            TP.l2_step(x, lt, tab, pw, pb, temp, gp, gq)
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        one()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return N_ROWS * done / dt, dt / done * 1e3, done, torch.get_num_threads(), kind, what


def run_reference(args, rank):
    if rank != 0:
        return
    rate, ms, done, cores, kind, what = cpu_fwd_bwd_rate(args.steps, args.warmup, budget_s=120.0)
    line = {"impl": "reference", "metric": "vq_fwd_bwd_frames_per_sec", "value": rate, "unit": "frames/s",
            "n_gpus": args.gpus, "steps": done, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(args.gpus),
            "cpu_baseline": {"value": rate, "unit": "frames/s", "cores": cores, "kind": kind,
                             "sample": "%d full steps of the same workload (%s)" % (done, what)},
            "e2e": {"value": rate, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def _time_kernels(V, m, sets, iters=24):
    """CUDA-event time of the two dominant kernels launched alone through the C ABI (ring-rotated inputs)."""
    import ctypes
    from semi_tts_b200 import functional as VF, _lib
    attr, pw, pb = m.phn_attr.weight, m.proj_attr.weight, m.proj_attr.bias
    table, enorm, _, cache = VF.assemble_table(m.learnable_table, attr, pw, pb, want_cache=True)
    flags = _lib.SCORE_L2 | _lib.STOP_GRAD | (_lib.TENSOR_CORES if m.tensor_cores else 0)     # no AFTER_ASSEMBLE: launched alone
    N_ROWS = sets[0][0].numel() // D            # (config 2: 51 200; the steady-state record passes 2^20-row sets)
    outs = [VF._run_forward(flags, s[0].view(N_ROWS, D), table, enorm, table, m.temp, True, None, False) for s in sets]
    stream = torch.cuda.current_stream()
    fwd_ms, bwd_ms = [], []
    lib = _lib.load()
    # pre-build argument structs so only the launch is inside the events
    fa, ba, keep = [], [], []
    lib = _lib.load()
    for s, o in zip(sets, outs):
        p_code, idx, q, _ = o
        a = _lib.FwdArgs(); a.struct_size = ctypes.sizeof(_lib.FwdArgs); a.flags = flags
        a.n_rows, a.dim, a.n_codes = N_ROWS, D, K
        a.x, a.score_w, a.score_b, a.gather_table = s[0].data_ptr(), table.data_ptr(), enorm.data_ptr(), table.data_ptr()
        a.temp, a.p_code, a.idx, a.new_latent = m.temp.data_ptr(), p_code.data_ptr(), idx.data_ptr(), q.data_ptr()
        a.operand_cache = cache.data_ptr()          # as the module passes it (built once per step with the table)
        nb = ctypes.c_size_t(0)
        _lib.check(lib.vqb_forward_workspace(ctypes.byref(a), ctypes.byref(nb)))
        ws = torch.empty(max(nb.value, 1), dtype=torch.uint8, device="cuda")
        a.workspace, a.workspace_bytes = ws.data_ptr(), nb.value
        keep.append(ws)
        fa.append(a)
        dx = torch.empty(N_ROWS, D, device="cuda"); dw = torch.zeros(K, D, device="cuda"); cs = torch.zeros(K, device="cuda")
        b = _lib.BwdArgs(); b.struct_size = ctypes.sizeof(_lib.BwdArgs); b.flags = flags
        b.n_rows, b.dim, b.n_codes, b.n_real_rows = N_ROWS, D, K, 0
        b.x, b.score_w, b.score_b, b.gather_table, b.temp = s[0].data_ptr(), table.data_ptr(), enorm.data_ptr(), table.data_ptr(), m.temp.data_ptr()
        b.p_code, b.idx, b.g_p, b.g_q = p_code.data_ptr(), idx.data_ptr(), s[1].data_ptr(), s[2].data_ptr()
        b.dx, b.d_score_w, b.colsum = dx.data_ptr(), dw.data_ptr(), cs.data_ptr()
        b.operand_cache = cache.data_ptr()
        nb = ctypes.c_size_t(0)
        _lib.check(lib.vqb_backward_workspace(ctypes.byref(b), ctypes.byref(nb)))
        wsb = torch.empty(max(nb.value, 1), dtype=torch.uint8, device="cuda")
        b.workspace, b.workspace_bytes = wsb.data_ptr(), nb.value
        ba.append(b); keep.append((dx, dw, cs, wsb))
    sp = ctypes.c_void_p(stream.cuda_stream)
    # the library records these events on the launch stream right before / after the dominant kernel of each call
    # (vqb_debug_set_kernel_events), so helper kernels of the same call are outside the measured interval
    lib.vqb_debug_set_kernel_events.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    lib.vqb_debug_set_kernel_events.restype = None
    # launches go out back to back, as in the timed region (no host synchronisation between them: an isolated launch on an
    # idle GPU adds its own start-up latency to the interval); every launch has its own event pair
    n = iters + 4
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n)]
    for e4 in evs:
        for e in e4:
            e.record(stream)              # (torch creates the CUDA event on first use; the library records it again)
    torch.cuda.synchronize()
    for i in range(n):
        j = i % len(sets)
        e = evs[i]
        lib.vqb_debug_set_kernel_events(ctypes.c_void_p(e[0].cuda_event), ctypes.c_void_p(e[1].cuda_event))
        _lib.check(lib.vqb_forward(ctypes.byref(fa[j]), sp))
        lib.vqb_debug_set_kernel_events(ctypes.c_void_p(e[2].cuda_event), ctypes.c_void_p(e[3].cuda_event))
        _lib.check(lib.vqb_backward(ctypes.byref(ba[j]), sp))
    lib.vqb_debug_set_kernel_events(None, None)
    torch.cuda.synchronize()
    for e in evs[4:]:
        fwd_ms.append(e[0].elapsed_time(e[1])); bwd_ms.append(e[2].elapsed_time(e[3]))
    kf = lib.vqb_forward_kernel_name(ctypes.byref(fa[0])).decode()
    kb = lib.vqb_backward_kernel_name(ctypes.byref(ba[0])).decode()
    return statistics.mean(fwd_ms), statistics.mean(bwd_ms), kf, kb


def _c3_sweep(local_rank):
    """BASELINE configs[2]: the large-codebook sweep (K x D at N = 2^20 frames, fused mode) as a sub-record of the bench line:
    tensor-pipe fraction of the search kernel (stated against BOTH the measured dense bf16 peak and its tf32 half -- the
    MMA runs kind::tf32) and HBM fraction of the scatter, with the clocks sampled during the sweep."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sweep_c3
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    pts = sweep_c3.main(quiet=True)
    clocks = sampler.stop()
    keep = ("N", "K", "D", "fwd_ms", "search_tflops", "tensor_frac_of_tf32_peak", "tensor_frac_of_dense_bf16_peak",
            "reranked_rows", "full_scan_rows", "scatter_ms", "scatter_gbs", "scatter_hbm_frac", "frames_per_s_fwd_bwd")
    return {"workload": "configs[2]: K in {256,1024,4096,8192} x D in {64,256}, N = 2^20 frames, fused mode (no p_code): "
                        "tcgen05 search forward + scatter-add backward",
            "points": [{k: p[k] for k in keep if k in p} for p in pts], "clocks": clocks, "seconds": time.perf_counter() - t0}


def _log(rank, msg):
    print("[bench rank %d %.1fs] %s" % (rank, time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


_T0 = time.perf_counter()


def _bind_to_gpu_numa_node(local_rank):
    """Multi-rank runs: pin this process (and with it the pinned host buffers it allocates from here on: Linux allocates on
    the node of the running CPU) to the NUMA node the GPU's PCIe root hangs off, so that the per-step upload does not cross
    the socket interconnect.  Returns what was done, for the JSON line; None if the topology is not visible."""
    if os.environ.get("VQB_NO_NUMA_BIND"):
        return None
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        addr = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % addr).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"gpu": addr, "node": node, "cpus": len(cpus)}
    except Exception:                 # noqa: BLE001 -- no topology in this container: leave the process where it is
        return None


def run_ours(args, rank, world, local_rank):
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)      # a wedged collective must not eat the GPU budget
    import semi_tts_b200 as V
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (GPU arm) needs a B200; there is no CPU fallback -- use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    numa = None
    if dist_on:
        import datetime
        import torch.distributed as dist
        numa = _bind_to_gpu_numa_node(local_rank)       # (N = 1 keeps every core: the CPU baseline runs in this process)
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
        _log(rank, "process group up (world %d), numa binding %s" % (world, numa))
    torch.manual_seed(0)
    m = V.L2Embedding(K, False, **_codebook_kwargs()).to(dev)
    m.train()
    if os.environ.get("VQB_NO_TAIL"):
        m.fused_tail.enabled = False             # developer switch: three-kernel backward tail
    no_exchange = bool(os.environ.get("VQB_BENCH_NO_EXCHANGE"))    # developer A/B: N independent replicas, nothing exchanged
    if dist_on and not os.environ.get("VQB_NCCL_ALLREDUCE") and not no_exchange:
        V.dist.enable_fused_allreduce(m)         # gradient sum inside the backward's tail kernel (NVLink peer memory)
    # ring of distinct device-resident input sets (weak scaling: every rank owns RING x 64 x 800 frames)
    sets = [_inputs(1000 * rank + i, dev) for i in range(RING)]
    for s in sets:
        s[0].requires_grad_(True)

    # Deferred exchange (N > 1): the backward's tail kernel pushes this rank's flat gradient into every peer's buffer over
    # NVLink; the poll + rank-ordered sum of step i runs on a side stream behind step i+1's table assembly and forward, and is
    # joined before step i+1's backward (grads are consumed by optimizer.step, src/solver.py:149; in a trainer the rest of
    # the model's backward sits there).  The rank skew and the NVLink round trip are then off the step's critical path.
    deferred = dist_on and m.fused_tail.exchange is not None and not os.environ.get("VQB_NO_DEFER")
    m.fused_tail.defer = deferred
    side_x = torch.cuda.Stream() if deferred else None
    join_early = bool(os.environ.get("VQB_DEFER_JOIN_EARLY"))      # developer A/B: join the exchange before the backward

    def step(s):
        if deferred:
            cur = torch.cuda.current_stream()
            side_x.wait_stream(cur)
            V.dist.finish_codebook_grads(m, stream=side_x)      # the previous step's exchange, concurrent with this forward
        p, q, _, _ = m(s[0])
        if deferred and join_early:
            cur.wait_stream(side_x)
        torch.autograd.backward([p, q], [s[1], s[2]])
        if deferred and not join_early:
            cur.wait_stream(side_x)
        if dist_on and not deferred and not no_exchange:
            V.dist.allreduce_codebook_grads(m)

    # ---- value: device-resident inputs, whole step (fwd + bwd [+ all-reduce]) captured in CUDA graphs ------
    for p_ in m.parameters():
        p_.grad = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for s in sets[:3]:
            step(s)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    _log(rank, "warm-up done, capturing graphs")
    graphs, use_graph = [], True
    lib_ = V._lib.load()
    launches_per_step = 0
    try:
        pool = None
        V.dist.finish_codebook_grads(m)          # nothing pending when the first capture starts
        # deferred exchange: every graph finishes its ring predecessor's exchange, so set 0 is captured once more at the end
        # (then with set RING-1's exchange pending) and that second capture is the one replayed
        for s in (sets + [sets[0]] if deferred else sets):
            l0 = lib_.vqb_launch_count()
            for p_ in m.parameters():
                p_.grad = None                  # as after optimizer.zero_grad(): backward assigns, it does not accumulate
            s[0].grad = None
            g = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread must not invalidate the capture
            with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                step(s)
            launches_per_step = int(lib_.vqb_launch_count() - l0)      # this library's kernels in one captured step
            pool = g.pool()
            graphs.append(g)
        if deferred:
            _keep_first_capture = graphs[0]                             # its buffers stay valid: set 1's graph finishes into them
            graphs = [graphs[RING]] + graphs[1:RING]                    # the re-captured set 0, then sets 1 .. RING-1
    except Exception as e:           # noqa: BLE001 -- report and fall back to eager launches of the same kernels
        use_graph = False
        graphs = []
        torch.cuda.synchronize()
        _log(rank, "CUDA-graph capture failed (%s); timing eager launches" % str(e).splitlines()[0])
    if dist_on:
        # every rank must take the same path, or the collectives no longer match up
        ok = torch.tensor([1 if use_graph else 0], device=dev)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        if int(ok.item()) == 0:
            use_graph, graphs = False, []
    if not use_graph:
        l0 = lib_.vqb_launch_count()
        step(sets[0])
        launches_per_step = int(lib_.vqb_launch_count() - l0)
    _log(rank, "launch mode: %s, %d libvqb200 kernels per step" % ("cuda_graph" if use_graph else "eager", launches_per_step))

    USAGE_EVERY = 500                       # the trainer reads (and resets) the usage histogram every 500 steps (:305, :310)

    def run_steps(n, first=0):
        for i in range(n):
            if use_graph:
                graphs[(first + i) % RING].replay()
            else:
                step(sets[(first + i) % RING])
            if dist_on and (i + 1) % USAGE_EVERY == 0:
                V.dist.allreduce_usage(m)   # at the reference's own cadence, inside the timed region when it falls there

    def barrier():
        if dist_on:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    run_steps(max(args.warmup, 3))
    _log(rank, "timed warm-up done")
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    run_steps(args.steps, first=args.warmup)
    if dist_on:
        V.dist.finish_codebook_grads(m)    # the last step's exchange (deferred mode) completes inside the timed region
    e1.record()
    if dist_on:
        V.dist.allreduce_usage(m)          # what is pending of the current 500-step window (outside the timed region)
    barrier()
    ms_total = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms_total], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms_total = float(t.item())
    value = N_ROWS * world * args.steps / (ms_total * 1e-3)
    _log(rank, "device-resident timing done: %.3f ms/step" % (ms_total / args.steps))

    # ---- e2e: public module API, the step's HOST input (pinned) uploaded and its results read back inside the region ----
    # The module's input is enc_embs; the upstream gradients g_p / g_q are produced ON THE DEVICE by whatever consumes the
    # module's outputs (CTC loss, Tacotron: bin/train_vqvae.py:208,:174), so they stay device-resident here as well.
    # Every step: H2D x (13.1 MB), forward + backward through the nn.Module, D2H of the picked indices, the parameter
    # gradients and the usage histogram.
    host_x = [_inputs(5000 + 1000 * rank + i, "cpu", pin=True)[0] for i in range(4)]
    n_grad = sum(p_.numel() for p_ in m.parameters() if p_.requires_grad)
    h_idx = torch.empty(B, S, dtype=torch.int64).pin_memory()
    h_grad = torch.empty(n_grad + K, dtype=torch.float32).pin_memory()
    g_dev = [(s_[1], s_[2]) for s_ in sets[:4]]

    # double-buffered pipeline: a copy stream uploads step i+1's pinned host input while the compute stream runs step i
    NBUF = 3
    comp, copy_s = torch.cuda.current_stream(), torch.cuda.Stream()
    dbuf = [torch.empty_like(host_x[0], device=dev).requires_grad_(True) for _ in range(NBUF)]
    ready = [torch.cuda.Event() for _ in range(NBUF)]
    free = [torch.cuda.Event() for _ in range(NBUF)]

    def upload(i):
        with torch.cuda.stream(copy_s):
            copy_s.wait_event(free[i % NBUF])
            with torch.no_grad():
                dbuf[i % NBUF].copy_(host_x[i % 4], non_blocking=True)
            ready[i % NBUF].record(copy_s)

    def e2e_body(k):
        # one step through the public module API on upload buffer k: forward, backward, gradient sum over the ranks,
        # read-back of the picked indices, the parameter gradients and the usage histogram
        x = dbuf[k]
        gp, gq = g_dev[k]
        x.grad = None
        for p_ in m.parameters():
            p_.grad = None
        p, q, _, _ = m(x)
        torch.autograd.backward([p, q], [gp, gq])
        if dist_on:
            V.dist.allreduce_codebook_grads(m)
        h_idx.copy_(m.last_idx, non_blocking=True)
        flat = torch.cat([p_.grad.reshape(-1) for p_ in m.parameters() if p_.requires_grad] + [m.usage.counts.float()])
        h_grad.copy_(flat, non_blocking=True)

    # the same step captured once per upload buffer (what a user does around a fixed-shape training step): the eager
    # Python of forward + backward (~0.3 ms) would otherwise hide the PCIe transfer it is supposed to overlap with
    e2e_graphs = []
    if use_graph and not os.environ.get("VQB_E2E_EAGER"):
        try:
            e2e_body(0)
            torch.cuda.synchronize()
            pool2 = None
            for k in range(NBUF):
                g2 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g2, pool=pool2, capture_error_mode="thread_local"):
                    e2e_body(k)
                pool2 = g2.pool()
                e2e_graphs.append(g2)
        except Exception as e:       # noqa: BLE001
            e2e_graphs = []
            torch.cuda.synchronize()
            _log(rank, "e2e graph capture failed (%s); eager e2e" % str(e).splitlines()[0])
    if dist_on:
        ok = torch.tensor([1 if e2e_graphs else 0], device=dev)
        torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN)
        if int(ok.item()) == 0:
            e2e_graphs = []

    def e2e_step(i):
        k = i % NBUF
        comp.wait_event(ready[k])
        if e2e_graphs:
            e2e_graphs[k].replay()
        else:
            e2e_body(k)
        free[k].record(comp)

    def e2e_run(n):
        for k in range(NBUF):
            free[k].record(comp)
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            e2e_step(i)
        comp.wait_stream(copy_s)

    e2e_steps = max(3, min(args.steps, 100))
    e2e_run(4)
    barrier()
    e0.record()
    e2e_run(e2e_steps)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([e2e_ms], device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    _log(rank, "e2e timing done")
    e2e_value = N_ROWS * world * e2e_steps / (e2e_ms * 1e-3)
    h2d = 4 * B * S * D
    d2h = 8 * B * S + 4 * (n_grad + K)

    if rank == 0:
        fwd_ms, bwd_ms, kf, kb_ = _time_kernels(V, m, sets)
        peak, peak_src = _peaks()
        fwd_bytes = N_ROWS * (8 * D + 8 + 4 * K)                      # read x, write new_latent, idx(int64), p_code
        bwd_bytes = N_ROWS * (12 * D + 8 * K + 8)                     # read x, g_q, p_code, g_p, idx; write dx
        dom = (kb_, bwd_ms, bwd_bytes) if bwd_ms >= fwd_ms else (kf, fwd_ms, fwd_bytes)
        achieved = dom[2] / (dom[1] * 1e-3) / 1e9
        line = {"metric": "vq_fwd_bwd_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": _config(world),
                "launch_mode": "cuda_graph" if use_graph else "eager",
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                        "launch_mode": "cuda_graph per upload buffer" if e2e_graphs else "eager",
                        "numa_binding_rank0": numa,
                        "note": "enc_embs uploaded from pinned host memory every step; the upstream gradients are device-resident "
                                "(the downstream losses produce them on the device); indices, parameter gradients and the usage "
                                "histogram read back every step"},
                "gpu_launches": launches_per_step * args.steps,
                "roofline": {"bound": "hbm", "kernel": dom[0], "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": _ncu_traffic(dom[0]), "peak_source": peak_src,
                             "kernel_ms": {kf: fwd_ms, kb_: bwd_ms},
                             "frac_per_kernel": {kf: fwd_bytes / (fwd_ms * 1e-3) / 1e9 / peak, kb_: bwd_bytes / (bwd_ms * 1e-3) / 1e9 / peak},
                             "step_frac": (fwd_bytes + bwd_bytes) / (ms_total / args.steps * 1e-3) / 1e9 / peak if world == 1 else None,
                             "note": "kernel_ms = CUDA-event time of the named kernel alone: the library records the two events on "
                                     "the launch stream immediately around that launch (vqb_debug_set_kernel_events), every launch its "
                                     "own pair; forward and backward launches alternate back to back without host "
                                     "synchronisation, as in the timed region; averaged over ring-rotated inputs (> L2)",
                             "algorithmic_bytes": {"fwd": fwd_bytes, "bwd": bwd_bytes}},
                "clocks": clocks}
        if world == 1 and not args.no_sweep:
            # the same two kernels where launch latency and the single wave no longer dominate: 2^20 rows (20 x config 2)
            big = [[torch.randn(1024, 1024, D, device=dev), torch.randn(1024, 1024, K, device=dev),
                    torch.randn(1024, 1024, D, device=dev)] for _ in range(2)]
            f1, b1, _, _ = _time_kernels(V, m, big, iters=8)
            del big
            n1 = 1 << 20
            line["roofline"]["at_2^20_rows"] = {
                "kernel_ms": {kf: f1, kb_: b1},
                "frac_per_kernel": {kf: n1 * (8 * D + 8 + 4 * K) / (f1 * 1e-3) / 1e9 / peak,
                                    kb_: n1 * (12 * D + 8 * K + 8) / (b1 * 1e-3) / 1e9 / peak},
                "note": "same kernels, same algorithmic bytes per row, 2^20 rows per launch (inputs and outputs far beyond L2)"}
        if world == 1:
            # the CPU arm beside it: measured once, at N = 1 only (at N > 1 the other ranks would spin in a barrier meanwhile)
            cpu_rate, cpu_ms, cpu_done, cores, kind, what = cpu_fwd_bwd_rate(400, 3, budget_s=12.0)
            line["cpu_baseline"] = {"value": cpu_rate, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "sample": "%d full steps of the same workload on the host (%s)" % (cpu_done, what)}
            if not args.no_sweep:
                line["sweep"] = _c3_sweep(local_rank)
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line))
    faulthandler.cancel_dump_traceback_later()
    sys.stdout.flush()
    if dist_on:
        _log(rank, "final barrier")
        torch.distributed.barrier()
        torch.cuda.synchronize()
        # CUDA graphs that captured NCCL work keep the communicator referenced; tearing it down from Python has been
        # seen to wedge, so leave the process without running destructors (all results are already flushed)
        sys.stderr.flush()
        os._exit(0)


def run_encode(args, rank, world, local_rank):
    """BASELINE configs[4]: the encode path -- forward-only nearest-codeword search + gather under torch.no_grad() (what
    VQVAE.speech_to_text does at src/vqvae.py:119 when the caller is bin/train_vqvae.py:343-346 / bin/gen_specgram.py) and the
    text-side lookup inference(txt) (src/vqvae.py:147, bin/gen_specgram.py:95-108) -- over 10 000 synthetic utterances of 400
    encoder frames (800 mel frames / time_reduce_factor 2), utterances sharded over the GPUs (dist.shard_bounds), batches of
    64 utterances issued EAGERLY through the nn.Module exactly as the solver's loop would.  No collective on the data path;
    the usage histogram is summed once at the end.  value = frames/s over all GPUs (device time, max over ranks)."""
    import semi_tts_b200 as V
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    utts, frames, text_len, batch, ring_n = 10000, 400, 67, 64, 24      # 24 x 64 x 400 x 64 x 4 B = 157 MB > 126 MB L2
    torch.manual_seed(0)
    m = V.L2Embedding(K, False, **_codebook_kwargs()).to(dev).eval()
    lo, hi = V.dist.shard_bounds(utts, rank, world)
    n_batches = (hi - lo + batch - 1) // batch
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ring = [torch.randn(batch, frames, D, device=dev, generator=g) for _ in range(ring_n)]
    txt = torch.randint(3, K - 1, (batch, text_len), device=dev, generator=g)
    lib = V._lib.load()

    def run(fused):
        m.fused_search = fused
        m.usage.reset()
        done = 0
        with torch.no_grad():
            for b in range(n_batches):
                nb = min(batch, hi - lo - done)
                x, t = ring[b % ring_n], txt
                if nb < batch:                                   # the shard's last, ragged batch
                    x, t = x[:nb], t[:nb]
                m(x)
                m.inference(t)
                done += nb
        return done

    out = {}
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches = 0
    for name, fused in (("parity_mode", False), ("fused_search", True)):
        run(fused)                                              # warm-up pass
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.vqb_launch_count()
        t0 = time.perf_counter()
        e0.record()
        done = run(fused)
        e1.record()
        host_s = time.perf_counter() - t0                       # time to ISSUE the work (no sync inside the loop)
        torch.cuda.synchronize()
        launches = int(lib.vqb_launch_count() - l0)
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        ms = float(ms.item())
        out[name] = {"ms": ms, "utts_per_s": utts / (ms * 1e-3), "frames_per_s": utts * frames / (ms * 1e-3),
                     "host_issue_us_per_batch": host_s / n_batches * 1e6, "device_us_per_batch": ms * 1e3 / n_batches}
        assert done == hi - lo
    if world > 1:
        V.dist.allreduce_usage(m)
    total = m.usage.total()
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        best = out["parity_mode"]
        print(json.dumps({"metric": "vq_encode_frames_per_sec", "value": best["frames_per_s"], "unit": "frames/s", "n_gpus": world,
                          "steps": n_batches, "warmup": n_batches, "ms_per_step": best["ms"] / n_batches,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": "configs[4]: no-grad search + gather and inference(txt), %d utterances x %d frames, "
                                                 "K=%d D=%d, batches of %d utterances issued eagerly through the nn.Module" % (
                                                     utts, frames, K, D, batch),
                                     "l2_policy": "ring of %d input batches (157 MB) > 126 MB L2" % ring_n,
                                     "parallelism": "dp%d (utterances sharded, no data-path collective)" % world},
                          "gpu_launches": launches, "usage_total": total, "usage_expected": utts * frames,
                          "parity_mode": out["parity_mode"], "fused_search": out["fused_search"], "clocks": clocks}))
    sys.stdout.flush()
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-3 roofline sweep sub-record (N = 1 only)")
    ap.add_argument("--workload", default="train", choices=["train", "encode"],
                    help="train: BASELINE configs[1] fwd+bwd (the bench line); encode: configs[4], no-grad search + gather")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL prints its version banner) must not pollute stdout: the contract is ONE JSON line there
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    if args.impl == "reference":
        run_reference(args, rank)
        sys.stdout.flush()
        return
    if args.workload == "encode":
        run_encode(args, rank, world, local_rank)
    else:
        run_ours(args, rank, world, local_rank)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
