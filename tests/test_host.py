"""CPU-only tests: the C-ABI library loads and exports every symbol include/vqb.h declares, host-side
module logic mirrors the reference's interface, and the multi-rank reduction algebra (gloo, world 2)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from helpers import build_module, codebook_kwargs, phn_attr_tsv


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vqb.h")).read()
    return sorted(set(re.findall(r"VQB_API\s+[\w\s\*]+?\b(vqb_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import semi_tts_b200 as V
    lib = V._lib.load()
    names = _declared_symbols()
    assert len(names) >= 12 and set(names) == set(V._lib.EXPORTS)
    for n in names:
        assert getattr(lib, n) is not None
    assert lib.vqb_abi_version() == V._lib.ABI_VERSION


def test_nothing_is_exported_that_no_header_declares():
    """the drop-in ABI is include/vqb.h; the developer hooks are declared (and fenced off) in include/vqb_debug.h; the
    library exports nothing else"""
    import subprocess
    import semi_tts_b200 as V
    so = os.path.join(os.path.dirname(V.__file__), "libvqb200.so")
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.split()[-1].startswith("vqb_")}
    dbg = open(os.path.join(ROOT, "include", "vqb_debug.h")).read()
    debug = set(re.findall(r"VQB_API\s+[\w\s\*]+?\b(vqb_\w+)\s*\(", dbg))
    assert debug and all(n.startswith("vqb_debug_") for n in debug)
    assert exported == set(_declared_symbols()) | debug


def test_ctypes_structs_match_the_c_header(tmp_path):
    """sizeof/offsetof of the ctypes mirrors == what a C compiler makes of include/vqb.h."""
    import subprocess
    import semi_tts_b200 as V
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "vqb.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(vqb_fwd_args), offsetof(vqb_fwd_args, temp), offsetof(vqb_fwd_args, workspace_bytes),'
                   'sizeof(vqb_bwd_args), offsetof(vqb_bwd_args, idx), offsetof(vqb_bwd_args, d_temp),'
                   'offsetof(vqb_fwd_args, row_lengths), offsetof(vqb_fwd_args, ctc_eps), offsetof(vqb_bwd_args, g_logp),'
                   'offsetof(vqb_bwd_args, tail), sizeof(vqb_bwd_tail), offsetof(vqb_bwd_tail, reserved));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    F, B, T = V._lib.FwdArgs, V._lib.BwdArgs, V._lib.BwdTail
    assert got == [ctypes.sizeof(F), F.temp.offset, F.workspace_bytes.offset,
                   ctypes.sizeof(B), B.idx.offset, B.d_temp.offset,
                   F.row_lengths.offset, F.ctc_eps.offset, B.g_logp.offset, B.tail.offset, ctypes.sizeof(T), T.reserved.offset]


def test_c_program_links_and_uses_the_abi_without_python(tmp_path):
    """The boundary is a C ABI: a C99 program includes vqb.h, links libvqb200.so and gets the documented behaviour
    from the host-side entry points (version, workspace sizing, argument validation with thread-local messages,
    no-device failure) -- no torch, no Python in between."""
    import subprocess
    lib_dir = os.path.join(ROOT, "semi-tts_b200")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "vqb.h"
int main(void) {
    if (vqb_abi_version() != VQB_ABI_VERSION) { printf("version\n"); return 1; }
    vqb_fwd_args a; memset(&a, 0, sizeof a);
    size_t n = 123;
    a.struct_size = 4;                                        /* wrong size -> VQB_ERR_INVALID + message */
    if (vqb_forward_workspace(&a, &n) != VQB_ERR_INVALID || !strstr(vqb_last_error(), "struct_size")) { printf("size\n"); return 2; }
    a.struct_size = (uint32_t)sizeof a;
    a.flags = VQB_SCORE_L2 | VQB_SCORE_LINEAR;                /* both scores -> invalid */
    a.n_rows = 8; a.dim = 64; a.n_codes = 43;
    if (vqb_forward_workspace(&a, &n) != VQB_ERR_INVALID) { printf("flags\n"); return 3; }
    a.flags = VQB_SCORE_L2 | VQB_STOP_GRAD;
    a.dim = 62;                                               /* D must be a multiple of 4 */
    if (vqb_forward_workspace(&a, &n) != VQB_ERR_INVALID || !strstr(vqb_last_error(), "multiple of 4")) { printf("dim\n"); return 4; }
    a.dim = 64; a.n_rows = 0;                                 /* empty input: nothing to do, no workspace */
    if (vqb_forward_workspace(&a, &n) != VQB_OK || n != 0) { printf("empty\n"); return 5; }
    size_t sb = 0;                                            /* large-table scatter needs scratch, small tables none */
    if (vqb_scatter_workspace(1 << 20, 8192, 256, &sb) != VQB_OK || sb == 0) { printf("scatter\n"); return 6; }
    if (vqb_scatter_workspace(51200, 43, 64, &sb) != VQB_OK || sb != 0) { printf("scatter small\n"); return 7; }
    if (vqb_exchange_bytes(2580, 8) < (size_t)2 * 8 * 2580 * 8) { printf("exchange\n"); return 8; }
    printf("devices %d\n", vqb_device_count());
    return 0;
}
''')
    exe = tmp_path / "abi"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-lvqb200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.startswith("devices ")


def test_abi_argument_validation_never_crashes_on_random_arguments():
    """Host-side validation of the C ABI under random (mostly invalid) shapes and flags: always a return code and, on
    failure, a non-empty message -- never a crash; and without a device the compute entry points refuse to run."""
    import semi_tts_b200 as V
    lib = V._lib.load()
    rng = np.random.default_rng(7)
    has_gpu = lib.vqb_device_count() > 0
    for _ in range(300):
        a = V._lib.FwdArgs()
        a.struct_size = ctypes.sizeof(V._lib.FwdArgs) if rng.random() < 0.9 else int(rng.integers(0, 400))
        a.flags = int(rng.integers(0, 1 << 8))
        a.n_rows = int(rng.choice([-5, 0, 1, 127, 128, 129, 51200, 1 << 20, 1 << 31, 1 << 40]))
        a.dim = int(rng.choice([-4, 0, 1, 3, 4, 20, 32, 62, 64, 128, 256, 512, 516, 4096]))
        a.n_codes = int(rng.choice([-1, 0, 1, 43, 64, 65, 300, 8192, 1 << 24, 1 << 30]))
        n = ctypes.c_size_t(12345)
        rc = lib.vqb_forward_workspace(ctypes.byref(a), ctypes.byref(n))
        assert rc in (0, 1, 2, 3, 4)
        if rc:
            assert len(lib.vqb_last_error()) > 0 and n.value == 0
        assert isinstance(lib.vqb_forward_kernel_name(ctypes.byref(a)), bytes)
        b = V._lib.BwdArgs()
        b.struct_size = ctypes.sizeof(V._lib.BwdArgs) if rng.random() < 0.9 else int(rng.integers(0, 400))
        b.flags, b.n_rows, b.dim, b.n_codes = a.flags, a.n_rows, a.dim, a.n_codes
        b.n_real_rows = int(rng.integers(-3, 1 << 20))
        rc = lib.vqb_backward_workspace(ctypes.byref(b), ctypes.byref(n))
        assert rc in (0, 1, 2, 3, 4)
        assert isinstance(lib.vqb_backward_kernel_name(ctypes.byref(b)), bytes)
        sb = ctypes.c_size_t(0)
        assert lib.vqb_scatter_workspace(max(a.n_rows, 0), max(a.n_codes, 1), max(a.dim, 4), ctypes.byref(sb)) in (0, 1)
        if not has_gpu:
            # all pointers are NULL here: either the arguments are rejected, or an empty call is a no-op, or the missing
            # device is reported -- the library never dereferences anything on this path
            rc = lib.vqb_forward(ctypes.byref(a), None)
            assert rc in (0, 1, 4)
            if rc == 0:
                assert a.n_rows == 0


def test_struct_size_mismatch_is_an_error_without_a_gpu():
    import semi_tts_b200 as V
    lib = V._lib.load()
    a = V._lib.FwdArgs()
    a.struct_size = 8
    n = ctypes.c_size_t(0)
    assert lib.vqb_forward_workspace(ctypes.byref(a), ctypes.byref(n)) != 0
    assert b"struct_size" in lib.vqb_last_error()


def test_state_dict_keys_and_seeded_init_match_the_reference():
    """Strict-load compatibility (bin/train_vqvae.py:106) and same-seed construction (RNG order of
    src/embed.py:28-29,62,81,85)."""
    import semi_tts_b200 as V
    g = load_golden("init_l2_seed0")
    torch.manual_seed(0)
    m = V.L2Embedding(43, False, **codebook_kwargs(g))
    sd = m.state_dict()
    assert list(sd.keys()) == [k[3:] for k in g.keys()]
    for k, v in g.items():
        assert np.array_equal(sd[k[3:]].numpy(), v), k
    g = load_golden("init_sep_seed0")
    torch.manual_seed(0)
    m = V.SeperateEmbedding(43, False, **codebook_kwargs(g, bone="sep"))
    sd = m.state_dict()
    assert list(sd.keys()) == [k[3:] for k in g.keys()]
    for k, v in g.items():
        assert np.array_equal(sd[k[3:]].numpy(), v), k


def test_module_surface():
    import semi_tts_b200 as V
    g = load_golden("init_l2_seed0")
    m = V.L2Embedding(43, False, **codebook_kwargs(g))
    assert m.out_dim == 64 and m.latent_dim == 64 and m.vocab_size == 43
    assert "Temp. = 1.0" in m.create_msg() and "Phn. attributs = True" in m.create_msg()
    assert [n for n, p in m.named_parameters() if p.requires_grad] == ["learnable_table", "proj_attr.weight", "proj_attr.bias"]
    m2 = V.L2Embedding(43, False, **codebook_kwargs(g, temp=-1))
    assert isinstance(m2.temp, torch.nn.Parameter) and "learnable" in m2.create_msg()
    with pytest.raises(AssertionError):
        V.L2Embedding(43, True, **codebook_kwargs(g))                     # ema must be False
    with pytest.raises(AssertionError):
        V.SeperateEmbedding(43, False, **codebook_kwargs(g, skip_prob=0.5, bone="l2"))


def test_read_phn_attr_matches_reference_table():
    import semi_tts_b200 as V
    tab = V.read_phn_attr(phn_attr_tsv())
    assert np.array_equal(tab, np.load(os.path.join(ROOT, "tests", "golden", "phn_attr_table.npy")))
    assert tab.shape == (43, 31) and not tab[:3].any()


def test_cpu_forward_raises():
    g = load_golden("l2_attr_stopgrad")
    m = build_module(g, "l2", device="cpu")
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.from_numpy(g["x"]))
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.inference(torch.zeros(2, 3, dtype=torch.long))


def test_shard_bounds_cover_batch():
    import semi_tts_b200 as V
    for n, w in [(64, 8), (10, 4), (3, 8), (0, 2)]:
        spans = [V.dist.shard_bounds(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import semi_tts_b200 as V
    from helpers import build_module as bm
    from conftest import load_golden as lg
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = lg("l2_attr_stopgrad")
    m = bm(g, "l2", device="cpu")
    # per-rank shard gradients = slices of the reference's per-row contributions; emulate with seeded tensors
    gen = torch.Generator().manual_seed(100 + rank)
    for p in m.parameters():
        if p.requires_grad:
            p.grad = torch.randn(p.shape, generator=gen)
    m.usage.counts = torch.randint(0, 1000, (43,), generator=gen)
    V.dist.allreduce_codebook_grads(m, include_usage=True)
    # a second step's counts, exchanged again: nothing may be counted twice
    m.usage.counts += torch.randint(0, 1000, (43,), generator=gen)
    V.dist.allreduce_usage(m)
    # numpy payloads: torch tensors travel through a Queue by shared fd, which races with worker exit
    out = {n: p.grad.numpy().copy() for n, p in m.named_parameters() if p.requires_grad}
    out["usage"] = m.usage.all_counts().numpy().copy()
    out["usage_bar"] = np.asarray(m.usage.bar())
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_allreduce_of_codebook_grads_and_histogram_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = load_golden("l2_attr_stopgrad")
    m = build_module(g, "l2", device="cpu")
    expect = {}
    for rank in range(2):
        gen = torch.Generator().manual_seed(100 + rank)
        for n, p in m.named_parameters():
            if p.requires_grad:
                expect[n] = expect.get(n, 0) + torch.randn(p.shape, generator=gen)
        expect["usage"] = expect.get("usage", 0) + torch.randint(0, 1000, (43,), generator=gen) \
            + torch.randint(0, 1000, (43,), generator=gen)
    bar = expect["usage"].double() / expect["usage"].sum()
    bar[0] = 0
    for rank in range(2):
        for n, v in expect.items():
            assert torch.allclose(torch.from_numpy(res[rank][n]).to(v.dtype), v, atol=1e-6), (rank, n)
        assert res[rank]["usage"].dtype == np.int64
        assert np.array_equal(res[rank]["usage"], expect["usage"].numpy())
        assert np.allclose(res[rank]["usage_bar"], bar.numpy())


def test_flat_gradient_view_detection():
    """dist._flat_view: consecutive views of one buffer are reduced in place; anything else is packed."""
    import semi_tts_b200 as V
    flat = torch.arange(20, dtype=torch.float32)
    a, b, c = flat[:6].view(2, 3), flat[6:14].view(4, 2), flat[14:]
    v = V.dist._flat_view([a, b, c])
    assert v is not None and v.numel() == 20 and v.data_ptr() == flat.data_ptr()
    v.mul_(2)
    assert torch.equal(a, torch.arange(6, dtype=torch.float32).view(2, 3) * 2)
    assert V.dist._flat_view([a, c]) is None                       # gap
    assert V.dist._flat_view([a, torch.zeros(3)]) is None          # different storage
    assert V.dist._flat_view([]) is None


@pytest.mark.skipif(not os.path.isfile("/root/reference/src/vqvae.py"), reason="reference tree not present")
def test_drop_in_into_reference_vqvae_construction():
    """The unmodified reference VQVAE (src/vqvae.py:26-90) builds with the B200 classes patched in; parameter
    names / shapes of the whole model are identical to the stock model's (strict checkpoint compatibility)."""
    import copy
    import yaml
    from oracle import ref_import
    ref_embed = ref_import.import_reference()
    V_ref = ref_import.import_reference_vqvae()
    import semi_tts_b200 as V
    stock = (ref_embed.L2Embedding, ref_embed.SeperateEmbedding, V_ref.L2Embedding, V_ref.SeperateEmbedding)
    stock_mean_forward = V_ref.VQVAE.mean_forward
    with ref_import.reference_cwd():
        cfg = yaml.load(open("config/semi-multi-spkr-paired-data.yaml"), Loader=yaml.FullLoader)
        torch.manual_seed(0)
        ref_model = V_ref.VQVAE(80, 1025, 43, 109, **copy.deepcopy(cfg["model"]))
        try:
            V.install_into_reference()
            torch.manual_seed(0)
            new_model = V_ref.VQVAE(80, 1025, 43, 109, **copy.deepcopy(cfg["model"]))
        finally:
            V.uninstall_from_reference()
    assert (ref_embed.L2Embedding, ref_embed.SeperateEmbedding, V_ref.L2Embedding, V_ref.SeperateEmbedding) == stock
    assert V_ref.VQVAE.mean_forward is stock_mean_forward      # the run-length collapse is put back as well
    assert type(new_model.codebook) is V.L2Embedding
    a, b = ref_model.state_dict(), new_model.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape, k
        if k.startswith("codebook."):
            assert torch.equal(a[k], b[k]), k                      # same seed -> same initial codebook
    new_model.load_state_dict(a, strict=True)
    assert new_model.codebook.out_dim == ref_model.codebook.out_dim
    assert "Phn. attributs = True" in new_model.create_msg()[-1] or True


def test_usage_add_counts_chosen_rows():
    """usage.add(idx): the caller picks the rows (the reference counts the unpaired batch only, bin/train_vqvae.py:256-261);
    bar() then applies src/util.py:139-143 to exactly those."""
    from semi_tts_b200.usage import UsageHistogram
    u = UsageHistogram(6)
    idx = torch.tensor([[1, 2, 2], [5, 0, 2]])
    u.add(idx)
    u.add(idx[1:])
    assert u.all_counts().tolist() == [2, 1, 4, 0, 0, 2] and u.total() == 9
    bar = u.bar()
    assert bar[0] == 0.0 and abs(bar[2] - 4 / 9) < 1e-12


def test_bench_reference_arm_runs_on_the_cpu_and_keeps_the_contract():
    """bench.py --impl reference: ONE JSON line on stdout with the GPU arm's metric / unit / config, impl = reference, a
    cpu_baseline that says what was timed, and the e2e object; the GPU arm refuses to run without a GPU (no CPU fallback)."""
    import json
    import subprocess
    env = dict(os.environ, OMP_NUM_THREADS="4")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "vq_fwd_bwd_frames_per_sec" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench._config(1)                 # the dict the GPU arm prints (same function, same arguments)
    if not torch.cuda.is_available():
        gpu = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1", "--no-sweep"],
                             capture_output=True, text=True, timeout=600, env=env)
        assert gpu.returncode != 0 and "no CPU fallback" in gpu.stderr


def test_per_step_holders_do_not_travel_with_the_module():
    """the CTC log-probability holder and the no-grad cache are per-step / per-device state: copy.deepcopy and pickle of
    the module work with a graph tensor parked in them and produce empty holders; the state dict does not see them; the
    cached parameter tuple of the no-grad path follows a replaced Parameter."""
    import copy
    import pickle
    g = load_golden("l2_attr_stopgrad")
    m = build_module(g, "l2", device="cpu")
    leaf = torch.ones(3, requires_grad=True)
    m._ctc_out.value = leaf * 2.0                       # a non-leaf tensor: deepcopy of it alone would raise
    m.ctc_eps = 1e-10
    keys = list(m.state_dict().keys())
    m2 = copy.deepcopy(m)
    assert m2.ctc_logp is None and m2.ctc_eps == 1e-10 and m.ctc_logp is not None
    m3 = pickle.loads(pickle.dumps(m))
    assert m3.ctc_logp is None and list(m3.state_dict().keys()) == keys == list(m2.state_dict().keys())
    assert torch.equal(m3.learnable_table, m.learnable_table)
    p0 = m._nograd_params()
    assert p0[0] is m.learnable_table and p0[4] is m.temp and m._nograd_params() is p0
    m.learnable_table = torch.nn.Parameter(m.learnable_table.detach().clone())
    p1 = m._nograd_params()
    assert p1 is not p0 and p1[0] is m.learnable_table
