"""N-GPU check of the fused backward tail (one-shot all-reduce over NVLink peer memory) against the NCCL route.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/dist_check.py
Every rank runs the module on its own shard; the gradients of the fused route must equal (to fp32 round-off of a
different summation order across ranks) the NCCL all-reduce of the unfused route, be bit-identical on all ranks, and
stay so across repeated calls and CUDA-graph replays.  Prints one JSON line on rank 0; exit code 1 on mismatch."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    import semi_tts_b200 as V
    from helpers import phn_attr_tsv
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    B, S, K, D = 16, 200, 43, 64
    torch.manual_seed(0)
    m = V.L2Embedding(K, False, softmax="normal", latent_dim=D, commit_weight=0, vq_weight=0, temp=1, skip_prob=0,
                      stop_grad=True, phn_attr_pth=phn_attr_tsv(), proj_attr=16).to(dev)
    m.train()
    g = torch.Generator().manual_seed(100 + rank)
    sets = [[torch.randn(B, S, D, generator=g).to(dev).requires_grad_(True), torch.randn(B, S, K, generator=g).to(dev),
             torch.randn(B, S, D, generator=g).to(dev)] for _ in range(3)]

    txt = torch.randint(0, K, (B, 37), generator=g).to(dev)
    g_inf = torch.randn(B, 37, D, generator=g).to(dev)
    with_lookup = [False]

    def step(s):
        for p_ in m.parameters():
            p_.grad = None
        s[0].grad = None
        p, q, _, _ = m(s[0])
        outs, grads = [p, q], [s[1], s[2]]
        if with_lookup[0]:
            # the text branch trains the codebook through inference() in every reference step (src/vqvae.py:147): a second,
            # non-fused contribution to the same parameter gradients
            outs.append(m.inference(txt)); grads.append(g_inf)
        torch.autograd.backward(outs, grads)
        V.dist.allreduce_codebook_grads(m)
        return torch.cat([p_.grad.reshape(-1) for p_ in m.parameters() if p_.requires_grad]).clone(), s[0].grad.clone()

    # reference route: unfused backward + NCCL all-reduce
    m.fused_tail.enabled = False
    ref = [step(s) for s in sets]
    assert not m.fused_tail.fused
    with_lookup[0] = True
    ref_lookup = [step(s) for s in sets]
    with_lookup[0] = False
    # fused route
    m.fused_tail.enabled = True
    V.dist.enable_fused_allreduce(m)
    ok, worst, worst_dx = True, 0.0, 0.0
    for rep in range(3):                       # repeated calls: both exchange slots, growing epochs
        for s, (rg, rdx) in zip(sets, ref):
            got, dx = step(s)
            assert m.fused_tail.fused
            err = float((got - rg).norm() / rg.norm())
            worst, worst_dx = max(worst, err), max(worst_dx, float((dx - rdx).abs().max()))
            gathered = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(gathered, got)
            same = all(torch.equal(gathered[0], t) for t in gathered)
            ok = ok and err < 2e-6 and same
    # mixed routes in one step: fused tail (forward path) + inference() lookups (NCCL inside their backward)
    with_lookup[0] = True
    worst_mixed = 0.0
    for s, (rg, rdx) in zip(sets, ref_lookup):
        got, dx = step(s)
        err = float((got - rg).norm() / rg.norm())
        worst_mixed = max(worst_mixed, err)
        gathered = [torch.empty_like(got) for _ in range(world)]
        dist.all_gather(gathered, got)
        ok = ok and err < 2e-6 and all(torch.equal(gathered[0], t) for t in gathered)
    with_lookup[0] = False
    # deferred exchange: the tail only pushes, allreduce_codebook_grads() (-> finish) completes the sum later
    m.fused_tail.defer = True
    worst_defer = 0.0
    for rep in range(2):
        for s, (rg, rdx) in zip(sets, ref):
            got, dx = step(s)
            err = float((got - rg).norm() / rg.norm())
            worst_defer = max(worst_defer, err)
            gathered = [torch.empty_like(got) for _ in range(world)]
            dist.all_gather(gathered, got)
            ok = ok and err < 2e-6 and all(torch.equal(gathered[0], t) for t in gathered)
    m.fused_tail.defer = False
    V.dist.check_exchange(m)
    # an empty shard on the last rank: it still joins the exchange (its tail runs over zero rows); expected = the sum of the
    # other ranks' local gradients
    ex = m.fused_tail.exchange
    m.fused_tail.exchange = None

    def local_grads(s):
        for p_ in m.parameters():
            p_.grad = None
        if s[0].shape[0]:
            p, q, _, _ = m(s[0])
            torch.autograd.backward([p, q], [s[1], s[2]])
            return torch.cat([p_.grad.reshape(-1) for p_ in m.parameters() if p_.requires_grad]).clone()
        return torch.zeros(sum(p_.numel() for p_ in m.parameters() if p_.requires_grad), device=dev)

    last = rank == world - 1
    s_e = [t[:0] if last else t for t in sets[1]]
    s_e[0] = s_e[0].detach().requires_grad_(True)
    loc = local_grads(s_e)
    gathered = [torch.empty_like(loc) for _ in range(world)]
    dist.all_gather(gathered, loc)
    want = torch.stack(gathered).sum(0)
    m.fused_tail.exchange = ex
    got, _ = step(s_e)
    err_empty = float((got - want).norm() / want.norm())
    ok = ok and err_empty < 2e-6
    V.dist.check_exchange(m)
    # CUDA-graph replay of the fused step
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(sets[0])
    torch.cuda.current_stream().wait_stream(side)
    for p_ in m.parameters():
        p_.grad = None
    sets[0][0].grad = None
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr, capture_error_mode="thread_local"):
        p, q, _, _ = m(sets[0][0])
        torch.autograd.backward([p, q], [sets[0][1], sets[0][2]])
    for _ in range(5):
        gr.replay()
    torch.cuda.synchronize()
    got = torch.cat([p_.grad.reshape(-1) for p_ in m.parameters() if p_.requires_grad])
    err_g = float((got - ref[0][0]).norm() / ref[0][0].norm())
    ok = ok and err_g < 2e-6
    # timing: graph replay of the fused step vs NCCL step
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        gr.replay()
    e1.record(); torch.cuda.synchronize()
    fused_us = e0.elapsed_time(e1) * 1e3 / 50
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"world": world, "ok": bool(flag.item()), "max_rel_err_vs_nccl": worst, "max_abs_dx_diff": worst_dx,
                          "graph_replay_rel_err": err_g, "mixed_routes_rel_err": worst_mixed, "empty_shard_rel_err": err_empty, "deferred_exchange_rel_err": worst_defer,
                          "fused_step_us_16x200": fused_us}), flush=True)
    dist.barrier(); torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
