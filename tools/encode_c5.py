"""BASELINE.json configs[4]: the --gen-specgram encode path -- forward-only nearest-codeword search + gather under
torch.no_grad() (what VQVAE.forward does at src/vqvae.py:119 when the caller is bin/train_vqvae.py:343 /
bin/gen_specgram.py) and the text-side lookup `inference(txt)` (src/vqvae.py:147) -- over 10 000 synthetic utterances
of 400 encoder frames (800 mel frames / time_reduce_factor 2), utterances sharded over the GPUs by `dist.shard_bounds`.
No collective on the data path; the usage histogram is summed once at the end.

  python tools/encode_c5.py [--utts 10000] [--frames 400] [--batch 64]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tools/encode_c5.py

Prints one JSON line on rank 0: utterances/s and frames/s (device time, max over ranks), HBM-resident inputs (a ring of
batches larger than L2), plus the p_code-free fused search next to the parity-mode forward.  A measurement tool, not a
bench line (bench.py stays on configs[1]).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utts", type=int, default=10000)
    ap.add_argument("--frames", type=int, default=400)
    ap.add_argument("--text-len", type=int, default=67)        # FRAME_PHN_RATIO = 6 (src/vqvae.py:18)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--ring", type=int, default=24)            # 24 x 64 x 400 x 64 x 4 B = 157 MB > 126 MB L2
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import semi_tts_b200 as V

    K, D = 43, 64
    torch.manual_seed(0)
    m = V.L2Embedding(K, False, softmax="normal", latent_dim=D, commit_weight=0, vq_weight=0, temp=1, skip_prob=0,
                      stop_grad=True).cuda().eval()
    lo, hi = V.dist.shard_bounds(args.utts, rank, world)
    n_batches = (hi - lo + args.batch - 1) // args.batch
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    ring = [torch.randn(args.batch, args.frames, D, device="cuda", generator=g) for _ in range(args.ring)]
    txt = torch.randint(3, K - 1, (args.batch, args.text_len), device="cuda", generator=g)

    def run(fused):
        m.fused_search = fused
        m.usage.reset()
        done = 0
        with torch.no_grad():
            for b in range(n_batches):
                nb = min(args.batch, hi - lo - done)
                x = ring[b % args.ring][:nb]
                _, q, _, _ = m(x)
                _ = m.inference(txt[:nb])
                done += nb
        return done

    out = {}
    for name, fused in (("parity_mode", False), ("fused_search", True)):
        run(fused)                                              # warm-up pass
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        done = run(fused)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        out[name] = {"ms": float(ms.item()), "utts_per_s": args.utts / (float(ms.item()) * 1e-3),
                     "frames_per_s": args.utts * args.frames / (float(ms.item()) * 1e-3)}
        assert done == hi - lo
    if world > 1:
        V.dist.allreduce_usage(m)
    total = m.usage.total()
    if rank == 0:
        print(json.dumps({"workload": "configs[4]: no-grad search + gather and inference(txt), %d utterances x %d frames, "
                                      "K=%d D=%d" % (args.utts, args.frames, K, D),
                          "n_gpus": world, "batch": args.batch, "usage_total": total,
                          "usage_expected": args.utts * args.frames, **out}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
