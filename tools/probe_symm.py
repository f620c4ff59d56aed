"""2-GPU probe: does torch symmetric memory rendezvous work on this box, and are peer pointers usable?"""
import os, sys, time
import torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
t = symm.empty(4096, dtype=torch.float32, device=torch.device("cuda", lr))
t.fill_(float(rank + 1))
try:
    h = symm.rendezvous(t, dist.group.WORLD)
    print(rank, "rendezvous ok", [hex(p) for p in h.buffer_ptrs], "multicast", h.has_multicast_support, hex(h.multicast_ptr) if h.has_multicast_support else None,
          "signal pad", h.signal_pad_size, flush=True)
    h.barrier()
    peer = h.get_buffer((rank + 1) % world, (4096,), torch.float32)
    torch.cuda.synchronize()
    print(rank, "peer value", float(peer[0].item()), flush=True)
    h.barrier()
    # NCCL small all-reduce latency in a CUDA graph, for comparison
    x = torch.ones(2700, device="cuda")
    for _ in range(5): dist.all_reduce(x)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, capture_error_mode="thread_local"):
        for _ in range(20): dist.all_reduce(x)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dist.barrier(); torch.cuda.synchronize()
    e0.record(); 
    for _ in range(10): g.replay()
    e1.record(); torch.cuda.synchronize()
    print(rank, "nccl all_reduce 10.8KB in graph: %.2f us each" % (e0.elapsed_time(e1) * 1e3 / 200), flush=True)
except Exception as e:
    print(rank, "FAILED", repr(e), flush=True)
dist.barrier(); torch.cuda.synchronize()
os._exit(0)
