"""How much host<->device bandwidth does this box give N ranks at once?  Every rank copies a page-locked 1 GiB buffer
device->host (and host->device) a few times, all ranks at the same time; prints per-rank and aggregate GB/s.
Run under torchrun.  (Context for the end-to-end numbers of bench.py at N > 1: 16 bytes per Pair record leave the GPUs.)"""
import os
import time

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
out = {}
for name, (dst, src) in (("d2h", (host, dev)), ("h2d", (dev, host))):
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    gbs = torch.tensor([4 * n / dt / 1e9], dtype=torch.float64, device="cuda")
    if world > 1:
        lo = gbs.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(gbs, op=dist.ReduceOp.SUM)
        out[name] = {"aggregate_gbs": float(gbs), "slowest_rank_gbs": float(lo)}
    else:
        out[name] = {"aggregate_gbs": float(gbs), "slowest_rank_gbs": float(gbs)}
if rank == 0:
    import json
    print(json.dumps({"ranks": world, "host_cores": os.cpu_count(), **out}))
if world > 1:
    dist.destroy_process_group()
