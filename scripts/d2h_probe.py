#!/usr/bin/env python
"""is the N-GPU e2e leg bound by the host?  Every rank copies a 157 MB device buffer
(the observations of 1M envs) to its own pinned host buffer, all ranks at once, nothing else running:
aggregate D2H GB/s of plain cudaMemcpyAsync at N ranks.  torchrun --nproc-per-node N scripts/d2h_probe.py"""
import json, os, time
import torch
import torch.distributed as dist

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        del os.environ["NCCL_DEBUG"]
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 157286400 // 4
dev = torch.empty(n, device="cuda")
host = torch.empty(n, pin_memory=True)
up = torch.empty(25165824 // 4, pin_memory=True); upd = torch.empty_like(up, device="cuda")
res = {}
for name, fn in (("d2h", lambda: host.copy_(dev, non_blocking=True)),
                 ("h2d", lambda: upd.copy_(up, non_blocking=True)),
                 ("both", lambda: (host.copy_(dev, non_blocking=True), upd.copy_(up, non_blocking=True)))):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    K = 20
    for _ in range(K):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    nbytes = {"d2h": host.numel() * 4, "h2d": up.numel() * 4, "both": host.numel() * 4 + up.numel() * 4}[name]
    res[name] = {"per_rank_GBs": nbytes * K / float(t) / 1e9, "aggregate_GBs": world * nbytes * K / float(t) / 1e9}
if rank == 0:
    try:
        aff = sorted(os.sched_getaffinity(0))
    except Exception:
        aff = []
    print(json.dumps({"probe": "concurrent pinned copies, no kernels", "ranks": world, "result": res,
                      "cpus_visible": len(aff), "numa_nodes": len([d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')]) if os.path.isdir('/sys/devices/system/node') else None}))
if world > 1:
    dist.destroy_process_group()
