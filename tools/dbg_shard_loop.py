#!/usr/bin/env python
"""Where the plane-sharded head's time goes in a bench-style loop (dev tool): device vs host time per call,
one vs two rotating input sets, gather vs scatter exchange.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/dbg_shard_loop.py"""
import importlib, os, sys, time
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
out = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dpv = importlib.import_module("probabilistic-depth_b200")
sh = importlib.import_module("probabilistic-depth_b200.sharding")
D, H, W = 256, 384, 1280
d = dpv.synth.depth_candidates(5, 40, D)
lo, hi = sh.plane_range(D, rank, world)
xs = [torch.randn((1, hi - lo, H, W), device="cuda") * 3 for _ in range(2)]
for exchange in ("gather", "scatter"):
    head = sh.PlaneShardedHead(D, exchange=exchange)
    for nset in (1, 2):
        for _ in range(5):
            head(xs[0], d)
        torch.cuda.synchronize(); dist.barrier()
        n = 50
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for i in range(n):
            head(xs[i % nset], d)
        b.record()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        if rank == 0:
            print("world %d %-7s nset %d: device %.3f ms/call, host issue %.3f ms/call" %
                  (world, exchange, nset, a.elapsed_time(b) / n, 1e3 * (t1 - t0) / n), file=out, flush=True)
dist.destroy_process_group()
