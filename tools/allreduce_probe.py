"""Latency of the step's one collective: all-reduce (sum) of the packed gradient buffer (56 B x P) -- NCCL and
libghr's NVLink peer-memory kernel -- launched eagerly and replayed from a CUDA graph, with the per-call
distribution (CUDA events around every call).
torchrun --nproc-per-node N tools/allreduce_probe.py [--P 60000] [--iters 300]"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guassianhand_b200.dist import PackedGrads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=60000)
ap.add_argument("--iters", type=int, default=300)
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
out = {"world": world, "bytes": a.P * 14 * 4}


def stats(ms):
    ms = np.asarray(ms) * 1000.0
    return {"median_us": round(float(np.median(ms)), 1), "p90_us": round(float(np.percentile(ms, 90)), 1),
            "p99_us": round(float(np.percentile(ms, 99)), 1), "max_us": round(float(ms.max()), 1)}


def measure(fn, n):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for s, e in ev:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return stats([s.elapsed_time(e) for s, e in ev])


for name, peer in (("nccl", False), ("peer", True)):
    g = PackedGrads(a.P, 0, device=dev, peer=peer)
    g.flat.fill_(1.0)
    out[name + "_eager"] = measure(lambda: g.all_reduce_(), a.iters)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g.all_reduce_()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        g.all_reduce_()
    out[name + "_graph"] = measure(lambda: graph.replay(), a.iters)
    if peer:
        out["peer_status"] = g.comm.status()
allr = [None] * world
dist.all_gather_object(allr, out)
if rank == 0:
    print(json.dumps({"per_rank": allr}), flush=True)
torch.cuda.synchronize()
os._exit(0)
