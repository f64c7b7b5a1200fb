"""Latency of the step's one collective: NCCL all-reduce (sum) of the packed gradient buffer (56 B x P),
eager and replayed from a CUDA graph.  torchrun --nproc-per-node N tools/allreduce_probe.py [--P 60000]"""
import argparse
import os

import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--P", type=int, default=60000)
ap.add_argument("--iters", type=int, default=200)
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
x = torch.ones(a.P * 14, device=dev)
for _ in range(10):
    dist.all_reduce(x)
torch.cuda.synchronize()
dist.barrier()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(a.iters):
    dist.all_reduce(x)
e.record()
torch.cuda.synchronize()
eager_us = s.elapsed_time(e) / a.iters * 1000
g = torch.cuda.CUDAGraph()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    dist.all_reduce(x)
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
with torch.cuda.graph(g):
    dist.all_reduce(x)
for _ in range(5):
    g.replay()
torch.cuda.synchronize()
dist.barrier()
s.record()
for _ in range(a.iters):
    g.replay()
e.record()
torch.cuda.synchronize()
graph_us = s.elapsed_time(e) / a.iters * 1000
if dist.get_rank() == 0:
    print({"world": dist.get_world_size(), "bytes": x.numel() * 4, "eager_us": round(eager_us, 1), "graph_us": round(graph_us, 1)},
          flush=True)
torch.cuda.synchronize()
os._exit(0)
