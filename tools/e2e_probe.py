"""Bisecting tool: the e2e pipeline of bench.py reduced to [fill + peer all-reduce] graphs with H2D / D2H copies
on side streams.  torchrun --nproc-per-node 2 tools/e2e_probe.py [--no-h2d] [--no-d2h] [--clone-d2h] [--work-us 400]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guassianhand_b200.dist import PackedGrads  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--no-h2d", action="store_true")
ap.add_argument("--no-d2h", action="store_true")
ap.add_argument("--clone-d2h", action="store_true")
ap.add_argument("--nccl", action="store_true")
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--work-us", type=int, default=400)
ap.add_argument("--slots", type=int, default=1)
ap.add_argument("--status-reads", action="store_true")
a = ap.parse_args()
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
rank, world = dist.get_rank(), dist.get_world_size()
P = 60000
host_in = torch.randn(P * 14).pin_memory()
work = torch.empty(16 << 20, dtype=torch.float32, device=dev)      # 64 MB: one mul_ pass ~ 25 us
n_pass = max(1, a.work_us // 25)
n_work = work.numel()


def do_work():
    for _ in range(n_pass):
        work.mul_(1.0001)
slots = []
pin_small = torch.zeros(8).pin_memory()
main = torch.cuda.current_stream()
up, down = torch.cuda.Stream(), torch.cuda.Stream()
for s in range(a.slots):
    g = PackedGrads(P, 0, device=dev, peer=not a.nccl)
    din = torch.empty(P * 14, device=dev)
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        do_work()
        g.flat.copy_(din[:g.flat.numel()])
        g.all_reduce_()
    main.wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph):
        do_work()
        g.flat.copy_(din[:g.flat.numel()])
        g.all_reduce_()
    slots.append(dict(g=g, din=din, graph=graph, host_out=torch.empty(g.flat.numel()).pin_memory(),
                      ev_up=torch.cuda.Event(), ev_used=torch.cuda.Event(), ev_down=torch.cuda.Event()))


def upload(i):
    sl = slots[i % len(slots)]
    with torch.cuda.stream(up):
        up.wait_event(sl["ev_used"])
        if not a.no_h2d:
            sl["din"].copy_(host_in, non_blocking=True)
        sl["ev_up"].record(up)


def render(i):
    sl = slots[i % len(slots)]
    main.wait_event(sl["ev_up"])
    main.wait_event(sl["ev_down"])
    sl["graph"].replay()
    if a.status_reads:
        pin_small.copy_(work[:8], non_blocking=True)
    sl["ev_used"].record(main)
    down.wait_stream(main)
    with torch.cuda.stream(down):
        if not a.no_d2h:
            src = sl["g"].flat.clone() if a.clone_d2h else sl["g"].flat
            sl["host_out"].copy_(src, non_blocking=True)
        sl["ev_down"].record(down)


def collect(i):
    slots[i % len(slots)]["ev_down"].synchronize()


def run(n):
    marks = []
    for sl in slots:
        sl["ev_used"].record(main)
        sl["ev_down"].record(main)
    upload(0)
    for i in range(n):
        if i + 1 < n:
            upload(i + 1)
        render(i)
        if i >= 1:
            collect(i - 1)
            marks.append(time.perf_counter())
    collect(n - 1)
    marks.append(time.perf_counter())
    torch.cuda.synchronize()
    return marks


run(8)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
m = run(a.steps)
steps = np.diff(np.array([t0] + m)) * 1e3
res = {"rank": rank, "median_ms": round(float(np.median(steps)), 3), "max_ms": round(float(steps.max()), 3),
       "steps": [round(float(x), 2) for x in steps]}
allr = [None] * world
dist.all_gather_object(allr, res)
if rank == 0:
    print(json.dumps({"args": vars(a), "per_rank": allr}), flush=True)
torch.cuda.synchronize()
os._exit(0)
