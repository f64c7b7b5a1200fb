"""Diagnostic 2: the e2e leg's pipelined shape (launch i, copy down on a side stream, wait for i-1) with the pieces
switched on one at a time.  torchrun --nproc-per-node 2 tools/peer_stall_probe2.py"""
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from guassianhand_b200.dist import PeerAllReduce  # noqa: E402

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
n = 60000 * 14
comms = [PeerAllReduce(n, device=dev) for _ in range(2)]
nccl_bufs = [torch.zeros(n, device=dev) for _ in range(2)]
pins = [torch.zeros(n).pin_memory() for _ in range(2)]
hsrc = torch.zeros(n).pin_memory()
dsts = [torch.zeros(n, device=dev) for _ in range(2)]
xs = [torch.randn(2048, 2048, device=dev) for _ in range(2)]
ys = [torch.empty(2048, 2048, device=dev) for _ in range(2)]
zs = [torch.zeros(1 << 20, device=dev) for _ in range(2)]
down, up, fork = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
main = torch.cuda.current_stream()
LATE = 0.004


def body(s, forkjoin, coll):
    cur = torch.cuda.current_stream()
    if forkjoin:
        fork.wait_stream(cur)
        with torch.cuda.stream(fork):
            zs[s].add_(1.0)
    torch.mm(xs[s], xs[s], out=ys[s])
    if forkjoin:
        cur.wait_stream(fork)
    if coll == "peer":
        comms[s].all_reduce_()
    elif coll == "nccl":
        dist.all_reduce(nccl_bufs[s])


def trial(name, graph=False, forkjoin=False, timing=True, upload=False, coll="peer", steps=10):
    graphs = []
    if graph:
        for s in range(2):
            torch.cuda.synchronize()
            side = torch.cuda.Stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                body(s, forkjoin, coll)
            main.wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                body(s, forkjoin, coll)
            graphs.append(g)
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    t0 = time.perf_counter()
    evd = [torch.cuda.Event(enable_timing=timing) for _ in range(2)]
    evu = [torch.cuda.Event(enable_timing=timing) for _ in range(2)]
    evused = [torch.cuda.Event(enable_timing=timing) for _ in range(2)]
    for s in range(2):
        evused[s].record(main)
    done_dev, seen, kms = [], [], []

    def do_upload(i):
        s = i % 2
        with torch.cuda.stream(up):
            up.wait_event(evused[s])
            dsts[s].copy_(hsrc, non_blocking=True)
            evu[s].record(up)

    def wait(i):
        s = i % 2
        while not evd[s].query():
            pass
        seen.append((time.perf_counter() - t0) * 1e3)

    if upload:
        do_upload(0)
    for i in range(steps):
        s = i % 2
        if i == 0 and rank == 1:
            time.sleep(LATE)
        if upload and i + 1 < steps:
            do_upload(i + 1)
        if upload:
            main.wait_event(evu[s])
        main.wait_event(evd[s])
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(main)
        if graph:
            graphs[s].replay()
        else:
            body(s, forkjoin, coll)
        e1.record(main)
        evused[s].record(main)
        down.wait_stream(main)
        with torch.cuda.stream(down):
            pins[s].copy_(comms[s].flat if coll == "peer" else nccl_bufs[s], non_blocking=True)
            e2.record(down)
            evd[s].record(down)
        kms.append((e0, e1))
        done_dev.append(e2)
        if i >= 1:
            wait(i - 1)
    wait(steps - 1)
    torch.cuda.synchronize()
    late = [round(seen[i] - base.elapsed_time(done_dev[i]), 3) for i in range(steps)]
    k = [round(a.elapsed_time(b), 3) for a, b in kms]
    out = [None] * world
    dist.all_gather_object(out, (k, late))
    if rank == 0:
        for r, (k_, late_) in enumerate(out):
            print(f"{name:44s} rank {r}: step ms {k_}  host-late ms {late_}", flush=True)


trial("eager pipelined")
trial("eager pipelined, non-timing events", timing=False)
trial("eager + uploads", upload=True, timing=False)
trial("graph", graph=True, timing=False)
trial("graph + fork/join", graph=True, forkjoin=True, timing=False)
trial("graph + fork/join + uploads", graph=True, forkjoin=True, upload=True, timing=False)
trial("graph + fork/join + uploads, NCCL", graph=True, forkjoin=True, upload=True, timing=False, coll="nccl")
trial("graph + fork/join + uploads, no collective", graph=True, forkjoin=True, upload=True, timing=False, coll="none")
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
