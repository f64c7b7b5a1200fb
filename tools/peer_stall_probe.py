"""Diagnostic: after a rank has waited milliseconds for a late peer inside a kernel, when does the HOST see the
completion of that GPU's following work?  (2 ranks; rank 1 is made late on purpose in iteration 0 and 4.)
    torchrun --nproc-per-node 2 tools/peer_stall_probe.py"""
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
comm = PeerAllReduce(n, device=dev)
small = PeerAllReduce(64, device=dev)
plain = torch.zeros(n, device=dev)
pin = torch.zeros(n).pin_memory()
side = torch.cuda.Stream()
main = torch.cuda.current_stream()
LATE = 0.004


def trial(name, launch, src, copy_stream):
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    t0 = time.perf_counter()
    rows = []
    for i in range(8):
        if i in (0, 4) and rank == 1:
            time.sleep(LATE)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        launch(i)
        e1.record()
        cs = side if copy_stream == "side" else main
        if cs is side:
            side.wait_stream(main)
        with torch.cuda.stream(cs):
            pin.copy_(src, non_blocking=True)
            e2.record()
        while not e2.query():
            pass
        seen = (time.perf_counter() - t0) * 1e3
        rows.append((round(e0.elapsed_time(e1), 3), round(base.elapsed_time(e2), 3), round(seen, 3)))
    late = [round(r[2] - r[1], 3) for r in rows]
    out = [None] * world
    dist.all_gather_object(out, (rows, late))
    if rank == 0:
        for r, (rows_, late_) in enumerate(out):
            print(f"{name:34s} rank {r}: kernel ms {[x[0] for x in rows_]}  host-late ms {late_}", flush=True)


def sleep_kernel(i):
    if rank == 0 and i in (0, 4):
        torch.cuda._sleep(int(LATE * 1.9e9))


nccl_buf = torch.zeros(n, device=dev)
trial("peer all-reduce 3.4MB, side copy", lambda i: comm.all_reduce_(), comm.flat, "side")
trial("peer all-reduce 3.4MB, main copy", lambda i: comm.all_reduce_(), comm.flat, "main")
trial("peer all-reduce, copy plain buffer", lambda i: comm.all_reduce_(), plain, "side")
trial("peer all-reduce 256 B", lambda i: small.all_reduce_(), plain, "side")
trial("local sleep kernel only (no peer)", sleep_kernel, plain, "side")
trial("nccl all-reduce 3.4MB", lambda i: dist.all_reduce(nccl_buf), nccl_buf, "side")
st = comm.status()
if rank == 0:
    print("comm status", st, "GHR_COMM_DBG", os.environ.get("GHR_COMM_DBG"))
torch.cuda.synchronize()
dist.barrier()
os._exit(0)
