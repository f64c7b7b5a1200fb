"""Where a multi-GPU step's time goes, per rank and per step: the graphed camera-sharded step (forward +
backward of the rank's views + all-reduce) timed with CUDA events around every replay, next to the same graph
WITHOUT the collective (local compute only) and the collective alone.

    torchrun --nproc-per-node N tools/rank_timeline.py [--steps 60] [--views 8] [--allreduce peer|nccl]
                                                        [--overlap 2] [--no-flush]
Prints one JSON object: per rank median / p90 / max of {step, local compute, all-reduce alone} and the indices
and durations of the slowest steps."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import _native as NV, api, scenes  # noqa: E402
from guassianhand_b200.dist import GraphedFitStep, PackedGrads, fit_step_grads  # noqa: E402
import util  # noqa: E402
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--views", type=int, default=8)
ap.add_argument("--allreduce", default="peer", choices=["peer", "nccl"])
ap.add_argument("--overlap", type=int, default=2)
ap.add_argument("--no-flush", action="store_true")
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
P, H, W, B = 60000, 512, 334, a.views
N = H * W
sc = scenes.two_hand_scene(P, seed=0)
cams = scenes.fibonacci_cameras(64, H, W, seed=0)
bg = np.zeros(3, np.float32)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
             colors_precomp=t(sc.colors))
# the bench's deal: the step's world*B views dealt to the ranks by blend pairs
costs = []
for c in cams[:B * world]:
    v1 = util.gpu_views([c], bg, dev)
    r1 = api.forward_raw(v1.cams(), gauss["means3D"], gauss["opacities"], gauss["scales"], gauss["rotations"], None, None,
                         gauss["colors_precomp"], 0, 1.0)
    lay1 = NV.layout(P, 1, H, W, 0, 0, r1.R_cap)
    costs.append(float(r1.state[lay1.off_ncontrib: lay1.off_ncontrib + N * 4].view(torch.int32).sum(dtype=torch.int64).item()))
mine = bench.deal_views(costs, list(range(B * world)), 0, 1, world)[rank] if world > 1 else list(range(B))
views = util.gpu_views([cams[i] for i in mine], bg, dev)
dL = t((np.random.default_rng(1 + rank).normal(size=(B, 3, H, W)) / N).astype(np.float32))
G = max(1, min(a.overlap, B))
peer = world > 1 and a.allreduce == "peer"
grads = PackedGrads(P, 0, device=dev, peer=peer)
r = fit_step_grads(gauss, views, dL, grads, overlap=G, group=False)
caps = [int(x.R * 1.25) + (1 << 14) for x in r.results] if G > 1 else int(r.R * 1.25) + (1 << 14)
pairs_mine = float(sum(costs[i] for i in mine))
step = GraphedFitStep(gauss, views, dL, grads, R_cap=caps, overlap=G)                 # with the all-reduce


class LocalGrads(PackedGrads):
    def all_reduce_(self, group=None, async_op=False):
        return None


lgrads = LocalGrads(P, 0, device=dev)
local_step = GraphedFitStep(gauss, views, dL, lgrads, R_cap=caps, overlap=G)            # local compute only
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n)]
    for s, e in ev:
        if not a.no_flush:
            flush.zero_()
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    return np.array([s.elapsed_time(e) for s, e in ev]) * 1000.0


def summ(us):
    worst = np.argsort(-us)[:5]
    return {"median_us": round(float(np.median(us)), 1), "p90_us": round(float(np.percentile(us, 90)), 1),
            "max_us": round(float(us.max()), 1), "mean_us": round(float(us.mean()), 1),
            "slowest": [(int(i), round(float(us[i]), 1)) for i in worst]}


out = {"rank": rank, "blend_pairs": pairs_mine}
out["step"] = summ(timed(step.replay, a.steps))
out["local_compute"] = summ(timed(local_step.replay, a.steps))
out["allreduce_alone"] = summ(timed(lambda: grads.all_reduce_(), a.steps)) if world > 1 else None
if peer:
    out["peer_status"] = grads.comm.status()
allr = [out]
if world > 1:
    allr = [None] * world
    dist.all_gather_object(allr, out)
if rank == 0:
    print(json.dumps({"world": world, "views_per_rank": B, "allreduce": a.allreduce if world > 1 else None,
                      "overlap": G, "flush": not a.no_flush, "per_rank": allr}), flush=True)
torch.cuda.synchronize()
os._exit(0)
