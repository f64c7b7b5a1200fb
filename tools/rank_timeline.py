"""Where a multi-GPU step's time goes, per rank: local compute (forward + backward of the rank's views,
graph replay without the collective) vs the all-reduce (which also absorbs the wait for the slowest
rank), with and without an L2 flush between steps.  Written to explain the 8-rank step (682 us against
444 us on one GPU while every kernel's duration is unchanged).

    torchrun --nproc-per-node N tools/rank_timeline.py [--steps 50] [--views 8]

NOT yet run on hardware (the round's GPU budget was spent when it was written)."""
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
from guassianhand_b200 import api, scenes  # noqa: E402
from guassianhand_b200.dist import GraphedFitStep, PackedGrads, balanced_shards, fit_step_grads  # noqa: E402
import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--views", type=int, default=8)
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
P, H, W, B = 60000, 512, 334, a.views
sc = scenes.two_hand_scene(P, seed=0)
cams = scenes.fibonacci_cameras(64, H, W, seed=0)
bg = np.zeros(3, np.float32)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
             colors_precomp=t(sc.colors))
# the bench's assignment: the step's world*B views dealt to the ranks by instance count
costs = []
for c in cams[:B * world]:
    v1 = util.gpu_views([c], bg, dev)
    costs.append(api.forward_raw(v1.cams(), gauss["means3D"], gauss["opacities"], gauss["scales"],
                                 gauss["rotations"], None, None, gauss["colors_precomp"], 0, 1.0).R)
mine = balanced_shards(costs, world)[rank]
views = util.gpu_views([cams[i] for i in mine], bg, dev)
dL = t((np.random.default_rng(1 + rank).normal(size=(B, 3, H, W)) / (H * W)).astype(np.float32))
grads = PackedGrads(P, 0, device=dev)
r = fit_step_grads(gauss, views, dL, grads, overlap=2, group=None)
caps = [int(x.R * 1.25) + (1 << 14) for x in r.results]


# graph of the LOCAL part only: PackedGrads.all_reduce_ is skipped by capturing on a private buffer whose
# all_reduce_ is a no-op
class LocalGrads(PackedGrads):
    def all_reduce_(self, group=None, async_op=False):
        return None


lgrads = LocalGrads(P, 0, device=dev)
step = GraphedFitStep(gauss, views, dL, lgrads, R_cap=caps, overlap=2)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
out = {}
for mode in ("no_flush", "flush"):
    K = a.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    for i in range(5):
        step.replay()
        if world > 1:
            dist.all_reduce(lgrads.flat)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    for i in range(K):
        if mode == "flush":
            flush.zero_()
        ev[i][0].record()
        step.replay()
        ev[i][1].record()
        if world > 1:
            dist.all_reduce(lgrads.flat)
        ev[i][2].record()
    torch.cuda.synchronize()
    comp = float(np.mean([e[0].elapsed_time(e[1]) for e in ev])) * 1000
    coll = float(np.mean([e[1].elapsed_time(e[2]) for e in ev])) * 1000
    both = torch.tensor([comp, coll], device=dev, dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(both) for _ in range(world)]
        dist.all_gather(allv, both)
    else:
        allv = [both]
    out[mode] = {"compute_us_per_rank": [round(float(v[0]), 1) for v in allv],
                 "allreduce_plus_wait_us_per_rank": [round(float(v[1]), 1) for v in allv]}
if rank == 0:
    out["world"], out["views_per_rank"] = world, B
    print(json.dumps(out), flush=True)
torch.cuda.synchronize()
os._exit(0)
