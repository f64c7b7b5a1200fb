"""Camera-sharded step on >= 2 GPUs (skipped on a single-GPU box): the all-reduced packed gradients of an N-rank
step -- through NCCL and through libghr's own NVLink peer-memory kernel (csrc/comm.cu), eagerly and replayed
from a CUDA graph -- against rank 0 rendering ALL views itself.  Replaces the reference's Lightning-DDP
gradient all-reduce (/root/reference/infer_one_shot.py:631,638)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from guassianhand_b200 import scenes
    from guassianhand_b200.dist import GraphedFitStep, PackedGrads, fit_step_grads, shard_views
    import util
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    res = {}
    try:
        P, V, H, W = 3001, 3 * world, 80, 96                  # odd P: padded segments
        sc = scenes.two_hand_scene(P, seed=31)
        cams = scenes.fibonacci_cameras(V, H, W, seed=31)
        bg = np.array([0.1, 0.0, 0.2], np.float32)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).float().to(dev)
        gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
                     colors_precomp=t(sc.colors))
        w_all = (np.random.default_rng(6).normal(size=(V, 3, H, W)) / (H * W)).astype(np.float32)
        mine = list(shard_views(V, rank, world))
        views = util.gpu_views([cams[i] for i in mine], bg, dev)
        w = t(w_all[mine])
        # reference: every rank renders all views itself (no communication)
        ref = PackedGrads(P, 0, device=dev)
        fit_step_grads(gauss, util.gpu_views(cams, bg, dev), t(w_all), ref, group=False)
        want = ref.flat.double().cpu().numpy()
        scale = np.abs(want).max()
        for name, peer in (("nccl", False), ("peer", True)):
            g = PackedGrads(P, 0, device=dev, peer=peer)
            assert (g.comm is not None) == peer
            r = fit_step_grads(gauss, views, w, g)
            torch.cuda.synchronize()
            res[name + "_eager"] = float(np.abs(g.flat.double().cpu().numpy() - want).max() / scale)
            cap = int(r.R * 1.25) + 1024
            step = GraphedFitStep(gauss, views, w, g, R_cap=cap)
            for _ in range(3):
                g.flat.zero_()
                step.replay()
            torch.cuda.synchronize()
            res[name + "_graph"] = float(np.abs(g.flat.double().cpu().numpy() - want).max() / scale)
            assert not step.status()[1]
            if peer:
                res["peer_status"] = g.comm.status()
            # identical bits on every rank (the peer kernel sums in rank order on the slice owner)
            gathered = [torch.empty_like(g.flat) for _ in range(world)]
            dist.all_gather(gathered, g.flat.clone())
            res[name + "_same_bits"] = all(torch.equal(gathered[0], x) for x in gathered[1:])
        torch.cuda.synchronize()
        dist.barrier()
    except Exception as e:  # noqa: BLE001
        import traceback
        res["error"] = traceback.format_exc()
    out.put((rank, res))
    out.close()
    out.join_thread()          # the result has left this process before it exits hard
    # leave without tearing NCCL down: destroy_process_group() with NCCL captured in live graphs can block
    torch.cuda.synchronize()
    os._exit(0)


@pytest.mark.parametrize("world", [2])
def test_sharded_step_allreduce_matches_single_rank(cuda_device, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=150) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for rank, res in results.items():
        assert "error" not in res, res.get("error")
        for k in ("nccl_eager", "nccl_graph", "peer_eager", "peer_graph"):
            assert res[k] <= 1e-5, (rank, k, res[k])
        assert res["peer_same_bits"]
        assert res["peer_status"][1] == 0 and res["peer_status"][0] >= 4
