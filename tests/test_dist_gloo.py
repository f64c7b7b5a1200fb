"""Host logic of the camera-sharded multi-GPU path on CPU: view sharding, the packed gradient
buffer and its all-reduce, with torch.distributed `gloo`, world_size 2 (no GPU needed)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from guassianhand_b200.dist import PackedGrads, shard_views


def test_shard_views_partitions_exactly():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            got = [i for r in range(world) for i in shard_views(n, r, world)]
            assert got == list(range(n))
            sizes = [len(shard_views(n, r, world)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_views(4, 2, 2)


def test_packed_grads_layout():
    g = PackedGrads(P=10, M=0)
    assert g.floats_per_gaussian == 14                                  # 56 B x P
    # every segment is padded to a multiple of 4 floats: 32 + 32 + 40 + 12 + 32 floats for P = 10
    assert g.offsets == [0, 32, 64, 104, 116] and g.nbytes == 148 * 4
    v = g.views()
    v["dL_dscales"][3, 1] = 5.0
    assert g.flat[32 + 3 * 3 + 1] == 5.0                               # aliasing, segment order
    # odd Gaussian counts (one reference hand has 49,281 points): every segment still starts 16-byte aligned
    for P in (1, 7, 49281):
        go = PackedGrads(P=P, M=0)
        assert all(o % 4 == 0 for o in go.offsets)
        assert all(t.data_ptr() % 16 == 0 for t in go.views().values())
        assert go.views()["dL_drotations"].shape == (P, 4)
    gs = PackedGrads(P=4, M=16)
    assert gs.floats_per_gaussian == 11 + 48 and gs.views()["dL_dsh"].shape == (4, 16, 3)
    assert g.all_reduce_() is None                                      # no process group: no-op


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_views, P, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank derives the same per-view "gradient" deterministically, renders its shard only
        g = PackedGrads(P=P, M=0)
        torch.manual_seed(0)
        per_view = torch.randn(n_views, g.flat.numel())
        for v in shard_views(n_views, rank, world):
            g.flat += per_view[v]                                       # stands in for ghr_backward(+=)
        g.all_reduce_()
        want = per_view.sum(0)
        ok = torch.allclose(g.flat, want, atol=1e-5)
        gathered = [None] * world
        dist.all_gather_object(gathered, (rank, bool(ok), list(shard_views(n_views, rank, world))))
        if rank == 0:
            out.put(gathered)
    finally:
        dist.destroy_process_group()


def test_camera_sharded_allreduce_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 7, 33, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert sorted(i for _, _, idx in res for i in idx) == list(range(7))


def test_balanced_shards():
    from guassianhand_b200.dist import balanced_shards
    import random
    rnd = random.Random(0)
    costs = [rnd.uniform(0.8, 1.3) for _ in range(64)]
    for world in (1, 2, 4, 8):
        sh = balanced_shards(costs, world)
        assert sorted(i for s in sh for i in s) == list(range(64)) and all(len(s) == 64 // world for s in sh)
        sums = [sum(costs[i] for i in s) for s in sh]
        naive = [sum(costs[r * (64 // world):(r + 1) * (64 // world)]) for r in range(world)]
        assert max(sums) - min(sums) <= max(naive) - min(naive) + 1e-12
        assert max(sums) / (sum(costs) / world) < 1.02
    try:
        balanced_shards(costs[:63], 2)
        assert False
    except ValueError:
        pass
