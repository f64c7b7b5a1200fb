"""Camera-sharded data parallelism for the multi-view fitting step.

The reference's only multi-GPU mechanism is Lightning DDP (/root/reference/infer_one_shot.py:631,638):
an NCCL all-reduce of ~444 MB of *parameter* gradients per step.  Here the views of a step are
sharded across ranks (one process per GPU), Gaussians are replicated, the forward needs no
communication, and the only exchange is ONE all-reduce (sum) of the packed Gaussian-attribute
gradient buffer -- 56 B x P for the colors_precomp path.  The rasterizer backward writes straight
into that packed buffer (one flat allocation, one segment per attribute), so there is no
pack/concat kernel before the collective.
"""
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world: int) -> range:
    """Contiguous, balanced split of view indices [0, n_views): the first (n_views % world) ranks
    take one extra view."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_views, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def balanced_shards(costs: Sequence[float], world: int) -> List[List[int]]:
    """Split view indices [0, len(costs)) into `world` shards of equal size (len(costs) must be a
    multiple of world) whose summed costs are as equal as a sort + snake deal makes them: views in
    descending cost order go to ranks 0..w-1, w-1..0, 0..w-1, ...  Every rank computes the same
    assignment from the same costs.  The step ends with an all-reduce, i.e. at the slowest rank: with
    per-view instance counts as costs the ranks finish together instead of a few percent apart."""
    n = len(costs)
    if world < 1 or n % world:
        raise ValueError(f"{n} views do not split evenly over {world} ranks")
    order = sorted(range(n), key=lambda i: (-float(costs[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for k, i in enumerate(order):
        rnd, pos = divmod(k, world)
        shards[pos if rnd % 2 == 0 else world - 1 - pos].append(i)
    return [sorted(s) for s in shards]


class _RawCuda:
    """Zero-copy view of library-owned device memory for torch.as_tensor (CUDA array interface v2)."""

    def __init__(self, ptr: int, nfloats: int):
        self.__cuda_array_interface__ = {"shape": (int(nfloats),), "typestr": "<f4", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


class PeerAllReduce:
    """Sum of one float buffer over the ranks of ONE node by libghr's own kernel over NVLink peer memory
    (include/ghr.h ghr_comm_*, csrc/comm.cu): two-shot, in place, deterministic, one launch per call,
    CUDA-graph capturable.  torch.distributed is used once, to exchange the 128-byte IPC handles.

    `flat` is the buffer (a torch view of the library's allocation): write into it, call all_reduce_().
    Every rank must issue the same sequence of all_reduce_ calls."""

    def __init__(self, nfloats: int, group=None, device=None):
        import ctypes as C
        from . import _native as N
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("PeerAllReduce needs an initialised torch.distributed process group")
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > N.GHR_COMM_MAX_RANKS:
            raise RuntimeError(f"PeerAllReduce: at most {N.GHR_COMM_MAX_RANKS} ranks (one NVLink node)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.nfloats = (int(nfloats) + 3) // 4 * 4
        self._N, self._comm = N, C.c_void_p()
        with torch.cuda.device(self.device):
            N.check(N.lib().ghr_comm_create(self.rank, self.world, self.nfloats * 4, C.byref(self._comm)), "ghr_comm_create")
            mine = C.create_string_buffer(N.GHR_COMM_HANDLE_BYTES)
            N.check(N.lib().ghr_comm_handle(self._comm, mine), "ghr_comm_handle")
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine.raw), group=group)
            N.check(N.lib().ghr_comm_connect(self._comm, b"".join(gathered)), "ghr_comm_connect")
        ptr = N.lib().ghr_comm_buffer(self._comm)
        self._raw = _RawCuda(ptr, self.nfloats)
        self.flat = torch.as_tensor(self._raw, device=self.device)
        dist.barrier(group=group)       # every rank has mapped every buffer before anyone launches

    def all_reduce_(self, nfloats: Optional[int] = None):
        n = self.nfloats if nfloats is None else (int(nfloats) + 3) // 4 * 4
        stream = torch.cuda.current_stream(self.device).cuda_stream
        self._N.check(self._N.lib().ghr_comm_allreduce(self._comm, n, stream), "ghr_comm_allreduce")

    def status(self) -> Tuple[int, int]:
        """(all-reduces completed, error word); synchronises the device."""
        import ctypes as C
        ep, err = C.c_uint32(), C.c_uint32()
        torch.cuda.synchronize(self.device)
        self._N.check(self._N.lib().ghr_comm_status(self._comm, C.byref(ep), C.byref(err)), "ghr_comm_status")
        return int(ep.value), int(err.value)

    def close(self):
        if self._comm:
            torch.cuda.synchronize(self.device)
            self.flat = None
            self._N.lib().ghr_comm_destroy(self._comm)
            self._comm = None


class PackedGrads:
    """One flat fp32 buffer holding every Gaussian-attribute gradient that is summed over views.

    Segment order (floats per Gaussian): means3D 3 | scales 3 | rotations 4 | opacity 1 |
    colors 3 (colors_precomp path) or sh 3*M (SH path).  `views()` returns tensors aliasing the
    buffer with the shapes ghr_backward writes."""

    def __init__(self, P: int, M: int = 0, device="cpu", with_cov3D: bool = False, peer: bool = False, group=None):
        """peer=True (N > 1 ranks on one node): the buffer lives in a PeerAllReduce allocation and
        all_reduce_() is libghr's NVLink kernel instead of NCCL."""
        self.P, self.M = int(P), int(M)
        self.comm: Optional[PeerAllReduce] = None
        segs: List[Tuple[str, Tuple[int, ...]]] = [
            ("dL_dmeans3D", (P, 3)), ("dL_dscales", (P, 3)), ("dL_drotations", (P, 4)), ("dL_dopacity", (P, 1))]
        segs.append(("dL_dsh", (P, M, 3)) if M > 0 else ("dL_dcolors", (P, 3)))
        if with_cov3D:
            segs.append(("dL_dcov3D", (P, 6)))
        self.segments = segs
        # every segment starts on a 16-byte boundary (the kernels write rotations and SH rows with 128-bit
        # stores): sizes are rounded up to 4 floats, the padding floats stay zero and are all-reduced along
        sizes = [(int(torch.Size(s).numel()) + 3) // 4 * 4 for _, s in segs]
        self.offsets = [sum(sizes[:i]) for i in range(len(sizes))]
        if peer and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            self.comm = PeerAllReduce(sum(sizes), group=group, device=device)
            self.flat = self.comm.flat[:sum(sizes)]
            self.flat.zero_()
        else:
            self.flat = torch.zeros(sum(sizes), dtype=torch.float32, device=device)
        self._views: Dict[str, torch.Tensor] = {}
        for (name, shape), o in zip(segs, self.offsets):
            k = int(torch.Size(shape).numel())
            self._views[name] = self.flat[o:o + k].view(*shape)

    @property
    def floats_per_gaussian(self) -> int:
        return sum(int(torch.Size(s).numel()) for _, s in self.segments) // max(self.P, 1)

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * 4

    def views(self) -> Dict[str, torch.Tensor]:
        return dict(self._views)

    def zero_(self):
        self.flat.zero_()
        return self

    def all_reduce_(self, group=None, async_op: bool = False):
        """Sum over ranks, in place.  No-op without an initialised process group (N=1)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if self.comm is not None:
            return self.comm.all_reduce_(self.flat.numel())
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class StepResult(NamedTuple):
    """Result of an overlapped step: one api.ForwardResult per view group (contiguous view blocks)."""
    results: list
    bounds: list              # range of views per group
    color: torch.Tensor       # [V,3,H,W]: the groups' images are blocks of one tensor
    radii: torch.Tensor       # [V,P]

    @property
    def R(self):
        return None if any(r.R is None for r in self.results) else sum(r.R for r in self.results)

    @property
    def R_cap(self):
        return [r.R_cap for r in self.results]



_side_streams: Dict[int, list] = {}


def _streams(dev: torch.device, n: int):
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    pool = _side_streams.setdefault(idx, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=idx))
    return pool[:n]


def _slice_views(views, lo: int, hi: int):
    bg = views.bg if views.bg.dim() == 1 else views.bg[lo:hi]
    return views._replace(viewmatrix=views.viewmatrix[lo:hi], projmatrix=views.projmatrix[lo:hi],
                          campos=views.campos[lo:hi], tanfov=views.tanfov[lo:hi], bg=bg)


def fit_step_grads(gauss: Dict[str, torch.Tensor], views, dL_dout: torch.Tensor, grads: PackedGrads,
                   group=None, sh_degree: int = 0, scale_modifier: float = 1.0,
                   R_cap=None, check: str = "poll", fwd_events=None, bwd_events=None,
                   overlap: int = 1, partials: Optional[Sequence[PackedGrads]] = None,
                   temps: Optional[Sequence[torch.Tensor]] = None):
    """One camera-sharded step on this rank: forward + backward of the LOCAL views, gradients written
    into `grads` (overwritten), then the all-reduce.  Returns the api.ForwardResult (color, radii, state).

    gauss: dict with means3D, opacities, scales, rotations and colors_precomp or shs (CUDA fp32).
    group: the process group of the all-reduce (None = default), or False to skip the all-reduce.
    views: guassianhand_b200.api.ViewBatch of the local shard.  dL_dout: [V,3,H,W].

    overlap = G > 1 cuts the local views into G contiguous groups and runs each group's forward ->
    backward chain on its own stream (group 0 on the current one).  The chains are independent until the
    gradient sum, so the single-CTA tile scan, the half-empty sort launches and the tails of the blend
    kernels of one group are filled with the other groups' work.  Group g > 0 writes into
    `partials[g-1]` (allocated when absent; pass static buffers under CUDA-graph capture), which are
    added into `grads` after the join.  `R_cap` may then be a list (one capacity per group).  Returns a
    StepResult.  `temps`: one caller-owned scratch buffer per group (see step_temp_bytes) instead of the
    per-stream workspace -- GraphedFitStep passes its own so that the captured pointers stay valid."""
    from . import api
    f32 = api._f32c
    means3D, opac = f32(gauss["means3D"]), f32(gauss["opacities"])
    sc, rot = f32(gauss["scales"]), api._f32a(gauss["rotations"])
    shs = f32(gauss["shs"]) if gauss.get("shs") is not None else None
    col = f32(gauss["colors_precomp"]) if gauss.get("colors_precomp") is not None else None
    V = int(views.viewmatrix.shape[0])
    G = max(1, min(int(overlap), V))
    if G == 1:
        cams = views.cams()
        cap = R_cap[0] if isinstance(R_cap, (list, tuple)) else R_cap
        tmp = temps[0] if temps else None
        res = api.forward_raw(cams, means3D, opac, sc, rot, None, shs, col, sh_degree, scale_modifier, check=check,
                              R_cap=cap, stage_events=fwd_events, temp=tmp)
        api.backward_raw(cams, res.state, res.R_cap, dL_dout, means3D, opac, sc, rot, None, shs, col, sh_degree,
                         scale_modifier, want_means2D=False, accumulate_into=grads.views(), accumulate=False,
                         stage_events=bwd_events, temp=tmp)
        if group is not False:
            grads.all_reduce_(group)
        return res
    if fwd_events is not None or bwd_events is not None:
        raise ValueError("per-stage events need overlap=1 (stages of different groups run concurrently)")
    dev = means3D.device
    bounds = [shard_views(V, g, G) for g in range(G)]
    if partials is None:
        partials = [PackedGrads(grads.P, grads.M, device=dev) for _ in range(G - 1)]
    main = torch.cuda.current_stream(dev)
    streams = [main] + _streams(dev, G - 1)
    P, H, W = means3D.shape[0], int(views.image_height), int(views.image_width)
    color = torch.empty(V, 3, H, W, dtype=torch.float32, device=dev)
    radii = torch.empty(V, max(P, 1), dtype=torch.int32, device=dev)
    for st in streams[1:]:
        st.wait_stream(main)          # fork before group 0 is enqueued: the chains must not wait for it
    results = []
    for g, rng in enumerate(bounds):
        st = streams[g]
        with torch.cuda.stream(st):
            cams = _slice_views(views, rng.start, rng.stop).cams()
            cap = R_cap[g] if isinstance(R_cap, (list, tuple)) else R_cap
            tmp = temps[g] if temps else None
            res = api.forward_raw(cams, means3D, opac, sc, rot, None, shs, col, sh_degree, scale_modifier,
                                  check=check, R_cap=cap, out=(color[rng.start:rng.stop], radii[rng.start:rng.stop], None),
                                  temp=tmp)
            target = grads if g == 0 else partials[g - 1]
            api.backward_raw(cams, res.state, res.R_cap, dL_dout[rng.start:rng.stop], means3D, opac, sc, rot, None,
                             shs, col, sh_degree, scale_modifier, want_means2D=False,
                             accumulate_into=target.views(), accumulate=False, temp=tmp)
        results.append(res)
    for g in range(1, G):
        main.wait_stream(streams[g])
    for p in partials[:G - 1]:
        grads.flat.add_(p.flat)
    if group is not False:
        grads.all_reduce_(group)
    return StepResult(results, bounds, color, radii[:, :P])


def step_temp_bytes(P: int, V: int, H: int, W: int, M: int, sh_degree: int, R_cap: int) -> int:
    """Scratch bytes one forward + backward of V views needs (max of the two layouts)."""
    from . import _native as N
    lay = N.layout(P, V, H, W, M, sh_degree, int(R_cap))
    return int(max(lay.temp_bytes, lay.temp_bwd_bytes))


class GraphedFitStep:
    """The camera-sharded step (forward + backward [+ all-reduce]) captured once into a CUDA graph
    and replayed: ~16 kernel launches, 2 memsets and the NCCL call become one graph launch, which is
    what makes the single-view (latency-bound) case GPU-bound instead of host-bound.

    All inputs are static device buffers: write new Gaussian attributes / cameras / dL_dout into
    `gauss`, `views` tensors and `dL_dout` (in place) before `replay()`.  Capacity is fixed
    (`R_cap`); check `status()` for overflow when the scene changes a lot."""

    def __init__(self, gauss: Dict[str, torch.Tensor], views, dL_dout: torch.Tensor, grads: PackedGrads,
                 R_cap, group=None, sh_degree: int = 0, scale_modifier: float = 1.0, warmup: int = 2,
                 overlap: int = 1):
        self.gauss, self.views, self.dL_dout, self.grads, self.R_cap = gauss, views, dL_dout, grads, R_cap
        V = int(views.viewmatrix.shape[0])
        self.overlap = max(1, min(int(overlap), V))
        self.partials = [PackedGrads(grads.P, grads.M, device=grads.flat.device) for _ in range(self.overlap - 1)]
        # the graph bakes in raw pointers: its scratch is its own (one buffer per view group), never the
        # shared per-stream workspace, which another call may regrow (= free) later
        P = int(gauss["means3D"].shape[0])
        H, W = int(views.image_height), int(views.image_width)
        bounds = [shard_views(V, g, self.overlap) for g in range(self.overlap)]
        caps = list(R_cap) if isinstance(R_cap, (list, tuple)) else [R_cap] * self.overlap
        if any(c is None for c in caps):
            raise ValueError("GraphedFitStep needs a fixed R_cap (per view group)")
        self.temps = [torch.empty(step_temp_bytes(P, len(b), H, W, grads.M, sh_degree, c), dtype=torch.uint8,
                                  device=grads.flat.device) for b, c in zip(bounds, caps)]
        kw = dict(group=group, sh_degree=sh_degree, scale_modifier=scale_modifier, R_cap=self.R_cap, check="none",
                  overlap=self.overlap, partials=self.partials, temps=self.temps)
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                fit_step_grads(gauss, views, dL_dout, grads, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fit_step_grads(gauss, views, dL_dout, grads, **kw)
        self._pin = torch.zeros(4, dtype=torch.int64).pin_memory()

    def replay(self):
        self.graph.replay()
        return self.result

    def states(self):
        """State blobs of the captured step (one per view group)."""
        r = self.result
        return [x.state for x in r.results] if isinstance(r, StepResult) else [r.state]

    def status(self):
        """(R, overflow) of the last replay, summed / or-ed over view groups; synchronises the current stream."""
        from . import _native as N
        s = torch.cuda.current_stream()
        R, ov = 0, 0
        for st in self.states():
            N.check(N.lib().ghr_read_status_async(st.data_ptr(), self._pin.data_ptr(), s.cuda_stream),
                    "ghr_read_status_async")
            s.synchronize()
            R += int(self._pin[0])
            ov |= int(self._pin[1]) & 0xFFFFFFFF
        return R, ov


class FitResult(NamedTuple):
    """One step of PipelinedFitLoop.  `grads` is the slot's pinned host buffer (PackedGrads layout): valid until
    the slot is reused, `slots` steps later -- copy what must live longer."""
    grads: torch.Tensor
    loss: float
    R: int
    overflow: int


class PipelinedFitLoop:
    """Host buffers in, host buffers out, at (almost) the device-resident rate of GraphedFitStep.

    `slots` GraphedFitStep instances with their own static inputs and result buffers rotate.  Per step: the step's
    Gaussian attributes and cameras are copied from PINNED host memory into the slot's static inputs on an upload
    stream, `slots - 1` steps ahead of their replay; the replay (forward + backward [+ all-reduce], one graph
    launch) follows the previous one back to back on the launch stream; the loss <color, dL_dout>, the status
    reads and the download of the packed gradients run on a high-priority download stream; the host reads a
    step's results `slots - 1` launches after it (an overflow of the instance capacity raises there).

        loop = PipelinedFitLoop(gauss, views, dL_dout, R_cap, overlap=2)
        for res in loop.run((loop.pack_attributes(g), loop.pack_cameras(v)) for g, v in steps):
            optimizer_consumes(res.grads, res.loss)

    gauss / views: device tensors that fix names, shapes and dtypes of a step's inputs (and their initial values).
    Two slots are enough when a GPU's uploads and downloads overlap each other; the third absorbs the jitter where
    they share one copy path (bench.py's e2e leg at 8 x B200: 128k -> 141k views/s)."""

    CAM_FIELDS = ("viewmatrix", "projmatrix", "campos", "tanfov")

    def __init__(self, gauss: Dict[str, torch.Tensor], views, dL_dout: torch.Tensor, R_cap, overlap: int = 2,
                 slots: int = 3, peer: bool = False, group=None, sh_degree: int = 0, scale_modifier: float = 1.0,
                 trace: bool = False):
        from . import api
        dev = dL_dout.device
        self.names = list(gauss.keys())
        self.shapes = {k: tuple(gauss[k].shape) for k in self.names}
        self.sizes = {k: int(gauss[k].numel()) for k in self.names}
        self.cam_shapes = {k: tuple(getattr(views, k).shape) for k in self.CAM_FIELDS}
        self.cam_sizes = {k: int(getattr(views, k).numel()) for k in self.CAM_FIELDS}
        n_attr, n_cam = sum(self.sizes.values()), sum(self.cam_sizes.values())
        P = int(gauss["means3D"].shape[0])
        M = int(gauss["shs"].shape[1]) if gauss.get("shs") is not None else 0
        self.slots, self.trace = max(1, int(slots)), trace
        self.main = torch.cuda.current_stream(dev)
        self.up, self.down = torch.cuda.Stream(dev), torch.cuda.Stream(dev, priority=-1)
        self.dL_flat = dL_dout.reshape(-1)
        self._slots = []
        for _ in range(self.slots):
            flat = torch.empty(n_attr, dtype=torch.float32, device=dev)
            camflat = torch.empty(n_cam, dtype=torch.float32, device=dev)
            flat.copy_(torch.cat([gauss[k].detach().reshape(-1).float() for k in self.names]))
            camflat.copy_(torch.cat([getattr(views, k).reshape(-1).float() for k in self.CAM_FIELDS]))
            gin = self._carve(flat, self.names, self.sizes, self.shapes)
            cin = self._carve(camflat, self.CAM_FIELDS, self.cam_sizes, self.cam_shapes)
            sviews = api.ViewBatch(image_height=views.image_height, image_width=views.image_width,
                                   viewmatrix=cin["viewmatrix"], projmatrix=cin["projmatrix"], campos=cin["campos"],
                                   tanfov=cin["tanfov"], bg=views.bg, sh_degree=views.sh_degree,
                                   scale_modifier=views.scale_modifier)
            grads = PackedGrads(P, M, device=dev, peer=peer, group=group)
            step = GraphedFitStep(gin, sviews, dL_dout, grads, R_cap=R_cap, group=group, sh_degree=sh_degree,
                                  scale_modifier=scale_modifier, overlap=overlap)
            n_states = len(step.states())
            self._slots.append(dict(
                flat=flat, camflat=camflat, step=step, grads=grads,
                host_grads=torch.empty(grads.flat.numel()).pin_memory(), host_loss=torch.zeros(1).pin_memory(),
                host_status=torch.zeros(n_states, 4, dtype=torch.int64).pin_memory(),
                ev_up=torch.cuda.Event(), ev_used=torch.cuda.Event(), ev_down=torch.cuda.Event()))
        self.h2d_bytes_per_step = (n_attr + n_cam) * 4
        self.d2h_bytes_per_step = self._slots[0]["host_grads"].numel() * 4 + 4
        self.events = {"h2d": [], "replay": [], "d2h": []}      # trace=True: (start, stop) CUDA events per step
        self.host_ms = {"upload": [], "launch": [], "result": []}
        self.reset()

    @staticmethod
    def _carve(flat, names, sizes, shapes):
        out, o = {}, 0
        for k in names:
            out[k] = flat[o:o + sizes[k]].view(shapes[k])
            o += sizes[k]
        return out

    def pack_attributes(self, gauss: Dict[str, torch.Tensor]) -> torch.Tensor:
        """One pinned host buffer with a step's Gaussian attributes in this loop's order."""
        return torch.cat([gauss[k].detach().reshape(-1).float().cpu() for k in self.names]).pin_memory()

    def pack_cameras(self, views) -> torch.Tensor:
        """One pinned host buffer with a step's cameras (viewmatrix | projmatrix | campos | tanfov)."""
        return torch.cat([getattr(views, k).detach().reshape(-1).float().cpu() for k in self.CAM_FIELDS]).pin_memory()

    def reset(self):
        """Forget the slots' history (and the trace)."""
        for sl in self._slots:
            sl["ev_used"].record(self.main)
            sl["ev_down"].record(self.main)
        for v in list(self.events.values()) + list(self.host_ms.values()):
            del v[:]

    def _timed(self, what, fn, *a):
        if not self.trace:
            return fn(*a)
        import time
        t0 = time.perf_counter()
        r = fn(*a)
        self.host_ms[what].append((time.perf_counter() - t0) * 1e3)
        return r

    def _pair(self, stream):
        if not self.trace:
            return None, None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        return a, b

    def upload(self, i: int, host_attrs: torch.Tensor, host_cams: torch.Tensor):
        """Stage step i's inputs into its slot (waits on the device for the slot's previous replay)."""
        sl = self._slots[i % self.slots]
        with torch.cuda.stream(self.up):
            self.up.wait_event(sl["ev_used"])
            a, b = self._pair(self.up)
            sl["flat"].copy_(host_attrs, non_blocking=True)
            sl["camflat"].copy_(host_cams, non_blocking=True)
            if a is not None:
                b.record(self.up)
                self.events["h2d"].append((a, b))
            sl["ev_up"].record(self.up)

    def launch(self, i: int):
        """Replay step i and enqueue its loss, status reads and downloads."""
        from . import _native as N
        sl = self._slots[i % self.slots]
        main, down = self.main, self.down
        main.wait_event(sl["ev_up"])
        main.wait_event(sl["ev_down"])              # the slot's previous results have left the device
        a, b = self._pair(main)
        res = sl["step"].replay()
        if a is not None:
            b.record(main)
            self.events["replay"].append((a, b))
        sl["ev_used"].record(main)                  # the slot's inputs have been consumed
        down.wait_stream(main)
        with torch.cuda.stream(down):
            loss = torch.vdot(res.color.reshape(-1), self.dL_flat)
            for j, st in enumerate(sl["step"].states()):
                N.check(N.lib().ghr_read_status_async(st.data_ptr(), sl["host_status"][j].data_ptr(),
                                                      down.cuda_stream), "ghr_read_status_async")
            a, b = self._pair(down)
            sl["host_grads"].copy_(sl["grads"].flat, non_blocking=True)
            sl["host_loss"].copy_(loss.reshape(1), non_blocking=True)
            if a is not None:
                b.record(down)
                self.events["d2h"].append((a, b))
            sl["ev_down"].record(down)

    def result(self, i: int) -> FitResult:
        """Block until step i's results are in host memory."""
        sl = self._slots[i % self.slots]
        sl["ev_down"].synchronize()
        st = sl["host_status"]
        overflow = int((st[:, 1] & 0xFFFFFFFF).sum())
        if overflow:
            raise RuntimeError("PipelinedFitLoop: a step exceeded its instance capacity R_cap (its results are invalid)")
        return FitResult(sl["host_grads"], float(sl["host_loss"][0]), int(st[:, 0].sum()), overflow)

    def run(self, inputs):
        """inputs: iterable of (pinned attributes, pinned cameras) per step (pack_attributes / pack_cameras).
        Yields one FitResult per step, in order, `slots - 1` launches behind the newest one."""
        depth = self.slots - 1
        it = iter(inputs)
        n_up = n_launch = n_out = 0
        pending = []                                   # uploaded, not yet launched (kept alive: async copies)
        done = False

        def pull():
            nonlocal n_up, done
            if done:
                return False
            try:
                a, c = next(it)
            except StopIteration:
                done = True
                return False
            self._timed("upload", self.upload, n_up, a, c)
            pending.append((a, c))
            n_up += 1
            return True

        for _ in range(max(depth, 1)):
            pull()
        while n_launch < n_up:
            if depth:
                pull()                                 # step n_launch + depth goes up before step n_launch launches
            self._timed("launch", self.launch, n_launch)
            n_launch += 1
            if len(pending) > self.slots:
                pending.pop(0)
            if not depth:
                yield self._timed("result", self.result, n_out)
                n_out += 1
                pull()
            elif n_launch - n_out > depth:
                yield self._timed("result", self.result, n_out)
                n_out += 1
        while n_out < n_launch:
            yield self._timed("result", self.result, n_out)
            n_out += 1
