"""Host side of the rasterizer: the `diff_gaussian_rasterization` surface GuassianHand calls,
plus a multi-view batched entry.

Mirrors (names, argument meaning, error behaviour) the third-party module the reference imports
at /root/reference/tgs/models/renderer_one_shot.py:3 and uses at :281-296, :338-346, :355-379:

    GaussianRasterizationSettings  -- 12-field NamedTuple, same field order
    GaussianRasterizer(raster_settings).forward(means3D, means2D, opacities, shs, colors_precomp,
                                                scales, rotations, cov3D_precomp) -> (color, radii)
    GaussianRasterizer.markVisible(positions) -> bool[P]

PyTorch is plumbing here: it owns device memory, the stream and autograd bookkeeping.  All math
runs in libghr.so (hand-written sm_100a CUDA behind the C ABI of include/ghr.h).  There is no CPU
or eager fallback: without the library, or without CUDA tensors, calls raise.
"""
import ctypes as C
import itertools
import threading
import time
import weakref
from typing import NamedTuple, Optional

import torch
from torch import nn

from . import _native as N


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# ------------------------------------------------------------------ workspaces

class _Workspace:
    """Per (device, stream) scratch: temp buffer, pinned status slots, instance-capacity policy.

    Upstream sizes its binning buffers after a blocking D2H read of num_rendered in the middle of
    every forward.  Here the buffers are sized from a running high-water mark and
    the library reports R to pinned host memory as soon as the scan finishes; the host polls that
    word while the GPU is already sorting and blending.  Overflow (R > capacity) re-runs the
    forward once with a larger capacity, so results are always exact.
    """

    def __init__(self, device):
        self.device = device
        self.temp = torch.empty(0, dtype=torch.uint8, device=device)
        self.cap = {}          # (P, V, H, W) -> entries
        self.hist = {}         # (P, V, H, W) -> [exact instance counts seen, their maximum]
        self.pinned = torch.zeros(64, 4, dtype=torch.int64).pin_memory()   # 64 slots of GhrStatus
        self.pinned_np = self.pinned.numpy()      # same memory: reading a status word costs no tensor op
        self.pinned_ptr = self.pinned.data_ptr()
        self.slot = 0
        self.pending = []      # deferred checks: (row, seq, key, cap, fixed_cap)

    def get_temp(self, nbytes: int) -> torch.Tensor:
        if self.temp.numel() < nbytes:
            self.temp = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=self.device)
        return self.temp

    def capacity(self, key, P, V) -> int:
        c = self.cap.get(key)
        if c is None:
            c = max(8 * P * V, 1 << 16)
            self.cap[key] = c
        return c

    def note(self, key, R):
        h = self.hist.setdefault(key, [0, 0])
        h[0] += 1
        h[1] = max(h[1], int(R))

    def stable(self, key, cap) -> bool:
        """check="auto": the capacity has covered several exact counts with room to spare, so the host need not
        wait for this call's count (it is verified later; an overflow then raises)."""
        h = self.hist.get(key)
        return h is not None and h[0] >= 3 and cap >= int(h[1] * 1.3) + 4096

    def next_slot(self):
        """Returns (numpy row view [4] int64, host address) of the next pinned GhrStatus slot."""
        self.slot = (self.slot + 1) % self.pinned_np.shape[0]
        if any(p[5] == self.slot for p in self.pending):
            self.verify_pending(wait=True)          # the ring wrapped around an unverified report
        row = self.pinned_np[self.slot]
        row[:] = 0
        return row, self.pinned_ptr + 32 * self.slot

    def verify_pending(self, wait: bool = False):
        """check="deferred": look at the status reports of earlier forwards that have arrived (all of
        them when `wait`); raises if one overflowed its capacity (its outputs were invalid)."""
        keep = []
        for row, seq, key, cap, fixed, slot in self.pending:
            if int(row[2]) != seq:
                if not wait:
                    keep.append((row, seq, key, cap, fixed, slot))
                    continue
                torch.cuda.synchronize(self.device)
                if int(row[2]) != seq:
                    raise RuntimeError("ghr_forward: status report never arrived")
            R, overflow = int(row[0]), int(row[1]) & 0xFFFFFFFF
            if overflow & N.GHR_STATUS_PREFILTER:
                self.pending = []
                raise RuntimeError(_PREFILTER_MSG)
            self.note(key, R)
            if not fixed and key in self.cap:
                self.cap[key] = max(self.cap[key], int(R * 1.5) + (1 << 14))
            if overflow:
                self.pending = []
                raise RuntimeError(f"ghr_forward (deferred check): an earlier forward produced {R} instances, more "
                                   f"than its capacity {cap}; its outputs were invalid. Re-run it (the capacity "
                                   f"has been raised) or use check='poll'.")
        self.pending = keep


# upstream's in_frustum() prints this and __trap()s (the CUDA context is lost); here the forward raises and the
# context survives
_PREFILTER_MSG = "Point is filtered although prefiltered is set. This shouldn't happen!"

_ws_lock = threading.Lock()
_workspaces = {}
_seq = itertools.count(1)


def _raw_stream(dev) -> int:
    """cudaStream_t of torch's current stream on `dev` (no Stream object is built)."""
    return torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device())


_layouts = {}
_grad_plans = {}     # (P, M, scales/rotations present) -> carving of backward_raw's gradient allocation


def _layout(P, V, H, W, M, sh_degree, cap):
    key = (P, V, H, W, M, sh_degree, cap)
    lay = _layouts.get(key)
    if lay is None:
        if len(_layouts) > 256:
            _layouts.clear()
        lay = _layouts[key] = N.layout(*key)
    return lay


def _workspace(device, stream) -> _Workspace:
    raw = stream if isinstance(stream, int) else stream.cuda_stream
    key = (device.index if device.index is not None else torch.cuda.current_device(), raw)
    with _ws_lock:
        ws = _workspaces.get(key)
        if ws is None:
            ws = _Workspace(torch.device("cuda", key[0]))
            _workspaces[key] = ws
        return ws


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None or t.numel() == 0 else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _f32a(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """_f32c plus 16-byte alignment: the kernels read rotations (and SH rows) with 128-bit loads; a contiguous
    view into a packed parameter buffer (flat[6P:10P].view(P, 4) with odd P) is only 4-byte aligned."""
    if t is None:
        return None
    t = _f32c(t)
    if t.numel() and t.data_ptr() % 16:
        t = t.clone()
    return t


def _require_cuda(*ts):
    for t in ts:
        if t is not None and t.numel() and not t.is_cuda:
            raise RuntimeError("guassianhand_b200: all tensors must be CUDA tensors (there is no CPU path)")


class _Cams(NamedTuple):
    V: int
    H: int
    W: int
    view: torch.Tensor        # [V,16]
    proj: torch.Tensor        # [V,16]
    campos: torch.Tensor      # [V,3]
    tanfov: Optional[torch.Tensor]   # [V,2] device, or None -> scalars
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor          # [3] or [V,3]
    bg_stride: int


def _fill_common(a, cams: _Cams, P, M, sh_degree, R_cap, scale_modifier, flags, means3D, opacities, scales,
                 rotations, cov3D, shs, colors):
    a.dims = N.GhrDims(P, cams.V, cams.H, cams.W, M, sh_degree, R_cap)
    a.flags = flags
    a.scale_modifier = scale_modifier
    a.tanfovx, a.tanfovy = cams.tanfovx, cams.tanfovy
    a.viewmatrix, a.projmatrix, a.campos = cams.view.data_ptr(), cams.proj.data_ptr(), cams.campos.data_ptr()
    a.tanfov = _ptr(cams.tanfov)
    a.bg, a.bg_stride = cams.bg.data_ptr(), cams.bg_stride
    a.means3D, a.opacities = _ptr(means3D), _ptr(opacities)
    a.scales, a.rotations, a.cov3D_precomp = _ptr(scales), _ptr(rotations), _ptr(cov3D)
    a.shs, a.colors_precomp = _ptr(shs), _ptr(colors)


class ForwardResult(NamedTuple):
    color: torch.Tensor       # [V,3,H,W]
    mask: Optional[torch.Tensor]   # [V,H,W] coverage 1 - T_final (want_mask) or None
    radii: torch.Tensor       # [V,P] int32
    state: torch.Tensor       # uint8 blob (geometry, sorted instances, ranges, final_T, n_contrib)
    R_cap: int
    R: Optional[int]          # known when check == "poll"
    debug: Optional[dict]


def forward_raw(cams: _Cams, means3D, opacities, scales, rotations, cov3D, shs, colors, sh_degree,
                scale_modifier, flags=0, check: str = "poll", want_debug: bool = False,
                R_cap: Optional[int] = None, stage_events=None, want_mask: bool = False, out=None,
                temp: Optional[torch.Tensor] = None, reuse=None) -> ForwardResult:
    """Enqueue one libghr forward (V views).  check: "poll" (exact: the host waits for the instance
    count, which arrives while the GPU is still sorting/blending, and re-runs on overflow),
    "deferred" (no host wait: the report is verified at the next call on this stream or by
    check_deferred(); an overflow raises there), "none" (caller checks GhrStatus itself; needed
    under CUDA-graph capture).  temp: caller-owned scratch (uint8, >= layout temp_bytes) instead of the
    per-(device, stream) workspace -- required for anything whose pointers outlive the call (CUDA graphs).
    check="auto": "deferred" once the capacity has proven itself on this shape (_Workspace.stable), else "poll".
    reuse=(state, M, R): geometry reuse (GhrForwardArgs.reuse_state) -- the state blob of an earlier forward
    with identical geometry inputs, camera and R_cap (V == 1); only colours are recomputed and the blend runs."""
    L = N.lib()
    dev = means3D.device
    stream = _raw_stream(dev)
    ws = _workspace(dev, stream)
    P = means3D.shape[0]
    M = 0 if shs is None or shs.numel() == 0 else shs.shape[1]
    key = (P, cams.V, cams.H, cams.W)
    if ws.pending:
        ws.verify_pending()
    cap = int(R_cap) if R_cap is not None else ws.capacity(key, P, cams.V)
    if reuse is not None:
        check = "none"                 # the instance count and the capacity are those of the earlier call
    elif check == "auto":
        check = "deferred" if ws.stable(key, cap) else "poll"
    while True:
        lay = _layout(P, cams.V, cams.H, cams.W, M, sh_degree, cap)
        state = torch.empty(lay.state_bytes, dtype=torch.uint8, device=dev)
        if temp is None:
            tmp = ws.get_temp(max(lay.temp_bytes, lay.temp_bwd_bytes))
        else:
            tmp = temp
            if tmp.numel() < lay.temp_bytes:
                raise RuntimeError(f"forward_raw: temp has {tmp.numel()} bytes, the layout needs {lay.temp_bytes}")
        if out is not None:
            # caller-owned outputs (contiguous [V,3,H,W] / [V,max(P,1)] / [V,H,W] blocks of a larger batch)
            color, radii, mask = out
        else:
            color = torch.empty(cams.V, 3, cams.H, cams.W, dtype=torch.float32, device=dev)
            radii = torch.empty(cams.V, max(P, 1), dtype=torch.int32, device=dev)
            mask = torch.empty(cams.V, cams.H, cams.W, dtype=torch.float32, device=dev) if want_mask else None
        dbg = None
        a = N.GhrForwardArgs()
        _fill_common(a, cams, P, M, sh_degree, cap, scale_modifier, flags, means3D, opacities, scales, rotations,
                     cov3D, shs, colors)
        a.out_color, a.radii = color.data_ptr(), radii.data_ptr()
        a.out_mask = _ptr(mask)
        a.state, a.state_bytes = state.data_ptr(), lay.state_bytes
        a.temp, a.temp_bytes = tmp.data_ptr(), tmp.numel()
        if want_debug:
            dbg = dict(keys=torch.zeros(max(cap, 1), dtype=torch.int64, device=dev),
                       point_list=torch.zeros(max(cap, 1), dtype=torch.int32, device=dev), layout=lay)
            a.dbg_keys_sorted, a.dbg_point_list = dbg["keys"].data_ptr(), dbg["point_list"].data_ptr()
        if stage_events is not None:
            a.stage_events = stage_events.ptr()
        if reuse is not None:
            a.reuse_state, a.reuse_M = reuse[0].data_ptr(), int(reuse[1])
        row = None
        seq = next(_seq)
        a.seq = seq
        if check in ("poll", "deferred"):
            row, a.host_status = ws.next_slot()
        N.check(L.ghr_forward(C.byref(a), stream), "ghr_forward")
        R = None if reuse is None else reuse[2]
        if check == "deferred":
            ws.pending.append((row, seq, key, cap, R_cap is not None, ws.slot))
        if check == "poll":
            t0 = time.perf_counter()
            while int(row[2]) != seq:           # reserved[0]
                if time.perf_counter() - t0 > 10.0:
                    torch.cuda.current_stream(dev).synchronize()
                    if int(row[2]) != seq:
                        raise RuntimeError("ghr_forward: status report never arrived")
            R = int(row[0])
            overflow = int(row[1]) & 0xFFFFFFFF
            if overflow & N.GHR_STATUS_PREFILTER:
                raise RuntimeError(_PREFILTER_MSG)
            ws.note(key, R)
            if R_cap is None:
                ws.cap[key] = max(ws.cap[key], int(R * 1.5) + (1 << 14))
            if overflow:
                if R_cap is not None:
                    raise RuntimeError(f"ghr_forward: {R} instances exceed the fixed capacity R_cap={cap}")
                cap = ws.cap[key]
                continue
        return ForwardResult(color, mask, radii[:, :P], state, cap, R, dbg)


def check_deferred(device=None):
    """Verify every outstanding check="deferred" forward on `device` (synchronises it)."""
    with _ws_lock:
        wss = list(_workspaces.items())
    for (dev_index, _), ws in wss:
        if device is None or torch.device(device).index in (None, dev_index):
            ws.verify_pending(wait=True)


def backward_raw(cams: _Cams, fwd_state, R_cap, dL_dout, means3D, opacities, scales, rotations, cov3D, shs,
                 colors, sh_degree, scale_modifier, flags=0, want_means2D=True, accumulate_into=None,
                 want_conic=False, accumulate=True, stage_events=None, dL_dmask=None, temp=None):
    """Enqueue one libghr backward.  Returns dict of gradient tensors (summed over views, except
    dL_dmeans2D which is per view)."""
    L = N.lib()
    dev = means3D.device
    stream = _raw_stream(dev)
    ws = _workspace(dev, stream)
    P = means3D.shape[0]
    M = 0 if shs is None or shs.numel() == 0 else shs.shape[1]
    lay = _layout(P, cams.V, cams.H, cams.W, M, sh_degree, R_cap)
    if ws.pending:
        ws.verify_pending()            # a deferred capacity report that has arrived (raises on overflow)
    if temp is None:
        temp = ws.get_temp(max(lay.temp_bytes, lay.temp_bwd_bytes))
    elif temp.numel() < lay.temp_bwd_bytes:
        raise RuntimeError(f"backward_raw: temp has {temp.numel()} bytes, the layout needs {lay.temp_bwd_bytes}")
    a = N.GhrBackwardArgs()
    _fill_common(a, cams, P, M, sh_degree, R_cap, scale_modifier, flags, means3D, opacities, scales, rotations, cov3D,
                 shs, colors)
    dL_dout = _f32c(dL_dout)
    a.dL_dout_color = dL_dout.data_ptr()
    if dL_dmask is not None:
        dL_dmask = _f32c(dL_dmask)
        a.dL_dout_mask = dL_dmask.data_ptr()
    a.state, a.state_bytes = fwd_state.data_ptr(), fwd_state.numel()
    a.temp, a.temp_bytes = temp.data_ptr(), temp.numel()
    g = None if accumulate_into is None else dict(accumulate_into)
    a.accumulate = 1 if (g is not None and accumulate) else 0
    if g is None:
        # one allocation, carved into the per-attribute gradients (each block 16-byte aligned); the carving plan
        # is cached per shape
        has_sr = scales is not None and scales.numel() > 0
        plan = _grad_plans.get((P, M, has_sr))
        if plan is None:
            shapes = [("dL_dmeans3D", (P, 3)), ("dL_dopacity", (P, 1)), ("dL_dcov3D", (P, 6))]
            shapes.append(("dL_dsh", (P, M, 3)) if M > 0 else ("dL_dcolors", (P, 3)))
            if has_sr:
                shapes += [("dL_dscales", (P, 3)), ("dL_drotations", (P, 4))]
            items, o = [], 0
            for name, sh in shapes:
                n = int(torch.Size(sh).numel())
                items.append((name, sh, o, n))
                o += (n + 3) // 4 * 4
            if len(_grad_plans) > 64:
                _grad_plans.clear()
            plan = _grad_plans[(P, M, has_sr)] = (items, o)
        flat = torch.empty(plan[1], dtype=torch.float32, device=dev)
        g = {"_flat": flat}
        base = flat.data_ptr()
        for name, sh, o, n in plan[0]:
            g[name] = flat[o:o + n].view(sh)
            setattr(a, name, base + 4 * o if n else None)
    else:
        for k in ("dL_dmeans3D", "dL_dcolors", "dL_dopacity", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations"):
            setattr(a, k, _ptr(g.get(k)))
    if want_means2D:
        g["dL_dmeans2D"] = torch.empty(cams.V, P, 3, dtype=torch.float32, device=dev)
        a.dL_dmeans2D = _ptr(g["dL_dmeans2D"])
    elif g.get("dL_dmeans2D") is not None:
        a.dL_dmeans2D = _ptr(g["dL_dmeans2D"])
    if want_conic:
        g["dL_dconic"] = torch.empty(cams.V, P, 4, dtype=torch.float32, device=dev)
        a.dL_dconic = _ptr(g["dL_dconic"])
    elif g.get("dL_dconic") is not None:
        a.dL_dconic = _ptr(g["dL_dconic"])
    if stage_events is not None:
        a.stage_events = stage_events.ptr()
    N.check(L.ghr_backward(C.byref(a), stream), "ghr_backward")
    return g


def _cams_from_settings(rs: GaussianRasterizationSettings) -> _Cams:
    _require_cuda(rs.viewmatrix, rs.projmatrix, rs.campos, rs.bg)
    return _Cams(V=1, H=int(rs.image_height), W=int(rs.image_width),
                 view=_f32c(rs.viewmatrix).view(1, 16), proj=_f32c(rs.projmatrix).view(1, 16),
                 campos=_f32c(rs.campos).view(1, 3), tanfov=None, tanfovx=float(rs.tanfovx),
                 tanfovy=float(rs.tanfovy), bg=_f32c(rs.bg).view(3), bg_stride=0)


def _flags(rs) -> int:
    return (N.GHR_FLAG_PREFILTERED if rs.prefiltered else 0) | (N.GHR_FLAG_DEBUG if rs.debug else 0)


def _opt(t):
    return None if t is None or t.numel() == 0 else _f32a(t)


class _GeomCache:
    """The last single-view forward per device, for the reference's call pattern: every view is rendered twice
    with IDENTICAL geometry (RGB, then an all-ones mask render; renderer_one_shot.py:338-346, :372-379).  The
    second call is recognised by the IDENTITY and the autograd version counters of the geometry and camera
    tensors (the entry keeps them alive, so an address cannot be recycled into a false hit) and reuses the
    first call's projected geometry, tile ranges, depth order and cull masks (GhrForwardArgs.reuse_state)."""
    entries = {}
    hits = 0

    @staticmethod
    def _ident(t):
        # (storage address, shape, strides, autograd version); inference tensors have no version: never cached
        return None if t is None else (t.data_ptr(), tuple(t.shape), t.stride(), t._version)

    @classmethod
    def lookup(cls, dev_index, tensors, scalars):
        e = cls.entries.get(dev_index)
        if e is None:
            return None
        _, idents, es, state_ref, M, R_cap, R = e
        try:
            if es != scalars or idents != tuple(cls._ident(t) for t in tensors):
                return None
        except RuntimeError:
            return None
        state = state_ref()
        if state is None:
            return None
        cls.hits += 1
        return state, M, R_cap, R

    @classmethod
    def store(cls, dev_index, tensors, scalars, state, M, R_cap, R):
        try:
            idents = tuple(cls._ident(t) for t in tensors)
        except RuntimeError:
            cls.entries.pop(dev_index, None)
            return
        # the entry holds the tensors: while it lives their addresses cannot be handed to other tensors
        cls.entries[dev_index] = (tensors, idents, scalars, weakref.ref(state), M, R_cap, R)


class _RasterizeGaussians(torch.autograd.Function):
    """Same argument order and gradient order as upstream's autograd node."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                raster_settings, want_mask=False):
        _require_cuda(means3D, opacities)
        if means3D.dim() != 2 or means3D.shape[1] != 3:
            raise RuntimeError("means3D must have dimensions (num_points, 3)")
        rs = raster_settings
        cams = _cams_from_settings(rs)
        means3D_c, opac_c = _f32c(means3D), _f32c(opacities)
        sh_c, col_c, sc_c, rot_c, cov_c = _opt(sh), _opt(colors_precomp), _opt(scales), _opt(rotations), _opt(cov3Ds_precomp)
        args = (cams, means3D_c, opac_c, sc_c, rot_c, cov_c, sh_c, col_c, int(rs.sh_degree), float(rs.scale_modifier))
        M = 0 if sh_c is None else sh_c.shape[1]
        dev_index = means3D_c.device.index
        # (sc_c / rot_c / cov_c are None exactly when the argument is absent or empty)
        gtensors = (means3D, opacities, scales if sc_c is not None else None, rotations if rot_c is not None else None,
                    cov3Ds_precomp if cov_c is not None else None, rs.viewmatrix, rs.projmatrix, rs.campos)
        gscalars = (means3D.shape[0], cams.H, cams.W, cams.tanfovx, cams.tanfovy, float(rs.scale_modifier), _flags(rs))
        hit = None if rs.debug else _GeomCache.lookup(dev_index, gtensors, gscalars)
        if hit is not None:
            res = forward_raw(*args, flags=_flags(rs), want_mask=want_mask, R_cap=hit[2], reuse=(hit[0], hit[1], hit[3]))
        elif rs.debug:
            cpu_args = [None if t is None or not torch.is_tensor(t) else t.detach().cpu().clone()
                        for t in (means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp)]
            try:
                res = forward_raw(*args, flags=_flags(rs), want_mask=want_mask)
            except Exception:
                torch.save(cpu_args, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            res = forward_raw(*args, flags=_flags(rs), want_mask=want_mask, check="poll" if rs.prefiltered else "auto")
            _GeomCache.store(dev_index, gtensors, gscalars, res.state, M, res.R_cap, res.R)
        ctx.raster_settings = rs
        ctx.want_mask = want_mask
        ctx.cams = cams
        ctx.R_cap = res.R_cap
        ctx.num_rendered = res.R
        ctx.save_for_backward(means3D_c, opac_c, sc_c, rot_c, cov_c, sh_c, col_c, res.state)
        radii = res.radii[0]
        ctx.mark_non_differentiable(radii)
        if want_mask:
            return res.color[0], radii, res.mask[0]
        return res.color[0], radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii, grad_mask=None):
        rs = ctx.raster_settings
        means3D, opac, sc, rot, cov, sh, col, state = ctx.saved_tensors
        call = lambda: backward_raw(ctx.cams, state, ctx.R_cap, grad_out_color.unsqueeze(0), means3D, opac, sc, rot,
                                    cov, sh, col, int(rs.sh_degree), float(rs.scale_modifier), flags=_flags(rs),
                                    dL_dmask=grad_mask.unsqueeze(0) if (ctx.want_mask and grad_mask is not None) else None)
        if rs.debug:
            try:
                g = call()
            except Exception:
                torch.save([None if t is None else t.detach().cpu() for t in ctx.saved_tensors[:-1]] + [grad_out_color.detach().cpu()],
                           "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        else:
            g = call()
        P = means3D.shape[0]
        need = ctx.needs_input_grad
        return (
            g["dL_dmeans3D"] if need[0] else None,
            g["dL_dmeans2D"][0] if need[1] else None,
            g.get("dL_dsh") if need[2] else None,
            g.get("dL_dcolors") if need[3] else None,
            g["dL_dopacity"].view(P, -1) if need[4] else None,
            g.get("dL_dscales") if need[5] else None,
            g.get("dL_drotations") if need[6] else None,
            g["dL_dcov3D"] if need[7] else None,
            None,
            None,
        )


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                        raster_settings, want_mask=False):
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations,
                                     cov3Ds_precomp, raster_settings, want_mask)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            rs = self.raster_settings
            _require_cuda(positions, rs.viewmatrix)
            pos = _f32c(positions)
            P = pos.shape[0]
            out = torch.zeros(P, dtype=torch.bool, device=pos.device)
            stream = torch.cuda.current_stream(pos.device)
            N.check(N.lib().ghr_mark_visible(P, _ptr(pos), _f32c(rs.viewmatrix).data_ptr(),
                                             _f32c(rs.projmatrix).data_ptr(), _ptr(out), stream.cuda_stream),
                    "ghr_mark_visible")
        return out

    def forward_with_mask(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                          rotations=None, cov3D_precomp=None):
        """forward() plus the coverage mask [H,W] = 1 - T_final of the same pass: replaces the second
        rasterizer call of renderer_one_shot.py:353-380 (colors = 1, bg = 0).  Returns
        (color [3,H,W], radii [P], mask [H,W]); all three renders' gradients flow through ONE backward."""
        return self.forward(means3D, means2D, opacities, shs, colors_precomp, scales, rotations, cov3D_precomp,
                            _want_mask=True)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, _want_mask=False):
        rs = self.raster_settings
        if (shs is None and colors_precomp is None) or (shs is not None and colors_precomp is not None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        if ((scales is None or rotations is None) and cov3D_precomp is None) or \
                ((scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        # (upstream substitutes torch.Tensor([]) for the absent arguments here; the node takes None directly --
        # three tensor constructions per call less on a host-bound path)
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, rs, _want_mask)


# ------------------------------------------------------------------ multi-view batched entry

class ViewBatch(NamedTuple):
    """V cameras with a shared image size: the per-view part of GaussianRasterizationSettings
    stacked along dim 0 (renderer_one_shot.py:494-503 loops over exactly these per view)."""
    image_height: int
    image_width: int
    viewmatrix: torch.Tensor     # [V,4,4]
    projmatrix: torch.Tensor     # [V,4,4]
    campos: torch.Tensor         # [V,3]
    tanfov: torch.Tensor         # [V,2]  (tanfovx, tanfovy)
    bg: torch.Tensor             # [3] or [V,3]
    sh_degree: int = 0
    scale_modifier: float = 1.0

    @staticmethod
    def from_w2c(w2c: torch.Tensor, intrinsics: torch.Tensor, image_height: int, image_width: int, bg: torch.Tensor,
                 sh_degree: int = 0, scale_modifier: float = 1.0, znear: float = 0.01, zfar: float = 1000.0) -> "ViewBatch":
        """V cameras from world-to-camera matrices [V,4,4] and intrinsics [V,3,3] (CUDA), built by ONE kernel
        of libghr with no host synchronisation: the device-side replacement of the reference's per-view
        Camera.from_w2c + getProjectionMatrix_refine + intrinsic_to_fov and of the two math.tan(cuda scalar)
        syncs (/root/reference/tgs/models/renderer_one_shot.py:61-112, :278-279; its view loop :494-503).
        The reference ignores the znear / zfar it is given and uses 0.01 / 1000 (:99-100): the defaults."""
        _require_cuda(w2c, intrinsics)
        w, k = _f32c(w2c).view(-1, 16), _f32c(intrinsics).view(-1, 9)
        V = w.shape[0]
        if k.shape[0] != V:
            raise RuntimeError("ViewBatch.from_w2c: w2c and intrinsics must hold the same number of views")
        dev = w.device
        view = torch.empty(V, 4, 4, dtype=torch.float32, device=dev)
        proj = torch.empty(V, 4, 4, dtype=torch.float32, device=dev)
        campos = torch.empty(V, 3, dtype=torch.float32, device=dev)
        tanfov = torch.empty(V, 2, dtype=torch.float32, device=dev)
        N.check(N.lib().ghr_cameras_from_w2c(V, w.data_ptr(), k.data_ptr(), int(image_height), int(image_width),
                                             float(znear), float(zfar), view.data_ptr(), proj.data_ptr(),
                                             campos.data_ptr(), tanfov.data_ptr(), _raw_stream(dev)),
                "ghr_cameras_from_w2c")
        return ViewBatch(image_height=int(image_height), image_width=int(image_width), viewmatrix=view, projmatrix=proj,
                         campos=campos, tanfov=tanfov, bg=bg, sh_degree=sh_degree, scale_modifier=scale_modifier)

    def cams(self) -> _Cams:
        V = self.viewmatrix.shape[0]
        bg = _f32c(self.bg)
        return _Cams(V=V, H=int(self.image_height), W=int(self.image_width), view=_f32c(self.viewmatrix).view(V, 16),
                     proj=_f32c(self.projmatrix).view(V, 16), campos=_f32c(self.campos).view(V, 3),
                     tanfov=_f32c(self.tanfov).view(V, 2), tanfovx=0.0, tanfovy=0.0, bg=bg,
                     bg_stride=3 if bg.dim() == 2 else 0)


_side_streams = {}


def _streams(dev, n):
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    pool = _side_streams.setdefault(idx, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=idx))
    return pool[:n]


def _groups(V: int, G: int):
    """Contiguous, balanced view blocks."""
    base, rem = divmod(V, G)
    out, lo = [], 0
    for g in range(G):
        hi = lo + base + (1 if g < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def _slice_cams(c: _Cams, lo: int, hi: int) -> _Cams:
    return c._replace(V=hi - lo, view=c.view[lo:hi], proj=c.proj[lo:hi], campos=c.campos[lo:hi],
                      tanfov=None if c.tanfov is None else c.tanfov[lo:hi],
                      bg=c.bg[lo:hi] if c.bg_stride else c.bg)


class _RasterizeViews(torch.autograd.Function):
    """V views, cut into `overlap` contiguous groups whose launch chains run on concurrent streams
    (group 0 on the current one): the chains are independent, so one group's single-CTA tile scan and
    kernel tails are filled with another group's work.  Outputs are blocks of one [V,...] tensor."""

    @staticmethod
    def forward(ctx, means3D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, views: ViewBatch,
                want_mask: bool, check: str = "poll", overlap: int = 1):
        _require_cuda(means3D, opacities, views.viewmatrix)
        cams = views.cams()
        means3D_c, opac_c = _f32c(means3D), _f32c(opacities)
        sh_c, col_c, sc_c, rot_c, cov_c = _opt(sh), _opt(colors_precomp), _opt(scales), _opt(rotations), _opt(cov3Ds_precomp)
        dev, V, P = means3D.device, cams.V, means3D.shape[0]
        G = max(1, min(int(overlap), V))
        color = torch.empty(V, 3, cams.H, cams.W, dtype=torch.float32, device=dev)
        radii = torch.empty(V, max(P, 1), dtype=torch.int32, device=dev)
        mask = torch.empty(V, cams.H, cams.W, dtype=torch.float32, device=dev) if want_mask else None
        main = torch.cuda.current_stream(dev)
        streams = [main] + _streams(dev, G - 1)
        for st in streams[1:]:
            st.wait_stream(main)
        groups, states, caps = _groups(V, G), [], []
        for g, (lo, hi) in enumerate(groups):
            with torch.cuda.stream(streams[g]):
                res = forward_raw(_slice_cams(cams, lo, hi), means3D_c, opac_c, sc_c, rot_c, cov_c, sh_c, col_c,
                                  int(views.sh_degree), float(views.scale_modifier), want_mask=want_mask, check=check,
                                  out=(color[lo:hi], radii[lo:hi], None if mask is None else mask[lo:hi]))
            states.append(res.state)
            caps.append(res.R_cap)
        for st in streams[1:]:
            main.wait_stream(st)
        ctx.cams, ctx.views, ctx.caps, ctx.want_mask, ctx.groups = cams, views, caps, want_mask, groups
        z = torch.empty(0)
        ctx.save_for_backward(means3D_c, opac_c, sc_c if sc_c is not None else z, rot_c if rot_c is not None else z,
                              cov_c if cov_c is not None else z, sh_c if sh_c is not None else z,
                              col_c if col_c is not None else z, *states)
        radii_out = radii[:, :P]
        ctx.mark_non_differentiable(radii_out)
        return color, (mask if want_mask else torch.empty(0, device=dev)), radii_out

    @staticmethod
    def backward(ctx, grad_color, grad_mask, _):
        means3D, opac, sc, rot, cov, sh, col, *states = ctx.saved_tensors
        nz = lambda t: None if t.numel() == 0 else t
        sc, rot, cov, sh, col = map(nz, (sc, rot, cov, sh, col))
        v, dev = ctx.views, means3D.device
        grad_color = _f32c(grad_color)
        grad_mask = _f32c(grad_mask) if ctx.want_mask else None
        G = len(states)
        main = torch.cuda.current_stream(dev)
        streams = [main] + _streams(dev, G - 1)
        for st in streams[1:]:
            st.wait_stream(main)
        parts = []
        for g, (lo, hi) in enumerate(ctx.groups):
            with torch.cuda.stream(streams[g]):
                parts.append(backward_raw(_slice_cams(ctx.cams, lo, hi), states[g], ctx.caps[g], grad_color[lo:hi],
                                          means3D, opac, sc, rot, cov, sh, col, int(v.sh_degree),
                                          float(v.scale_modifier), want_means2D=False,
                                          dL_dmask=None if grad_mask is None else grad_mask[lo:hi]))
        g = parts[0]
        for st, part in zip(streams[1:], parts[1:]):
            main.wait_stream(st)
            part["_flat"].record_stream(main)
            g["_flat"].add_(part["_flat"])        # every gradient is a block of one flat allocation
        need = ctx.needs_input_grad
        P = means3D.shape[0]
        return (g["dL_dmeans3D"] if need[0] else None, g.get("dL_dsh") if need[1] else None,
                g.get("dL_dcolors") if need[2] else None, g["dL_dopacity"].view(P, -1) if need[3] else None,
                g.get("dL_dscales") if need[4] else None, g.get("dL_drotations") if need[5] else None,
                g["dL_dcov3D"] if need[6] else None, None, None, None, None)


def rasterize_views(means3D, opacities, views: ViewBatch, shs=None, colors_precomp=None, scales=None,
                    rotations=None, cov3D_precomp=None, return_mask: bool = False, check: str = "poll",
                    overlap: Optional[int] = None):
    """Render V views of one Gaussian set in a single launch chain.
    Returns (color [V,3,H,W], radii [V,P]) -- or (color, mask [V,H,W], radii) with return_mask --
    differentiable w.r.t. the Gaussian attributes with the gradient summed over views (what the
    per-view Python loop + autograd of the reference yields).

    `mask` is the coverage 1 - T_final of the SAME pass: what the reference obtains from a second
    rasterizer call with colors = 1 and bg = 0 (renderer_one_shot.py:353-380), at no extra render.

    check="deferred" removes the one host wait of a call (for pipelined loops that keep several steps
    in flight): the instance-capacity report is verified at the next call / by check_deferred().

    overlap: number of contiguous view groups whose launch chains run on concurrent streams (default 1).
    Results are identical; only the schedule changes.  It pays when the GPU is the bottleneck (large
    batches, CUDA-graph replay: +4 % on 8 views of the C2 scene); an eager Python loop is host-bound at
    these sizes and the extra launches cost more than the overlap returns."""
    if (shs is None) == (colors_precomp is None):
        raise Exception('Please provide excatly one of either SHs or precomputed colors!')
    if ((scales is None or rotations is None) and cov3D_precomp is None) or \
            ((scales is not None or rotations is not None) and cov3D_precomp is not None):
        raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
    e = lambda t: torch.Tensor([]) if t is None else t
    V = int(views.viewmatrix.shape[0])
    if overlap is None:
        overlap = 1
    color, mask, radii = _RasterizeViews.apply(means3D, e(shs), e(colors_precomp), opacities, e(scales),
                                               e(rotations), e(cov3D_precomp), views, bool(return_mask), check,
                                               int(overlap))
    return (color, mask, radii) if return_mask else (color, radii)
