"""ctypes front-end of the C oracle (oracle/gs_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) import this.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_fp = C.POINTER(C.c_float)


class _GsoIn(C.Structure):
    _fields_ = [
        ("P", C.c_int), ("H", C.c_int), ("W", C.c_int), ("M", C.c_int), ("D", C.c_int),
        ("tanfovx", C.c_float), ("tanfovy", C.c_float), ("scale_modifier", C.c_float),
        ("prefiltered", C.c_int),
        ("bg", _fp), ("view", _fp), ("proj", _fp), ("campos", _fp), ("means3D", _fp),
        ("opacities", _fp), ("scales", _fp), ("rotations", _fp), ("cov3D_precomp", _fp),
        ("shs", _fp), ("colors_precomp", _fp),
    ]


class _GsoFwd(C.Structure):
    _fields_ = [
        ("depths", _fp), ("radii", C.POINTER(C.c_int32)), ("xy", _fp), ("cov3D", _fp),
        ("conic_opacity", _fp), ("rgb", _fp), ("clamped", C.POINTER(C.c_uint8)),
        ("tiles_touched", C.POINTER(C.c_uint32)), ("offsets", C.POINTER(C.c_uint32)),
        ("R", C.c_uint64),
        ("keys_unsorted", C.POINTER(C.c_uint64)), ("vals_unsorted", C.POINTER(C.c_uint32)),
        ("keys", C.POINTER(C.c_uint64)), ("point_list", C.POINTER(C.c_uint32)),
        ("ranges", C.POINTER(C.c_uint32)), ("out_color", _fp), ("final_T", _fp),
        ("n_contrib", C.POINTER(C.c_uint32)), ("ambig", C.POINTER(C.c_uint8)),
        ("n_pairs", C.c_uint64),
    ]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libgs_oracle.so")
        if not os.path.exists(path):
            from oracle.build import build
            build()
        L = C.CDLL(path)
        L.gso_forward.restype = C.POINTER(_GsoFwd)
        L.gso_forward.argtypes = [C.POINTER(_GsoIn)]
        L.gso_free.argtypes = [C.POINTER(_GsoFwd)]
        L.gso_backward.restype = C.c_int
        L.gso_backward.argtypes = [C.POINTER(_GsoIn), C.POINTER(_GsoFwd)] + [_fp] * 10
        L.gso_higher_msb.restype = C.c_int
        L.gso_higher_msb.argtypes = [C.c_uint32]
        L.gso_mark_visible.argtypes = [C.c_int, _fp, _fp, _fp, C.POINTER(C.c_uint8)]
        L.gso_num_threads.restype = C.c_int
        L.gso_set_num_threads.argtypes = [C.c_int]
        _LIB = L
    return _LIB


def _f32(a):
    return None if a is None else np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(_fp)


class OracleScene:
    """Holds the (numpy, fp32) inputs of one rasterizer call: the arguments of
    GaussianRasterizer.forward plus the settings tuple (renderer_one_shot.py:281-346)."""

    def __init__(self, *, H, W, tanfovx, tanfovy, bg, viewmatrix, projmatrix, campos, means3D,
                 opacities, scales=None, rotations=None, cov3D_precomp=None, shs=None,
                 colors_precomp=None, sh_degree=0, scale_modifier=1.0, prefiltered=False):
        self.arr = dict(
            bg=_f32(bg).reshape(3), view=_f32(viewmatrix).reshape(16), proj=_f32(projmatrix).reshape(16),
            campos=_f32(campos).reshape(3), means3D=_f32(means3D).reshape(-1, 3),
            opacities=_f32(opacities).reshape(-1),
            scales=None if scales is None else _f32(scales).reshape(-1, 3),
            rotations=None if rotations is None else _f32(rotations).reshape(-1, 4),
            cov3D_precomp=None if cov3D_precomp is None else _f32(cov3D_precomp).reshape(-1, 6),
            shs=None if shs is None else _f32(shs),
            colors_precomp=None if colors_precomp is None else _f32(colors_precomp).reshape(-1, 3),
        )
        P = self.arr["means3D"].shape[0]
        M = 0 if shs is None else self.arr["shs"].shape[1]
        self.P, self.H, self.W, self.M, self.D = P, int(H), int(W), M, int(sh_degree)
        s = _GsoIn()
        s.P, s.H, s.W, s.M, s.D = P, int(H), int(W), M, int(sh_degree)
        s.tanfovx, s.tanfovy, s.scale_modifier = float(tanfovx), float(tanfovy), float(scale_modifier)
        s.prefiltered = int(bool(prefiltered))
        for k, v in self.arr.items():
            setattr(s, k, _ptr(v))
        self.c = s


def _np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape).copy()


def forward(scene: OracleScene, keep_handle: bool = False):
    L = lib()
    h = L.gso_forward(C.byref(scene.c))
    f = h.contents
    P, H, W = scene.P, scene.H, scene.W
    gx, gy = (W + 15) // 16, (H + 15) // 16
    R = int(f.R)
    out = dict(
        depths=_np(f.depths, (P,), np.float32), radii=_np(f.radii, (P,), np.int32),
        xy=_np(f.xy, (P, 2), np.float32), cov3D=_np(f.cov3D, (P, 6), np.float32),
        conic_opacity=_np(f.conic_opacity, (P, 4), np.float32), rgb=_np(f.rgb, (P, 3), np.float32),
        clamped=_np(f.clamped, (P, 3), np.uint8), tiles_touched=_np(f.tiles_touched, (P,), np.uint32),
        offsets=_np(f.offsets, (P,), np.uint32), R=R,
        keys_unsorted=_np(f.keys_unsorted, (R,), np.uint64),
        vals_unsorted=_np(f.vals_unsorted, (R,), np.uint32),
        keys=_np(f.keys, (R,), np.uint64), point_list=_np(f.point_list, (R,), np.uint32),
        ranges=_np(f.ranges, (gx * gy, 2), np.uint32),
        out_color=_np(f.out_color, (3, H, W), np.float32), final_T=_np(f.final_T, (H, W), np.float32),
        n_contrib=_np(f.n_contrib, (H, W), np.uint32), ambig=_np(f.ambig, (H, W), np.uint8),
        n_pairs=int(f.n_pairs),
    )
    if keep_handle:
        out["_handle"] = h
    else:
        L.gso_free(h)
    return out


def forward_backward(scene: OracleScene, dL_dout):
    """Returns (forward dict, grads dict).  dL_dout: [3,H,W] fp32."""
    L = lib()
    fwd = forward(scene, keep_handle=True)
    h = fwd.pop("_handle")
    P, M = scene.P, scene.M
    g = np.ascontiguousarray(np.asarray(dL_dout, dtype=np.float32)).reshape(3, scene.H, scene.W)
    Pa = max(P, 1)
    outs = dict(
        dL_dmeans3D=np.zeros((Pa, 3), np.float32), dL_dmeans2D=np.zeros((Pa, 3), np.float32),
        dL_dcolors=np.zeros((Pa, 3), np.float32), dL_dconic=np.zeros((Pa, 4), np.float32),
        dL_dopacity=np.zeros((Pa,), np.float32), dL_dcov3D=np.zeros((Pa, 6), np.float32),
        dL_dsh=np.zeros((Pa, max(M, 1), 3), np.float32), dL_dscales=np.zeros((Pa, 3), np.float32),
        dL_drots=np.zeros((Pa, 4), np.float32),
    )
    order = ["dL_dmeans3D", "dL_dmeans2D", "dL_dcolors", "dL_dconic", "dL_dopacity", "dL_dcov3D",
             "dL_dsh", "dL_dscales", "dL_drots"]
    rc = L.gso_backward(C.byref(scene.c), h, _ptr(g), *[_ptr(outs[k]) for k in order])
    L.gso_free(h)
    assert rc == 0
    outs = {k: v[:P] for k, v in outs.items()}
    if M == 0:
        outs["dL_dsh"] = np.zeros((P, 0, 3), np.float32)
    return fwd, outs


def mark_visible(means3D, viewmatrix, projmatrix):
    L = lib()
    m = _f32(means3D).reshape(-1, 3)
    v = _f32(viewmatrix).reshape(16)
    p = _f32(projmatrix).reshape(16)
    out = np.zeros((m.shape[0],), np.uint8)
    L.gso_mark_visible(m.shape[0], _ptr(m), _ptr(v), _ptr(p), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out.astype(bool)


def higher_msb(n: int) -> int:
    return int(lib().gso_higher_msb(int(n)))
