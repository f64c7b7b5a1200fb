"""ctypes binding of libghr.so -- mirrors include/ghr.h field by field.

The library is the product; there is no fallback.  If it is missing or does not export the ABI
this module raises, and every rasterizer call fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libghr.so")

GHR_OK, GHR_EINVAL, GHR_ENOSPC, GHR_ECUDA, GHR_EOVERFLOW = 0, -1, -2, -3, -4
GHR_FLAG_PREFILTERED, GHR_FLAG_DEBUG = 1, 2
GHR_STATUS_OVERFLOW, GHR_STATUS_PREFILTER = 1, 2
GHR_ABI_VERSION = 11
GHR_NSTAGES_FWD, GHR_NSTAGES_BWD = 5, 2
FWD_STAGES = ["preprocess", "tile_scan", "duplicate", "sort_gather", "blend_forward"]
BWD_STAGES = ["blend_backward", "preprocess_backward"]

EXPORTS = ["ghr_abi_version", "ghr_last_error", "ghr_struct_size", "ghr_layout", "ghr_forward", "ghr_backward",
           "ghr_mark_visible", "ghr_read_status_async", "ghr_event_create", "ghr_event_destroy", "ghr_event_record",
           "ghr_event_elapsed_ms", "ghr_fp32_probe", "ghr_attributes_forward", "ghr_attributes_backward",
           "ghr_sh_blend_forward", "ghr_sh_blend_backward",
           "ghr_cameras_from_w2c", "ghr_comm_create", "ghr_comm_handle", "ghr_comm_connect", "ghr_comm_buffer", "ghr_comm_allreduce",
           "ghr_comm_status", "ghr_comm_destroy"]
GHR_COMM_MAX_RANKS, GHR_COMM_HANDLE_BYTES = 8, 128
GHR_ATTR_XYZ_OFFSET, GHR_ATTR_RESTRICT_OFFSET, GHR_ATTR_CLIP_SCALING = 1, 2, 4

_vp = C.c_void_p


class GhrDims(C.Structure):
    _fields_ = [("P", C.c_int32), ("V", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("M", C.c_int32),
                ("sh_degree", C.c_int32), ("R_cap", C.c_int64)]


class GhrLayout(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in (
        "state_bytes", "temp_bytes", "temp_bwd_bytes", "off_status", "off_geom", "off_clamped", "off_ranges",
        "off_tilemax", "off_records", "off_final_T", "off_ncontrib", "off_order", "off_masks", "off_tilefinal",
        "off_ckpt", "off_units")]


class GhrStatus(C.Structure):
    _fields_ = [("R", C.c_uint64), ("overflow", C.c_uint32), ("n_visible", C.c_uint32),
                ("reserved", C.c_uint64 * 2)]


class GhrForwardArgs(C.Structure):
    _fields_ = [
        ("dims", GhrDims), ("flags", C.c_uint32), ("scale_modifier", C.c_float), ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("viewmatrix", _vp), ("projmatrix", _vp), ("campos", _vp), ("tanfov", _vp), ("bg", _vp),
        ("bg_stride", C.c_int32),
        ("means3D", _vp), ("opacities", _vp), ("scales", _vp), ("rotations", _vp), ("cov3D_precomp", _vp),
        ("shs", _vp), ("colors_precomp", _vp),
        ("out_color", _vp), ("radii", _vp), ("out_mask", _vp),
        ("state", _vp), ("state_bytes", C.c_size_t), ("temp", _vp), ("temp_bytes", C.c_size_t),
        ("dbg_keys_sorted", _vp), ("dbg_point_list", _vp),
        ("host_status", _vp), ("seq", C.c_uint64), ("stage_events", _vp),
        ("reuse_state", _vp), ("reuse_M", C.c_int32),
    ]


class GhrBackwardArgs(C.Structure):
    _fields_ = [
        ("dims", GhrDims), ("flags", C.c_uint32), ("scale_modifier", C.c_float), ("tanfovx", C.c_float),
        ("tanfovy", C.c_float),
        ("viewmatrix", _vp), ("projmatrix", _vp), ("campos", _vp), ("tanfov", _vp), ("bg", _vp),
        ("bg_stride", C.c_int32),
        ("means3D", _vp), ("opacities", _vp), ("scales", _vp), ("rotations", _vp), ("cov3D_precomp", _vp),
        ("shs", _vp), ("colors_precomp", _vp),
        ("dL_dout_color", _vp), ("dL_dout_mask", _vp),
        ("state", _vp), ("state_bytes", C.c_size_t), ("temp", _vp), ("temp_bytes", C.c_size_t),
        ("accumulate", C.c_int32),
        ("dL_dmeans3D", _vp), ("dL_dmeans2D", _vp), ("dL_dcolors", _vp), ("dL_dopacity", _vp),
        ("dL_dcov3D", _vp), ("dL_dsh", _vp), ("dL_dscales", _vp), ("dL_drotations", _vp), ("dL_dconic", _vp),
        ("stage_events", _vp),
    ]


class GhrAttributeArgs(C.Structure):
    _fields_ = [("P", C.c_int32), ("flags", C.c_uint32), ("clip_scaling", C.c_float)] + [(n, _vp) for n in (
        "xyz_raw", "pts", "scaling_raw", "rotation_raw", "opacity_raw", "rgb_raw", "xyz_b", "opacity_b", "color_w0",
        "color_w1", "color_b0", "means3D", "scales", "rotations", "opacities", "colors")]


class GhrAttributeGrads(C.Structure):
    _fields_ = [(n, _vp) for n in (
        "dL_dmeans3D", "dL_dscales", "dL_drotations", "dL_dopacity", "dL_dcolors", "d_xyz_raw", "d_pts",
        "d_scaling_raw", "d_rotation_raw", "d_opacity_raw", "d_rgb_raw", "d_xyz_b", "d_opacity_b", "d_color_w0",
        "d_color_w1", "d_color_b0")]


_lib = None


def lib():
    """Load libghr.so (built in-tree by guassianhand_b200.build).  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA library is required (there is no CPU fallback). "
            "Build it with `python -m guassianhand_b200.build`.")
    L = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(L, name):
            raise RuntimeError(f"libghr.so does not export {name}")
    L.ghr_abi_version.restype = C.c_int
    L.ghr_last_error.restype = C.c_char_p
    L.ghr_layout.restype = C.c_int
    L.ghr_layout.argtypes = [C.POINTER(GhrDims), C.POINTER(GhrLayout)]
    L.ghr_forward.restype = C.c_int
    L.ghr_forward.argtypes = [C.POINTER(GhrForwardArgs), _vp]
    L.ghr_backward.restype = C.c_int
    L.ghr_backward.argtypes = [C.POINTER(GhrBackwardArgs), _vp]
    L.ghr_mark_visible.restype = C.c_int
    L.ghr_mark_visible.argtypes = [C.c_int32, _vp, _vp, _vp, _vp, _vp]
    L.ghr_read_status_async.restype = C.c_int
    L.ghr_read_status_async.argtypes = [_vp, _vp, _vp]
    L.ghr_event_create.argtypes = [C.POINTER(_vp)]
    L.ghr_event_destroy.argtypes = [_vp]
    L.ghr_event_record.argtypes = [_vp, _vp]
    L.ghr_event_elapsed_ms.argtypes = [_vp, _vp, C.POINTER(C.c_float)]
    L.ghr_fp32_probe.argtypes = [C.c_int32, C.c_int32, _vp, _vp, C.POINTER(C.c_double), _vp]
    L.ghr_struct_size.restype = C.c_size_t
    L.ghr_struct_size.argtypes = [C.c_char_p]
    L.ghr_attributes_forward.restype = C.c_int
    L.ghr_attributes_forward.argtypes = [C.POINTER(GhrAttributeArgs), _vp]
    L.ghr_attributes_backward.restype = C.c_int
    L.ghr_attributes_backward.argtypes = [C.POINTER(GhrAttributeArgs), C.POINTER(GhrAttributeGrads), _vp]
    L.ghr_sh_blend_forward.restype = C.c_int
    L.ghr_sh_blend_forward.argtypes = [C.c_int64, _vp, _vp, _vp, _vp, _vp]
    L.ghr_sh_blend_backward.restype = C.c_int
    L.ghr_sh_blend_backward.argtypes = [C.c_int64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]
    L.ghr_cameras_from_w2c.restype = C.c_int
    L.ghr_cameras_from_w2c.argtypes = [C.c_int32, _vp, _vp, C.c_int32, C.c_int32, C.c_float, C.c_float, _vp, _vp, _vp,
                                       _vp, _vp]
    L.ghr_comm_create.restype = C.c_int
    L.ghr_comm_create.argtypes = [C.c_int32, C.c_int32, C.c_size_t, C.POINTER(_vp)]
    L.ghr_comm_handle.restype = C.c_int
    L.ghr_comm_handle.argtypes = [_vp, _vp]
    L.ghr_comm_connect.restype = C.c_int
    L.ghr_comm_connect.argtypes = [_vp, _vp]
    L.ghr_comm_buffer.restype = _vp
    L.ghr_comm_buffer.argtypes = [_vp]
    L.ghr_comm_allreduce.restype = C.c_int
    L.ghr_comm_allreduce.argtypes = [_vp, C.c_size_t, _vp]
    L.ghr_comm_status.restype = C.c_int
    L.ghr_comm_status.argtypes = [_vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.ghr_comm_destroy.restype = C.c_int
    L.ghr_comm_destroy.argtypes = [_vp]
    for cls in (GhrDims, GhrLayout, GhrStatus, GhrForwardArgs, GhrBackwardArgs, GhrAttributeArgs, GhrAttributeGrads):
        want = L.ghr_struct_size(cls.__name__.encode())
        if want != C.sizeof(cls):
            raise RuntimeError(f"ctypes mirror of {cls.__name__} is {C.sizeof(cls)} bytes, libghr.so says {want}")
    if L.ghr_abi_version() != GHR_ABI_VERSION:
        raise RuntimeError("libghr.so ABI version mismatch; rebuild with `python -m guassianhand_b200.build --force`")
    _lib = L
    return L


def check(rc: int, what: str = "libghr"):
    if rc != GHR_OK:
        msg = lib().ghr_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def layout(P, V, H, W, M, sh_degree, R_cap) -> GhrLayout:
    d = GhrDims(P, V, H, W, M, sh_degree, R_cap)
    out = GhrLayout()
    check(lib().ghr_layout(C.byref(d), C.byref(out)), "ghr_layout")
    return out


class StageEvents:
    """2*n cudaEvent_t handles (start, stop per stage) for GhrForwardArgs/GhrBackwardArgs.stage_events."""

    def __init__(self, n: int):
        self.n = n
        self.arr = (_vp * (2 * n))()
        for i in range(2 * n):
            e = _vp()
            check(lib().ghr_event_create(C.byref(e)), "ghr_event_create")
            self.arr[i] = e.value

    def ptr(self):
        return C.cast(self.arr, _vp).value

    def elapsed_ms(self, i: int) -> float:
        ms = C.c_float()
        check(lib().ghr_event_elapsed_ms(self.arr[2 * i], self.arr[2 * i + 1], C.byref(ms)), "ghr_event_elapsed_ms")
        return float(ms.value)

    def close(self):
        for i in range(2 * self.n):
            if self.arr[i]:
                lib().ghr_event_destroy(self.arr[i])
                self.arr[i] = None
