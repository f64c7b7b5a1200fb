"""The C-ABI library loads and exports every symbol include/ghr.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def native():
    from guassianhand_b200 import build, _native
    build.build()
    return _native


def _declared():
    src = open(os.path.join(ROOT, "include", "ghr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ghr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(native):
    L = native.lib()
    names = _declared()
    assert set(names) == set(native.EXPORTS)
    for n in names:
        assert hasattr(L, n), n


def test_abi_version_and_struct_sizes(native):
    L = native.lib()
    assert L.ghr_abi_version() == native.GHR_ABI_VERSION
    assert C.sizeof(native.GhrStatus) == 32
    assert C.sizeof(native.GhrDims) == 32


def test_layout_is_pure_host_logic(native):
    lay = native.layout(60000, 1, 512, 334, 0, 0, 1 << 20)
    assert lay.off_status == 0 and lay.off_geom >= 32
    assert lay.off_records - lay.off_geom >= 60000 * 64
    assert lay.state_bytes >= lay.off_ncontrib + 512 * 334 * 4
    assert lay.temp_bwd_bytes >= 60000 * 12 * 4
    # multi-view scales the per-view parts
    lay4 = native.layout(60000, 4, 512, 334, 0, 0, 1 << 20)
    assert lay4.off_ranges - lay4.off_geom >= 4 * 60000 * 64


def test_bad_dims_are_rejected_with_message(native):
    with pytest.raises(RuntimeError, match="bad dims"):
        native.layout(10, 0, 16, 16, 0, 0, 100)
    with pytest.raises(RuntimeError, match="2\\^30"):
        native.layout(10, 1, 16, 16, 0, 0, 1 << 31)


def test_null_args_do_not_crash(native):
    L = native.lib()
    assert L.ghr_forward(None, None) == native.GHR_EINVAL
    assert b"NULL" in L.ghr_last_error()
    assert L.ghr_backward(None, None) == native.GHR_EINVAL
    a = native.GhrForwardArgs()
    a.dims = native.GhrDims(4, 1, 16, 16, 0, 0, 64)
    assert L.ghr_forward(C.byref(a), None) == native.GHR_EINVAL      # required pointers missing
    assert b"required" in L.ghr_last_error()


def test_view_group_helpers_cpu():
    """Host logic of the overlapped schedule: contiguous balanced view groups, camera slicing."""
    import torch
    from guassianhand_b200 import api
    for V in (1, 2, 5, 8, 13):
        for G in (1, 2, 3, 8):
            G = min(G, V)
            gs = api._groups(V, G)
            assert gs[0][0] == 0 and gs[-1][1] == V and len(gs) == G
            assert all(a[1] == b[0] for a, b in zip(gs, gs[1:]))
            sizes = [hi - lo for lo, hi in gs]
            assert max(sizes) - min(sizes) <= 1 and min(sizes) >= 1
    V = 5
    cams = api._Cams(V=V, H=4, W=6, view=torch.arange(V * 16.).view(V, 16), proj=torch.zeros(V, 16),
                     campos=torch.zeros(V, 3), tanfov=torch.ones(V, 2), tanfovx=0.0, tanfovy=0.0,
                     bg=torch.arange(V * 3.).view(V, 3), bg_stride=3)
    c = api._slice_cams(cams, 2, 4)
    assert c.V == 2 and torch.equal(c.view, cams.view[2:4]) and torch.equal(c.bg, cams.bg[2:4])
    shared = cams._replace(bg=torch.zeros(3), bg_stride=0)
    assert api._slice_cams(shared, 1, 3).bg.shape == (3,)


def test_argument_checks_of_the_small_entry_points(native):
    """Entry points that validate before they launch: bad arguments come back as GHR_EINVAL with a message, without a GPU."""
    L = native.lib()
    assert L.ghr_sh_blend_forward(-1, None, None, None, None, None) == native.GHR_EINVAL
    assert L.ghr_sh_blend_forward(8, None, None, None, None, None) == native.GHR_EINVAL
    assert L.ghr_sh_blend_forward(0, None, None, None, None, None) == native.GHR_OK            # empty: nothing to do
    buf = (C.c_float * 8)()
    p = C.cast(buf, C.c_void_p)
    assert L.ghr_sh_blend_forward(8, p, None, p, p, None) == native.GHR_EINVAL                  # color_b without color_w
    assert b"color_w" in L.ghr_last_error()
    assert L.ghr_sh_blend_backward(8, p, None, None, p, None, p, None, None) == native.GHR_EINVAL   # d_w of an absent w
    assert L.ghr_cameras_from_w2c(1, None, None, 16, 16, 0.1, 100.0, None, None, None, None, None) == native.GHR_EINVAL
    assert L.ghr_cameras_from_w2c(0, None, None, 16, 16, 100.0, 0.1, None, None, None, None, None) == native.GHR_EINVAL
    assert L.ghr_mark_visible(-1, None, None, None, None, None) == native.GHR_EINVAL
    a = native.GhrAttributeArgs()
    a.P = 4
    assert L.ghr_attributes_forward(C.byref(a), None) == native.GHR_EINVAL
    assert b"required" in L.ghr_last_error()


def test_bench_only_standin_builds_and_exports():
    """baseline_standin/ (the upstream-structured GPU denominator of bench.py's gpu_baseline leg) compiles for
    sm_100a and exports its entry points; it is never imported by the product package."""
    from baseline_standin import build as sb
    so = sb.build()
    L = C.CDLL(so)
    for name in ("sgs_create", "sgs_destroy", "sgs_unpack", "sgs_forward", "sgs_backward", "sgs_num_rendered",
                 "sgs_n_contrib", "sgs_final_T", "sgs_ranges", "sgs_sorted_keys", "sgs_copy"):
        assert hasattr(L, name), name
    pkg = os.path.join(ROOT, "guassianhand_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "baseline_standin" not in src and "import oracle" not in src and "from oracle" not in src, fn


def test_geometry_cache_identity_and_version_rules_cpu():
    """api._GeomCache (the caller-invisible sharing between the reference's RGB and mask calls,
    renderer_one_shot.py:338-346 / :372-379): a hit needs the SAME tensors at the SAME autograd versions and the
    same scalars, and a state blob that is still alive.  Pure host logic: checked with CPU tensors."""
    import gc
    import torch
    from guassianhand_b200 import api
    C_ = api._GeomCache
    C_.entries.clear()
    mk = lambda *s: torch.zeros(*s)
    tensors = (mk(5, 3), mk(5, 1), mk(5, 3), mk(5, 4), None, mk(4, 4), mk(4, 4), mk(3))
    scalars = (5, 32, 32, 0.5, 0.5, 1.0, 0)
    state = torch.zeros(16, dtype=torch.uint8)
    assert C_.lookup(0, tensors, scalars) is None
    C_.store(0, tensors, scalars, state, 0, 1000, 123)
    hit = C_.lookup(0, tensors, scalars)
    assert hit is not None and hit[0] is state and hit[1:] == (0, 1000, 123)
    assert C_.lookup(1, tensors, scalars) is None                       # another device
    assert C_.lookup(0, tensors, scalars[:3] + (0.6,) + scalars[4:]) is None   # another tanfovx
    clone = tuple(None if t is None else t.clone() for t in tensors)
    assert C_.lookup(0, clone, scalars) is None                         # equal values, different tensors
    tensors[0].add_(1.0)                                                # in-place update: autograd version moves
    assert C_.lookup(0, tensors, scalars) is None
    C_.store(0, tensors, scalars, state, 0, 1000, 123)
    assert C_.lookup(0, tensors, scalars) is not None
    del state, hit
    gc.collect()
    assert C_.lookup(0, tensors, scalars) is None                       # the first call's state is gone
    C_.entries.clear()
