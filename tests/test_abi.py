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
