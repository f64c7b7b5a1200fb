"""The oracle reproduces the committed golden fixtures (tests/golden/*.npz, made by
tests/golden/make_golden.py): integer/byte intermediates bit-exact, floats to 1e-6."""
import os

import numpy as np
import pytest

import util
from golden.make_golden import CASES, case_inputs

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    sc, cam, bg, dL = case_inputs(name)
    f, g = util.run_oracle(sc, cam, bg, dL)
    for k in ("radii", "tiles_touched", "keys", "point_list", "ranges"):
        assert np.array_equal(f[k], z[k]), k
    for k in ("depths", "xy", "conic_opacity", "rgb"):
        assert np.array_equal(f[k].view(np.uint32), z[k].view(np.uint32)), k      # no exp involved: bit-exact
    amb = (f["ambig"] != 0) | (z["ambig"] != 0)
    assert np.array_equal(f["n_contrib"][~amb], z["n_contrib"][~amb])
    assert np.abs(f["out_color"] - z["out_color"]).max() < 1e-6
    for k, v in g.items():
        if v.size:
            assert util.rel_err(v, z["g_" + k]) < 1e-5, k
