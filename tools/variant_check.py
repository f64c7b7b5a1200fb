"""A/B helper: gradients of a build variant (guassianhand_b200/build.py VARIANTS) against the shipped library on the
same scenes (forward outputs must be identical when only the backward differs).
    python tools/variant_check.py halfq"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util  # noqa: E402
from guassianhand_b200 import scenes  # noqa: E402

name = sys.argv[1]
worst = 0.0
for P, H, W, nv, seed in ((4000, 96, 80, 2, 1), (60000, 512, 334, 3, 0), (20000, 129, 257, 2, 5), (300, 33, 47, 1, 7)):
    sc = scenes.two_hand_scene(P, seed=seed)
    cams = scenes.fibonacci_cameras(nv, H, W, seed=seed)
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    dL = (np.random.default_rng(seed).normal(size=(nv, 3, H, W)) / (H * W)).astype(np.float32)
    o0, g0, _ = util.run_gpu(sc, cams, bg, dL)
    with util.use_library_variant(name):
        o1, g1, _ = util.run_gpu(sc, cams, bg, dL)
    for v in range(nv):
        assert np.array_equal(o0[v]["out_color"], o1[v]["out_color"]) and np.array_equal(o0[v]["n_contrib"], o1[v]["n_contrib"])
    for k in g0:
        if k.startswith("_"):
            continue
        e = util.rel_err(g1[k], g0[k])
        gv = util.grad_violation(g1[k], g0[k], floor=5e-5)
        worst = max(worst, e)
        print(f"P={P} {H}x{W} {k:14s} rel_err {e:.2e}  violation {gv:.3f}")
print("worst rel_err", worst)
assert worst <= 1e-4
