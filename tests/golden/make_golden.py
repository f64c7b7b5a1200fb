"""Generates the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference has no golden vectors for this path, so these are produced by
oracle/gs_oracle.c (itself pinned by tests/test_oracle_kat.py and tests/test_oracle_cross.py).
They freeze the oracle's output so that (a) drift of the oracle on another host/libm is caught and
(b) the GPU box can check the CUDA path without re-deriving anything.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from guassianhand_b200 import scenes  # noqa: E402
import util  # noqa: E402

CASES = {
    # name: (scene factory, camera factory, bg)
    "rand_rgb_40x56": (lambda: scenes.random_scene(400, seed=11), lambda: scenes.simple_camera(40, 56),
                       [0.1, 0.2, 0.3]),
    "rand_sh3_33x47": (lambda: scenes.random_scene(300, seed=12, sh_degree=3), lambda: scenes.simple_camera(33, 47),
                       [0.0, 0.0, 0.0]),
    "hands_rgb_64x48": (lambda: scenes.two_hand_scene(1500, seed=13),
                        lambda: scenes.fibonacci_cameras(3, 64, 48, seed=13)[1], [1.0, 1.0, 1.0]),
}


def case_inputs(name):
    mk_s, mk_c, bg = CASES[name]
    sc, cam = mk_s(), mk_c()
    rng = np.random.default_rng(abs(hash(name)) % 1000 if False else len(name))
    dL = (rng.normal(size=(3, cam.H, cam.W)) / (cam.H * cam.W)).astype(np.float32)
    return sc, cam, np.asarray(bg, np.float32), dL


def main():
    for name in CASES:
        sc, cam, bg, dL = case_inputs(name)
        f, g = util.run_oracle(sc, cam, bg, dL)
        keep = {k: f[k] for k in ("radii", "tiles_touched", "depths", "xy", "conic_opacity", "rgb", "keys",
                                  "point_list", "ranges", "n_contrib", "ambig", "out_color", "final_T")}
        keep.update({"g_" + k: v for k, v in g.items()})
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **keep)
        print(name, "R", f["R"], "pairs", f["n_pairs"])


if __name__ == "__main__":
    main()
