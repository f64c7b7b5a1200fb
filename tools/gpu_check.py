"""Stage-by-stage parity report of the CUDA path against the CPU oracle (diagnostic; the asserting
version of the same checks lives in tests/test_gpu_parity.py).  Run on the GPU box:
    python tools/gpu_check.py [--quick]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from guassianhand_b200 import scenes  # noqa: E402
import util  # noqa: E402


def compare(name, scene, cams, bg, seed=0, scale_modifier=1.0, cov3D=None):
    H, W = cams[0].H, cams[0].W
    V = len(cams)
    rng = np.random.default_rng(seed + 7)
    dL = (rng.normal(size=(V, 3, H, W)) / (H * W)).astype(np.float32)
    t0 = time.time()
    gout, ggrad, info = util.run_gpu(scene, cams, bg, dL, scale_modifier=scale_modifier, cov3D=cov3D)
    t1 = time.time()
    rep = dict(name=name, P=scene.P, V=V, H=H, W=W, R=info["R"], R_cap=info["R_cap"], overflow=info["overflow"],
               gpu_s=round(t1 - t0, 3), views=[])
    sums = {}
    for v, cam in enumerate(cams):
        fo, go = util.run_oracle(scene, cam, bg, dL[v], scale_modifier=scale_modifier, cov3D=cov3D)
        g = gout[v]
        r = dict(
            R_oracle=fo["R"], R_gpu=g["R"],
            radii_mismatch=int((fo["radii"] != g["radii"]).sum()),
            tiles_mismatch=int((fo["tiles_touched"] != g["tiles_touched"]).sum()),
            depth_bits_mismatch=int((fo["depths"].view(np.uint32) != g["depths"].view(np.uint32)).sum()),
            xy_bits_mismatch=int((fo["xy"].view(np.uint32) != g["xy"].view(np.uint32)).sum()),
            conic_bits_mismatch=int((fo["conic_opacity"].view(np.uint32) != g["conic_opacity"].view(np.uint32)).sum()),
            rgb_bits_mismatch=int((fo["rgb"].view(np.uint32) != g["rgb"].view(np.uint32)).sum()),
            keys_equal=bool(fo["keys"].shape == g["keys"].shape and (fo["keys"] == g["keys"]).all()),
            plist_equal=bool(fo["point_list"].shape == g["point_list"].shape and (fo["point_list"] == g["point_list"]).all()),
            ranges_equal=bool((fo["ranges"] == g["ranges"]).all()),
            ncontrib_mismatch=int((fo["n_contrib"] != g["n_contrib"]).sum()),
            ncontrib_mismatch_unambiguous=int(((fo["n_contrib"] != g["n_contrib"]) & (fo["ambig"] == 0)).sum()),
            ambig=int(fo["ambig"].sum()),
            image_maxabs=float(np.abs(fo["out_color"] - g["out_color"]).max()),
            finalT_maxabs=float(np.abs(fo["final_T"] - g["final_T"]).max()),
            pairs=fo["n_pairs"],
        )
        r["dL_dmeans2D"] = util.rel_err(ggrad["dL_dmeans2D"][v], go["dL_dmeans2D"])
        r["dL_dconic"] = util.rel_err(ggrad["dL_dconic"][v], go["dL_dconic"])
        for k in go:
            if k in ("dL_dmeans2D", "dL_dconic"):
                continue
            sums[k] = sums.get(k, 0) + go[k].astype(np.float64)
        rep["views"].append(r)
    gk = {"dL_dmeans3D": "dL_dmeans3D", "dL_dcolors": "dL_dcolors", "dL_dopacity": "dL_dopacity",
          "dL_dcov3D": "dL_dcov3D", "dL_dsh": "dL_dsh", "dL_dscales": "dL_dscales", "dL_drots": "dL_drotations"}
    rep["grads"] = {}
    for ok, gkk in gk.items():
        if gkk in ggrad and ok in sums and sums[ok].size:
            rep["grads"][ok] = util.rel_err(ggrad[gkk].reshape(sums[ok].shape), sums[ok])
    print(json.dumps(rep))
    sys.stdout.flush()
    return rep


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "gpu_check.json"))
    a = ap.parse_args()
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    reps = []
    reps.append(compare("rand2k_64x80", scenes.random_scene(2000, seed=0), [scenes.simple_camera(64, 80)], bg))
    reps.append(compare("rand2k_sh3_70x45", scenes.random_scene(2000, seed=3, sh_degree=3),
                        [scenes.simple_camera(70, 45)], bg))
    reps.append(compare("hands6k_3views_128x96", scenes.two_hand_scene(6000, seed=1),
                        scenes.fibonacci_cameras(3, 128, 96, seed=1), bg))
    reps.append(compare("hands6k_sh2_2views", scenes.two_hand_scene(6000, seed=2, sh_degree=2),
                        scenes.fibonacci_cameras(2, 100, 120, seed=2), bg))
    if not a.quick:
        reps.append(compare("C2_60k_512x334", scenes.two_hand_scene(60000, seed=0),
                            scenes.fibonacci_cameras(1, 512, 334, seed=0), np.zeros(3, np.float32)))
        reps.append(compare("C2_60k_4views", scenes.two_hand_scene(60000, seed=0),
                            scenes.fibonacci_cameras(4, 512, 334, seed=0), np.zeros(3, np.float32)))
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(reps, f, indent=1)


if __name__ == "__main__":
    main()
