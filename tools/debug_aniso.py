import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import scenes
import util
from test_gpu_parity import _aniso_scene, BG
cam = scenes.simple_camera(96, 128, fx=150.0)
for seed in (41, 42, 43):
    sc = _aniso_scene(seed, [1.0, 3.0, 8.0], [1.0, 2.0, 4.0])
    H, W = cam.H, cam.W
    dL = (np.random.default_rng(seed).normal(size=(1, 3, H, W)) / (H * W)).astype(np.float32)
    gout, gg, info = util.run_gpu(sc, [cam], BG, dL)
    f, go = util.run_oracle(sc, cam, BG, dL[0])
    print("seed", seed, "R", info["R"])
    for ok, gk in [("dL_dmeans2D", "dL_dmeans2D"), ("dL_dconic", "dL_dconic"), ("dL_dopacity", "dL_dopacity"), ("dL_dcolors", "dL_dcolors"),
                   ("dL_dcov3D", "dL_dcov3D"), ("dL_dmeans3D", "dL_dmeans3D"), ("dL_dscales", "dL_dscales"), ("dL_drots", "dL_drotations")]:
        a = gg[gk].reshape(go[ok].shape).astype(np.float64); b = go[ok].astype(np.float64)
        d = np.abs(a - b)
        i = np.unravel_index(d.argmax(), d.shape)
        print(f"  {ok:12s} rel={d.max()/np.abs(b).max():.3e} worst idx={i} gpu={a[i]:.6e} ref={b[i]:.6e} maxref={np.abs(b).max():.3e}")
        if ok == "dL_dmeans3D":
            g = i[0]
            print("     gaussian", g, "scale", sc.scales[g], "opac", sc.opacities[g], "radius", f["radii"][g], "conic", f["conic_opacity"][g], "xy", f["xy"][g], "depth", f["depths"][g])
            print("     m2d gpu/ref", gg["dL_dmeans2D"][0][g], go["dL_dmeans2D"][g], "conic gpu/ref", gg["dL_dconic"][0][g], go["dL_dconic"][g])
            print("     cov3D grad gpu/ref", gg["dL_dcov3D"][g], go["dL_dcov3D"][g])
