import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from guassianhand_b200 import scenes
import util
worst = {}
def run(P,H,W,deg,mod,V,seed,huge):
    sc = scenes.random_scene(P, seed=seed, sh_degree=deg, behind_frac=0.1, huge_frac=huge)
    rng = np.random.default_rng(seed)
    bg = rng.uniform(0, 1, size=3).astype(np.float32)
    cams = [scenes.simple_camera(H, W, fx=float(rng.uniform(40, 300)))] if V==1 else scenes.fibonacci_cameras(V,H,W,seed=seed)
    rng2 = np.random.default_rng(seed)
    dL = (rng2.normal(size=(V, 3, H, W)) / (H * W)).astype(np.float32)
    gout, ggrad, info = util.run_gpu(sc, cams, bg, dL, scale_modifier=mod)
    sums = {}
    for v, cam in enumerate(cams):
        f, go = util.run_oracle(sc, cam, bg, dL[v], scale_modifier=mod)
        for k, a in go.items():
            sums[k] = sums.get(k, 0) + a.astype(np.float64)
    keys = {"dL_dmeans3D": "dL_dmeans3D", "dL_dcolors": "dL_dcolors", "dL_dopacity": "dL_dopacity", "dL_dcov3D": "dL_dcov3D", "dL_dsh": "dL_dsh", "dL_dscales": "dL_dscales", "dL_drots": "dL_drotations"}
    for ok, gk in keys.items():
        if gk in ggrad and ok in sums and sums[ok].size:
            got = ggrad[gk].reshape(sums[ok].shape)
            for fl in (1e-6, 1e-5, 1e-4):
                v_ = util.grad_violation(got, sums[ok], 1e-4, fl)
                worst[(ok, fl)] = max(worst.get((ok, fl), 0), v_)
            worst[(ok,'rel')] = max(worst.get((ok,'rel'),0), util.rel_err(got, sums[ok]))
run(2368,84,46,None,1.0,1,53,0.01)
print({k: round(v,3) if k[1]!='rel' else v for k,v in worst.items()})
rng = np.random.default_rng(0)
for i in range(25):
    run(int(rng.integers(1,3000)), int(rng.integers(1,150)), int(rng.integers(1,150)), [None,0,1,2,3][i%5], [1.0,0.6,1.9][i%3], 1+i%3, int(rng.integers(0,10000)), [0.0,0.01,0.05][i%3])
print({k: round(v,3) if k[1]!='rel' else v for k,v in worst.items()})
sc = scenes.two_hand_scene(60000, seed=0)
