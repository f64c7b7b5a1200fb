"""Backward work units the forward emits (GhrStatus.reserved[1]) against the backward launch's grid upper bound
(R_cap / 128 + V * T) on the bench scene.  Finding (B200, 8 views): 8.4k units in a 24.1k-CTA grid; cutting the grid
to 65 % changed nothing measurable -- surplus CTAs exit on their first load and cost no visible time.
    python tools/unit_count.py"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import scenes, api, _native as NV
import util
dev = torch.device("cuda", 0)
sc = scenes.two_hand_scene(60000, seed=0)
cams = scenes.fibonacci_cameras(64, 512, 334, seed=0)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
for V in (1, 4, 8):
    views = util.gpu_views(cams[:V], np.zeros(3, np.float32), dev)
    r = api.forward_raw(views.cams(), t(sc.means3D), t(sc.opacities), t(sc.scales), t(sc.rotations), None, None, t(sc.colors), 0, 1.0)
    cap = int(r.R * 1.25) + (1 << 14)
    r = api.forward_raw(views.cams(), t(sc.means3D), t(sc.opacities), t(sc.scales), t(sc.rotations), None, None, t(sc.colors), 0, 1.0, R_cap=cap)
    torch.cuda.synchronize()
    st = r.state[:32].cpu().numpy().view(np.uint64)
    T = 32 * 21
    print("V", V, "R", int(st[0]), "units", int(st[3]), "grid upper bound", cap // 128 + V * T, "R/128", int(st[0]) // 128)
