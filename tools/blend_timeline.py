"""Per-CTA timeline of the forward blend (profiling variant libghr_timeline.so, built with
`python -m guassianhand_b200.build timeline`): when every non-empty tile's CTA started and stopped, on which SM.
usage: python tools/blend_timeline.py [--views 8]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from guassianhand_b200 import _native as NV, build  # noqa: E402
NV.LIB_PATH = build.variant_path("timeline")
from guassianhand_b200 import scenes  # noqa: E402
from guassianhand_b200.dist import PackedGrads, fit_step_grads  # noqa: E402
import util  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--views", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
L = NV.lib()
sc = scenes.two_hand_scene(60000, seed=0)
cams = scenes.fibonacci_cameras(64, 512, 334, seed=0)
t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).float().to(dev)
gauss = dict(means3D=t(sc.means3D), opacities=t(sc.opacities), scales=t(sc.scales), rotations=t(sc.rotations),
             colors_precomp=t(sc.colors))
V = a.views
views = util.gpu_views(cams[:V], np.zeros(3, np.float32), dev)
dL = t((np.random.default_rng(1).normal(size=(V, 3, 512, 334)) / (512 * 334)).astype(np.float32))
grads = PackedGrads(60000, 0, device=dev)
res = fit_step_grads(gauss, views, dL, grads)
cap = int(res.R * 1.25) + (1 << 14)
buf = np.zeros((1 << 17, 4), np.uint64)
cnt = C.c_uint()
for it in range(3):
    fit_step_grads(gauss, views, dL, grads, R_cap=cap, check="none")
    torch.cuda.synchronize()
    L.ghr_debug_timeline(buf.ctypes.data_as(C.c_void_p), C.c_uint(buf.shape[0]), C.byref(cnt))
n = cnt.value
tl = buf[:n]
kind = (tl[:, 2] >> np.uint64(32)).astype(int)
fw = tl[kind == 0]
t0 = fw[:, 0].min()
start = (fw[:, 0] - t0).astype(np.int64) / 1e3
stop = (fw[:, 1] - t0).astype(np.int64) / 1e3
smid = (fw[:, 2] & np.uint64(0xFFFFFFFF)).astype(int)
nlist = (fw[:, 3] & np.uint64(0xFFFFFFFF)).astype(int)
dur = stop - start
print(f"forward CTAs (non-empty tiles): {len(fw)}  span {stop.max():.1f} us  sum of durations {dur.sum():.0f} us "
      f"(= {dur.sum() / stop.max():.1f} CTAs in flight on average, {dur.sum() / stop.max() / 148:.2f} per SM)")
vts = (fw[:, 3] >> np.uint64(32)).astype(int)
# per-warp counters (kind 1 + warp): {survivors << 12 | iterations, passes << 12 | live stages}, keyed by start time + SM
ck = tl[kind >= 16]
ckey = {}
for row in ck:
    k = (int(row[0]), int(row[2] & np.uint64(0xFFFFFFFF)))
    a_, b_ = int(row[3] >> np.uint64(32)), int(row[3] & np.uint64(0xFFFFFFFF))
    ckey.setdefault(k, []).append((int(row[2] >> np.uint64(32)) - 16, a_ * 16, (b_ >> 16) * 16, (b_ & 0xFFFF) * 16))
wk = tl[(kind >= 1) & (kind < 16)]
wkey = {}
for row in wk:
    k = (int(row[0]), int(row[2] & np.uint64(0xFFFFFFFF)))
    a_, b_ = int(row[3] >> np.uint64(32)), int(row[3] & np.uint64(0xFFFFFFFF))
    wkey.setdefault(k, []).append((int(row[2] >> np.uint64(32)) - 1, a_ >> 12, a_ & 0xFFF, b_ >> 12, b_ & 0xFFF,
                                   (int(row[1]) - int(row[0])) / 1e3))
order = np.argsort(-dur)[:12]
print("longest CTAs: list length, start, stop, duration (us), SM; per warp: survivors/iterations/passes/live stages/us")
for i in order:
    print(f"  n={nlist[i]:5d} start {start[i]:6.1f} stop {stop[i]:6.1f} dur {dur[i]:6.1f} sm {smid[i]}")
    ws = sorted(wkey.get((int(fw[i, 0]), int(smid[i])), []))
    print("     " + "  ".join(f"w{w}:{sv}/{it}/{ps}/{rd}/{us:.0f}" for w, sv, it, ps, rd, us in ws))
    print("     cycles wait/compaction/blend: " + "  ".join(f"w{w}:{a_}/{b_}/{c_}" for w, a_, b_, c_ in sorted(ckey.get((int(fw[i, 0]), int(smid[i])), []))))
allw = [x for v_ in wkey.values() for x in v_]
if allw:
    sv = np.array([x[1] for x in allw]); it = np.array([x[2] for x in allw]); ps = np.array([x[3] for x in allw])
    print(f"all warps: survivors {sv.sum()}  iterations {it.sum()} (x4 = {4 * it.sum()} slots, {sv.sum() / max(4 * it.sum(), 1):.2f} filled)  "
          f"passes {ps.sum()}  survivors/pass {sv.sum() / max(ps.sum(), 1):.1f}")
edges = np.arange(0, stop.max() + 10, 10.0)
print("CTAs running at t (us):", [(int(e), int(((start <= e) & (stop > e)).sum())) for e in edges])
print("CTA starts per 10 us:", np.histogram(start, bins=edges)[0].tolist())
