"""Seeded synthetic two-hand Gaussian scenes and cameras.

Host-side only (numpy); identical arrays feed the oracle, the CUDA path and the benchmark.
Camera construction restates GuassianHand's own math so that the matrices handed to the
rasterizer have exactly the layout the reference produces:
  * projection from intrinsics  -- /root/reference/tgs/models/renderer_one_shot.py:61-81
  * world_view_transform = w2c^T, full_proj = view @ proj^T-form, camera_center -- :90-107
  * tanfov = tan(0.5 * 2*atan2(W, 2 fx))  -- :83-87, :278-279
No MANO data is used (licence); hands are ellipsoid palms + capsule fingers.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np


@dataclass
class CameraParams:
    """One view, in the exact form GaussianRasterizationSettings carries it."""
    H: int
    W: int
    tanfovx: float
    tanfovy: float
    viewmatrix: np.ndarray   # [4,4] fp32, = w2c^T (row-major memory reads column-major in CUDA)
    projmatrix: np.ndarray   # [4,4] fp32, = (P @ w2c)^T
    campos: np.ndarray       # [3] fp32


@dataclass
class GaussianScene:
    means3D: np.ndarray      # [P,3]
    scales: np.ndarray       # [P,3]
    rotations: np.ndarray    # [P,4] normalised (r,x,y,z)
    opacities: np.ndarray    # [P,1]
    colors: Optional[np.ndarray]  # [P,3] precomputed colours
    shs: Optional[np.ndarray]     # [P,M,3]
    sh_degree: int = 0

    @property
    def P(self) -> int:
        return int(self.means3D.shape[0])


def projection_from_intrinsics(K: np.ndarray, H: int, W: int, znear: float, zfar: float) -> np.ndarray:
    """renderer_one_shot.py:61-81 (getProjectionMatrix_refine)."""
    fx, fy, cx, cy, s = K[0, 0], K[1, 1], K[0, 2], K[1, 2], K[0, 1]
    P = np.zeros((4, 4), dtype=np.float32)
    P[0, 0] = 2 * fx / W
    P[0, 1] = 2 * s / W
    P[0, 2] = -1 + 2 * (cx / W)
    P[1, 1] = 2 * fy / H
    P[1, 2] = -1 + 2 * (cy / H)
    P[2, 2] = (zfar + znear) / (zfar - znear)
    P[2, 3] = -1 * 2 * zfar * znear / (zfar - znear)
    P[3, 2] = 1.0
    return P


def camera_from_w2c(w2c: np.ndarray, K: np.ndarray, H: int, W: int) -> CameraParams:
    """renderer_one_shot.py:90-112 (Camera.from_w2c) + :278-279 (tanfov)."""
    w2c = w2c.astype(np.float32)
    K = K.astype(np.float32)
    view = np.ascontiguousarray(w2c.T)
    proj = projection_from_intrinsics(K, H, W, 0.01, 1000.0).T
    full = (view @ proj).astype(np.float32)
    campos = np.linalg.inv(view.astype(np.float64))[3, :3].astype(np.float32)
    fovx = 2 * np.arctan2(np.float32(W), 2 * K[0, 0])
    fovy = 2 * np.arctan2(np.float32(H), 2 * K[1, 1])
    return CameraParams(H=H, W=W, tanfovx=float(np.tan(fovx * 0.5)), tanfovy=float(np.tan(fovy * 0.5)),
                        viewmatrix=view, projmatrix=np.ascontiguousarray(full), campos=campos)


def look_at_w2c(eye: np.ndarray, target: np.ndarray, up=(0.0, -1.0, 0.0)) -> np.ndarray:
    """OpenCV-style camera (x right, y down, z forward) looking from `eye` to `target`."""
    f = target - eye
    f = f / np.linalg.norm(f)
    upv = np.asarray(up, dtype=np.float64)
    r = np.cross(f, -upv)
    if np.linalg.norm(r) < 1e-6:
        r = np.cross(f, np.array([1.0, 0, 0]))
    r = r / np.linalg.norm(r)
    d = np.cross(f, r)
    Rm = np.stack([r, d, f], axis=0)
    w2c = np.eye(4)
    w2c[:3, :3] = Rm
    w2c[:3, 3] = -Rm @ eye
    return w2c


def fibonacci_cameras(V: int, H: int, W: int, seed: int = 0, dist: float = 1.0) -> List[CameraParams]:
    """V views on a Fibonacci sphere, radius U(0.9,1.1)*dist, fx=fy=1300*W/334, off-centre principal
    point (cx=W/2+3.7, cy=H/2-2.1) -- SURVEY.md §8(d)."""
    rng = np.random.default_rng(seed + 1000)
    cams = []
    fx = 1300.0 * W / 334.0
    K = np.array([[fx, 0, W / 2 + 3.7], [0, fx, H / 2 - 2.1], [0, 0, 1]], dtype=np.float64)
    ga = np.pi * (3.0 - np.sqrt(5.0))
    for v in range(V):
        if V == 1:
            dirv = np.array([0.0, 0.0, -1.0])
        else:
            # front-biased band so the palm plane is never seen edge-on only
            zc = 1.0 - 2.0 * (v + 0.5) / V
            rad = np.sqrt(max(0.0, 1.0 - zc * zc))
            th = ga * v
            dirv = np.array([np.cos(th) * rad, 0.6 * zc, -abs(np.sin(th) * rad) - 0.35])
            dirv /= np.linalg.norm(dirv)
        r = dist * rng.uniform(0.9, 1.1)
        eye = dirv * r
        w2c = look_at_w2c(eye, np.zeros(3))
        cams.append(camera_from_w2c(w2c, K, H, W))
    return cams


def _sample_ellipsoid(rng, n, semi):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v * np.asarray(semi)[None, :]


def _sample_capsule(rng, n, radius, length):
    """capsule along +y from 0 to length"""
    frac_cap = (2 * radius) / (2 * radius + length)
    on_cap = rng.uniform(size=n) < frac_cap
    th = rng.uniform(0, 2 * np.pi, size=n)
    y = rng.uniform(0, length, size=n)
    pts = np.stack([radius * np.cos(th), y, radius * np.sin(th)], axis=1)
    s = rng.normal(size=(n, 3))
    s /= np.linalg.norm(s, axis=1, keepdims=True)
    s *= radius
    s[:, 1] = np.where(s[:, 1] > 0, s[:, 1] + length, s[:, 1])
    pts[on_cap] = s[on_cap]
    return pts


def _one_hand(rng, n, mirror: bool):
    n_palm = int(round(0.40 * n))
    n_f = [(n - n_palm) // 5] * 5
    n_f[-1] = n - n_palm - sum(n_f[:-1])
    parts = [_sample_ellipsoid(rng, n_palm, (0.045, 0.050, 0.012))]
    angles = np.deg2rad(np.linspace(-25, 25, 5))
    for k in range(5):
        radius = rng.uniform(0.008, 0.010)
        length = rng.uniform(0.060, 0.090)
        pts = _sample_capsule(rng, n_f[k], radius, length)
        ca, sa = np.cos(angles[k]), np.sin(angles[k])
        rot = np.array([[ca, -sa, 0], [sa, ca, 0], [0, 0, 1]])
        pts = pts @ rot.T
        pts += np.array([0.036 * np.sin(angles[k]) / np.sin(np.deg2rad(25)), 0.045, 0.0])
        parts.append(pts)
    pts = np.concatenate(parts, axis=0)
    # right hand at x=-45 mm, fingers pointing toward +x (the interaction zone)
    rz = np.deg2rad(-70.0)
    rot = np.array([[np.cos(rz), -np.sin(rz), 0], [np.sin(rz), np.cos(rz), 0], [0, 0, 1]])
    pts = pts @ rot.T
    pts[:, 0] -= 0.045
    if mirror:
        pts[:, 0] = -pts[:, 0]
        pts[:, 1] += 0.012   # interleave fingers
        pts[:, 2] += 0.006
    return pts


def two_hand_scene(P: int = 60000, seed: int = 0, sh_degree: Optional[int] = None, hands: int = 2,
                   tile: int = 1) -> GaussianScene:
    """SURVEY.md §8(d) "two-hand" geometry.  `sh_degree=None` -> precomputed colours (the path the
    reference uses, config_one_shot.yaml:188); otherwise SH coefficients of that degree.
    `tile>1` replicates the pair on a tile x tile grid with jitter (config C4)."""
    rng = np.random.default_rng(seed)
    reps = tile * tile
    per = P // (hands * reps)
    chunks = []
    for rep in range(reps):
        off = np.zeros(3)
        if tile > 1:
            off = np.array([(rep % tile - (tile - 1) / 2) * 0.30, (rep // tile - (tile - 1) / 2) * 0.24, 0.0])
            off += rng.normal(scale=0.01, size=3)
        for h in range(hands):
            n = per if not (rep == reps - 1 and h == hands - 1) else P - per * (hands * reps - 1)
            chunks.append(_one_hand(rng, n, mirror=(h == 1)) + off[None, :])
    means = np.concatenate(chunks, axis=0).astype(np.float32)
    Pn = means.shape[0]
    scales = np.exp(rng.normal(np.log(0.0015), 0.35, size=(Pn, 3)))
    scales = np.clip(scales, 0.0003, 0.006).astype(np.float32)
    q = rng.normal(size=(Pn, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = (1.0 / (1.0 + np.exp(-rng.normal(1.0, 1.5, size=(Pn, 1))))).astype(np.float32)
    colors = None
    shs = None
    deg = 0
    if sh_degree is None:
        colors = rng.uniform(0, 1, size=(Pn, 3)).astype(np.float32)
    else:
        deg = int(sh_degree)
        M = (deg + 1) ** 2
        shs = np.concatenate([rng.normal(0, 1.0, size=(Pn, 1, 3)), rng.normal(0, 0.1, size=(Pn, M - 1, 3))],
                             axis=1).astype(np.float32)
    return GaussianScene(means3D=means, scales=scales, rotations=q.astype(np.float32), opacities=opac,
                         colors=colors, shs=shs, sh_degree=deg)


def random_scene(P: int, seed: int = 0, extent: float = 0.15, scale_mean: float = 0.004,
                 sh_degree: Optional[int] = None, behind_frac: float = 0.05, huge_frac: float = 0.01):
    """Adversarial small scene for property tests: some Gaussians behind the camera plane
    (for a camera at z=-1 looking at +z), some huge, duplicated depths (ties)."""
    rng = np.random.default_rng(seed)
    means = rng.uniform(-extent, extent, size=(P, 3))
    nb = int(behind_frac * P)
    if nb:
        means[:nb, 2] = rng.uniform(-1.6, -0.9, size=nb)
    if P >= 8:
        means[P // 2: P // 2 + 3] = means[P // 2]      # exact ties in depth and position
    scales = np.exp(rng.normal(np.log(scale_mean), 0.6, size=(P, 3)))
    nh = int(huge_frac * P)
    if nh:
        scales[-nh:] *= 30.0
    q = rng.normal(size=(P, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    opac = 1.0 / (1.0 + np.exp(-rng.normal(0.5, 2.0, size=(P, 1))))
    colors = shs = None
    deg = 0
    if sh_degree is None:
        colors = rng.uniform(0, 1, size=(P, 3)).astype(np.float32)
    else:
        deg = int(sh_degree)
        M = (deg + 1) ** 2
        shs = rng.normal(0, 0.5, size=(P, M, 3)).astype(np.float32)
    return GaussianScene(means3D=means.astype(np.float32), scales=scales.astype(np.float32),
                         rotations=q.astype(np.float32), opacities=opac.astype(np.float32),
                         colors=colors, shs=shs, sh_degree=deg)


def simple_camera(H: int, W: int, dist: float = 1.0, fx: Optional[float] = None) -> CameraParams:
    fx = fx if fx is not None else 1300.0 * W / 334.0
    K = np.array([[fx, 0, W / 2 + 3.7], [0, fx, H / 2 - 2.1], [0, 0, 1]], dtype=np.float64)
    w2c = look_at_w2c(np.array([0.0, 0.0, -dist]), np.zeros(3))
    return camera_from_w2c(w2c, K, H, W)
