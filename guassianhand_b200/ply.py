"""3DGS `.ply` interchange: writes/reads exactly the vertex layout of
`GaussianModel.save_ply` (/root/reference/tgs/models/renderer_one_shot.py:120-154) so that point
clouds exchanged with the reference or any 3DGS viewer round-trip:

    x y z nx ny nz f_dc_* f_rest_* opacity scale_* rot_*      (all little-endian float32)

with opacity stored as inverse-sigmoid(clamp(o, 1e-3, 1-1e-3)), scales as log, rotations raw, SH
coefficients flattened [coeff, channel] (the reference flattens `shs[:, :1]` and `shs[:, 1:]` without
the channel-major transpose upstream 3DGS applies -- kept as the reference does it).
No plyfile dependency: binary_little_endian 1.0 is written and parsed directly."""
from typing import Dict

import numpy as np


def attribute_names(n_sh: int, n_scale: int = 3, n_rot: int = 4):
    names = ["x", "y", "z", "nx", "ny", "nz"]
    names += [f"f_dc_{i}" for i in range(3)]
    names += [f"f_rest_{i}" for i in range((n_sh - 1) * 3)]
    names.append("opacity")
    names += [f"scale_{i}" for i in range(n_scale)]
    names += [f"rot_{i}" for i in range(n_rot)]
    return names


def save_ply(path: str, xyz, opacity, rotation, scaling, shs) -> None:
    """xyz [P,3], opacity [P,1] in (0,1), rotation [P,4], scaling [P,3] > 0, shs [P,M,3]."""
    xyz = np.asarray(xyz, np.float32)
    P = xyz.shape[0]
    shs = np.asarray(shs, np.float32).reshape(P, -1, 3)
    o = np.clip(np.asarray(opacity, np.float32).reshape(P, 1), 1e-3, 1 - 1e-3)
    cols = [xyz, np.zeros_like(xyz), shs[:, :1].reshape(P, -1), shs[:, 1:].reshape(P, -1),
            np.log(o / (1 - o)), np.log(np.asarray(scaling, np.float32).reshape(P, -1)),
            np.asarray(rotation, np.float32).reshape(P, -1)]
    table = np.ascontiguousarray(np.concatenate(cols, axis=1).astype("<f4"))
    names = attribute_names(shs.shape[1], cols[5].shape[1], cols[6].shape[1])
    assert table.shape[1] == len(names)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {P}"]
    header += [f"property float {n}" for n in names]
    header.append("end_header")
    with open(path, "wb") as f:
        f.write(("\n".join(header) + "\n").encode("ascii"))
        f.write(table.tobytes())


def load_ply(path: str) -> Dict[str, np.ndarray]:
    """Inverse of save_ply: returns xyz, opacity (sigmoid applied), rotation, scaling (exp applied), shs."""
    with open(path, "rb") as f:
        data = f.read()
    end = data.index(b"end_header\n") + len(b"end_header\n")
    lines = data[:end].decode("ascii").splitlines()
    if lines[0] != "ply" or "binary_little_endian" not in lines[1]:
        raise ValueError("only binary_little_endian ply files are supported")
    n = 0
    names = []
    for ln in lines:
        t = ln.split()
        if t[:2] == ["element", "vertex"]:
            n = int(t[2])
        elif t[0] == "property":
            if t[1] not in ("float", "float32"):
                raise ValueError(f"unsupported property type {t[1]}")
            names.append(t[2])
    table = np.frombuffer(data, dtype="<f4", count=n * len(names), offset=end).reshape(n, len(names))
    col = {k: i for i, k in enumerate(names)}
    take = lambda prefix: table[:, [col[k] for k in sorted((k for k in names if k.startswith(prefix)),
                                                            key=lambda s: int(s.rsplit("_", 1)[1]))]]
    dc, rest = take("f_dc_"), take("f_rest_")
    shs = np.concatenate([dc.reshape(n, 1, 3), rest.reshape(n, -1, 3)], axis=1)
    return dict(xyz=table[:, [col["x"], col["y"], col["z"]]].copy(),
                opacity=(1.0 / (1.0 + np.exp(-table[:, [col["opacity"]]]))).astype(np.float32),
                scaling=np.exp(take("scale_")).astype(np.float32), rotation=take("rot_").copy(),
                shs=shs.astype(np.float32))
