"""`.ply` interchange in the layout of GaussianModel.save_ply (renderer_one_shot.py:120-154)."""
import numpy as np

from guassianhand_b200 import ply, scenes


def test_ply_roundtrip_and_header(tmp_path):
    sc = scenes.random_scene(257, seed=3, sh_degree=3)
    path = str(tmp_path / "g.ply")
    ply.save_ply(path, sc.means3D, sc.opacities, sc.rotations, sc.scales, sc.shs)
    head = open(path, "rb").read(2048).split(b"end_header")[0].decode().splitlines()
    props = [l.split()[2] for l in head if l.startswith("property")]
    # the reference's attribute order: x,y,z,nx,ny,nz,f_dc_0..2,f_rest_0..44,opacity,scale_0..2,rot_0..3
    assert props[:9] == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"]
    assert props[9] == "f_rest_0" and props[9 + 44] == "f_rest_44" and props[54] == "opacity"
    assert props[55:] == ["scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert "element vertex 257" in head
    g = ply.load_ply(path)
    assert np.array_equal(g["xyz"], sc.means3D) and np.array_equal(g["rotation"], sc.rotations)
    assert np.array_equal(g["shs"], sc.shs)
    assert np.allclose(g["scaling"], sc.scales, rtol=1e-6)
    assert np.allclose(g["opacity"], np.clip(sc.opacities, 1e-3, 1 - 1e-3), atol=1e-6)


def test_ply_degree0_has_no_rest(tmp_path):
    sc = scenes.random_scene(5, seed=1, sh_degree=0)
    path = str(tmp_path / "g0.ply")
    ply.save_ply(path, sc.means3D, sc.opacities, sc.rotations, sc.scales, sc.shs)
    g = ply.load_ply(path)
    assert g["shs"].shape == (5, 1, 3)
