"""CPU tests of the on-disk formats (photo_slam_b200.io) and of the host SH utilities against the oracle."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
from photo_slam_b200 import io, sh_utils, synthetic as syn  # noqa: E402


def _scene(n=257, seed=3):
    W, H, fx, fy = syn.CAMERAS["tum"]
    cam = syn.make_camera(W, H, fx, fy)
    return cam, syn.make_scene(n, cam, seed=seed)


def test_ply_roundtrip_and_reference_layout(tmp_path):
    cam, sc = _scene()
    path = str(tmp_path / "point_cloud" / "iteration_7" / "point_cloud.ply")
    io.save_ply(path, sc["xyz"], sc["features_dc"], sc["features_rest"], sc["opacity"], sc["scaling"], sc["rotation"])
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().strip().split("\n")
    assert lines[:3] == ["ply", "format binary_little_endian 1.0", f"element vertex {sc['xyz'].shape[0]}"]
    props = [l.split()[2] for l in lines[3:]]
    # reference savePly order (gaussian_model.cpp:968-1056): xyz, normals, f_dc, f_rest, opacity, scale, rot
    assert props == ["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{i}" for i in range(45)] + \
        ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"]
    assert all(l.split()[1] == "float" for l in lines[3:])
    assert len(body) == sc["xyz"].shape[0] * 62 * 4
    tab = np.frombuffer(body, "<f4").reshape(-1, 62)
    # channel-major flattening of the SH tensors: f_rest_{c*15+k} = features_rest[:, k, c]
    assert np.array_equal(tab[:, 9 + 1], sc["features_rest"][:, 1, 0]) and np.array_equal(tab[:, 9 + 15], sc["features_rest"][:, 0, 1])
    assert np.array_equal(tab[:, 6:9], sc["features_dc"][:, 0, :]) and not tab[:, 3:6].any()
    back = io.load_ply(path)
    for k in ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation"):
        assert back[k].dtype == np.float32 and back[k].shape == sc[k].shape and np.array_equal(back[k], sc[k]), k


def test_ply_reader_accepts_ascii_and_reordered_properties(tmp_path):
    cam, sc = _scene(5)
    names = io._vertex_properties(3)
    order = names[::-1]                                   # by-name lookup like the reference's tinyply requests
    cols = {"x": sc["xyz"][:, 0], "y": sc["xyz"][:, 1], "z": sc["xyz"][:, 2], "nx": 0 * sc["xyz"][:, 0], "ny": 0 * sc["xyz"][:, 0],
            "nz": 0 * sc["xyz"][:, 0], "opacity": sc["opacity"][:, 0]}
    for c in range(3):
        cols[f"f_dc_{c}"] = sc["features_dc"][:, 0, c]
        cols[f"scale_{c}"] = sc["scaling"][:, c]
        for k in range(15):
            cols[f"f_rest_{c * 15 + k}"] = sc["features_rest"][:, k, c]
    for c in range(4):
        cols[f"rot_{c}"] = sc["rotation"][:, c]
    path = str(tmp_path / "a.ply")
    with open(path, "w") as f:
        f.write("ply\nformat ascii 1.0\ncomment test\nelement vertex 5\n" + "".join(f"property float {n}\n" for n in order) + "end_header\n")
        for i in range(5):
            f.write(" ".join(repr(float(cols[n][i])) for n in order) + "\n")
    back = io.load_ply(path)
    for k in ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation"):
        assert np.array_equal(back[k], sc[k]), k
    with open(str(tmp_path / "bad.ply"), "w") as f:
        f.write("ply\nformat ascii 1.0\nelement vertex 1\nproperty float x\nend_header\n0\n")
    with pytest.raises(ValueError):
        io.load_ply(str(tmp_path / "bad.ply"))


def test_sparse_points_ply_and_result_dir(tmp_path):
    rng = np.random.default_rng(0)
    xyz, col = rng.normal(size=(10, 3)).astype(np.float32), rng.random((10, 3)).astype(np.float32)
    cam, sc = _scene(16)
    R, t = syn.random_pose(rng, max_angle=0.3, max_trans=1.0)
    kfs = [dict(fid=4, img_name="rgb/4.png", width=640, height=480, R=R, t=t, FoVx=io.focal2fov(520.9, 640), FoVy=io.focal2fov(521.0, 480))]
    io.save_result_dir(str(tmp_path / "out"), 300, {k: sc[k] for k in ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")},
                       kfs, xyz, col, sh_degree=3, model_path="out")
    assert os.path.exists(tmp_path / "out" / "point_cloud" / "iteration_300" / "point_cloud.ply")
    raw = open(tmp_path / "out" / "input.ply", "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    assert b"property uchar red" in head and len(body) == 10 * (6 * 4 + 3)
    rec = np.frombuffer(body, np.dtype([("p", "<f4", 6), ("c", "u1", 3)]))
    assert np.array_equal(rec["p"][:, :3], xyz) and np.array_equal(rec["c"], (col * 255.0).astype(np.uint8))
    cams = json.load(open(tmp_path / "out" / "cameras.json"))
    assert cams[0]["id"] == 4 and cams[0]["width"] == 640
    # camera-to-world: position = -R^T t, rotation = R^T (gaussian_mapper.cpp:1693-1700)
    assert np.allclose(cams[0]["position"], -R.T @ t, atol=1e-5) and np.allclose(cams[0]["rotation"], R.T, atol=1e-5)
    assert abs(cams[0]["fx"] - 520.9) < 1e-3 and abs(cams[0]["fy"] - 521.0) < 1e-3
    assert open(tmp_path / "out" / "cfg_args").read() == ("Namespace(eval=False, images='images', model_path='out', resolution=-1, sh_degree=3, "
                                                          "source_path='', white_background=False, )")


def test_eval_sh_matches_the_oracle_rasterizer_colour():
    import oracle_c
    cam, sc = _scene(2000, seed=5)
    act = syn.activate(sc)
    f = oracle_c.forward(cam, act)
    vis = f["radii"] > 0
    sh = np.concatenate([sc["features_dc"], sc["features_rest"]], axis=1).transpose(0, 2, 1)       # [N,3,16]
    d = sc["xyz"] - cam["campos"][None, :]
    d = d / np.linalg.norm(d, axis=1, keepdims=True)
    rgb = np.maximum(sh_utils.eval_sh(3, sh.astype(np.float64), d.astype(np.float64)) + 0.5, 0.0)
    assert vis.sum() > 500 and np.allclose(rgb[vis], f["rgb"][vis], rtol=1e-4, atol=1e-5)
    for deg in (0, 1, 2):
        assert sh_utils.eval_sh(deg, sh, d).shape == (2000, 3)
    assert np.allclose(sh_utils.SH2RGB(sh_utils.RGB2SH(np.array([0.1, 0.5, 0.9]))), [0.1, 0.5, 0.9])
    assert abs(sh_utils.psnr(np.zeros((3, 4, 4)), np.full((3, 4, 4), 0.1)) - 20.0) < 1e-9
