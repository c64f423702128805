"""Synthetic scenes and cameras for tests and bench.

Camera conventions restate reference src/gaussian_keyframe.cpp:119-204 and include/graphics_utils.h:42-50:
  world_view_transform = Rt^T (so the flat memory is column-major Rt: m[4*c + r]),
  projection P: P00 = 1/tan(fovx/2), P11 = 1/tan(fovy/2), P22 = zfar/(zfar-znear),
                P23 = -zfar*znear/(zfar-znear), P32 = 1;   full_proj = (P @ Rt)^T,
  camera_center = inverse(world_view_transform)[3, :3], fov = 2*atan(pixels / (2*focal)).
numpy only; the caller moves arrays to the device.
"""
import math

import numpy as np

C0 = 0.28209479177387814

# name -> (W, H, fx, fy)   intrinsics of the reference's shipped configs (cfg/ORB_SLAM3/...)
CAMERAS = {
    "tum": (640, 480, 520.9, 520.9),        # Monocular/TUM/tum_freiburg2_xyz.yaml:11-14,29-30
    "replica": (1200, 680, 600.0, 600.0),   # RGB-D/Replica/office0.yaml:11-14,29-30
    "euroc": (752, 480, 458.654, 457.296),  # Stereo/EuRoC/EuRoC.yaml:23-26,43-44
}


def make_camera(W, H, fx, fy, R=None, t=None, znear=0.01, zfar=100.0):
    """Returns dict(viewmatrix[16], projmatrix[16], campos[3], tanfovx, tanfovy, W, H) as float32 numpy."""
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64)
    t = np.zeros(3) if t is None else np.asarray(t, dtype=np.float64)
    Rt = np.eye(4)
    Rt[:3, :3] = R
    Rt[:3, 3] = t
    fovx = 2.0 * math.atan(W / (2.0 * fx))
    fovy = 2.0 * math.atan(H / (2.0 * fy))
    tanx, tany = math.tan(fovx * 0.5), math.tan(fovy * 0.5)
    Pm = np.zeros((4, 4))
    Pm[0, 0] = 1.0 / tanx
    Pm[1, 1] = 1.0 / tany
    Pm[3, 2] = 1.0
    Pm[2, 2] = zfar / (zfar - znear)
    Pm[2, 3] = -(zfar * znear) / (zfar - znear)
    view_T = Rt.T.astype(np.float32)                       # world_view_transform_ (row-major tensor)
    proj_T = Pm.T.astype(np.float32)                       # projection_matrix_
    full_T = (view_T.astype(np.float64) @ proj_T.astype(np.float64)).astype(np.float32)  # full_proj_transform_
    campos = np.linalg.inv(view_T.astype(np.float64))[3, :3].astype(np.float32)
    return dict(viewmatrix=view_T.reshape(-1).copy(), projmatrix=full_T.reshape(-1).copy(), campos=campos,
                tanfovx=np.float32(tanx), tanfovy=np.float32(tany), W=int(W), H=int(H), Rt=Rt)


def random_pose(rng, max_angle=0.4, max_trans=0.5):
    """Small random rigid world->camera transform (exercises general matrices; bench uses identity)."""
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = rng.uniform(-max_angle, max_angle)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    R = np.eye(3) + math.sin(ang) * K + (1 - math.cos(ang)) * (K @ K)
    t = rng.uniform(-max_trans, max_trans, size=3)
    return R, t


def make_scene(P, cam, seed=0, sh_degree_max=3, culled_frac=0.05, scale_px=2.4):
    """P Gaussians distributed as SURVEY.md §8(d) in front of `cam`. Returns raw trainer parameters
    (xyz, features_dc [P,1,3], features_rest [P,15,3], scaling (log), rotation (unnormalised), opacity (logit))."""
    rng = np.random.default_rng(seed)
    W, H = cam["W"], cam["H"]
    fx = W / (2.0 * float(cam["tanfovx"]))
    fy = H / (2.0 * float(cam["tanfovy"]))
    z = rng.uniform(0.5, 8.0, size=P)
    n_cull = int(P * culled_frac)
    if n_cull:
        z[rng.choice(P, n_cull, replace=False)] = rng.uniform(-1.0, 0.2, size=n_cull)
    u = rng.uniform(-0.1 * W, 1.1 * W, size=P)
    v = rng.uniform(-0.1 * H, 1.1 * H, size=P)
    zc = np.where(np.abs(z) < 0.05, 0.05, z)
    xc = (u - 0.5 * W) / fx * zc
    yc = (v - 0.5 * H) / fy * zc
    pc = np.stack([xc, yc, z], axis=1)
    Rt = cam["Rt"]
    Rinv = Rt[:3, :3].T
    xyz = (pc - Rt[:3, 3]) @ Rinv.T                         # camera -> world
    sigma = (scale_px / fx) * np.maximum(np.abs(z), 0.5)    # ~scale_px pixels at its depth
    scaling = np.log(sigma)[:, None] + rng.normal(0.0, 0.5, size=(P, 3))
    rotation = rng.normal(size=(P, 4))
    opacity = rng.normal(0.0, 1.5, size=(P, 1))
    rgb = rng.uniform(0.0, 1.0, size=(P, 3))
    M = (sh_degree_max + 1) ** 2
    f_dc = ((rgb - 0.5) / C0)[:, None, :]
    f_rest = rng.normal(0.0, 0.05, size=(P, M - 1, 3))
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(xyz=f32(xyz), features_dc=f32(f_dc), features_rest=f32(f_rest), scaling=f32(scaling),
                rotation=f32(rotation), opacity=f32(opacity))


def activate(scene):
    """The reference's activations (src/gaussian_model.cpp:48-71) in numpy float32: for CPU-side tests."""
    s = scene
    rot = s["rotation"] / np.maximum(np.linalg.norm(s["rotation"], axis=1, keepdims=True), 1e-12)
    return dict(means3D=s["xyz"], shs=np.concatenate([s["features_dc"], s["features_rest"]], axis=1),
                opacities=(1.0 / (1.0 + np.exp(-s["opacity"]))).astype(np.float32),
                scales=np.exp(s["scaling"]).astype(np.float32), rotations=rot.astype(np.float32))


def target_image(H, W, seed=1):
    rng = np.random.default_rng(seed)
    return rng.uniform(0.0, 1.0, size=(3, H, W)).astype(np.float32)
