"""Shared helpers for the parity tests: scene -> device tensors, run mine / run the reference build."""
import numpy as np
import torch

import photo_slam_b200.synthetic as syn


def scene_tensors(P, camname="tum", seed=0, pose_seed=None, dev="cuda", scale=1.0, wh=None, scale_px=2.4):
    W, H, fx, fy = syn.CAMERAS[camname]
    if wh is not None:
        fx, fy = fx * wh[0] / W, fy * wh[1] / H
        W, H = wh
    R = t = None
    if pose_seed is not None:
        R, t = syn.random_pose(np.random.default_rng(pose_seed))
    cam = syn.make_camera(W, H, fx, fy, R, t)
    sc = syn.make_scene(P, cam, seed=seed, scale_px=scale_px)
    act = syn.activate(sc)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    g = {k: T(v) for k, v in act.items()}
    c = dict(viewmatrix=T(cam["viewmatrix"]), projmatrix=T(cam["projmatrix"]), campos=T(cam["campos"]),
             tanfovx=float(cam["tanfovx"]), tanfovy=float(cam["tanfovy"]), W=W, H=H)
    return cam, sc, act, g, c


def rel_close(a, b, rtol=1e-4, atol=1e-7):
    """SURVEY §8(d) float gate: |a-b| <= rtol*max(|a|,|b|) + atol. Returns fraction of elements violating it."""
    a = a.double().flatten()
    b = b.double().flatten()
    bad = (a - b).abs() > (rtol * torch.maximum(a.abs(), b.abs()) + atol)
    return bad.double().mean().item() if a.numel() else 0.0
