"""CPU: host-side logic — camera conventions of the synthetic generator against the reference's formulas."""
import math

import numpy as np

import photo_slam_b200.synthetic as syn


def test_camera_matrices_follow_reference_conventions():
    rng = np.random.default_rng(0)
    R, t = syn.random_pose(rng)
    W, H, fx, fy = syn.CAMERAS["replica"]
    cam = syn.make_camera(W, H, fx, fy, R, t)
    vm = cam["viewmatrix"].reshape(4, 4)      # row-major tensor = Rt^T  (gaussian_keyframe.cpp:122-125)
    assert np.allclose(vm.T[:3, :3], R, atol=1e-6) and np.allclose(vm.T[:3, 3], t, atol=1e-6)
    pm = cam["projmatrix"].reshape(4, 4)      # (P @ Rt)^T  (gaussian_keyframe.cpp:138-139)
    p = np.array([0.3, -0.2, 2.0, 1.0])
    clip = p @ pm                             # row-vector convention of the transposed matrices
    cam_pt = R @ p[:3] + t
    assert np.isclose(clip[3], cam_pt[2], rtol=1e-5)                       # P32 = 1  -> w = view z
    assert np.isclose(clip[0] / clip[3], cam_pt[0] / cam_pt[2] / float(cam["tanfovx"]), rtol=1e-5)
    assert np.isclose(float(cam["tanfovx"]), W / (2 * fx), rtol=1e-6)      # tan(fov/2), fov = 2 atan(W / 2f)  (graphics_utils.h:47-50)
    assert np.allclose(R @ cam["campos"] + t, 0, atol=1e-5)                # camera centre maps to the origin of the view frame
    # memory convention used by the kernels: m[4*c + r] is row r, column c of Rt
    flat = cam["viewmatrix"]
    assert np.isclose(flat[4 * 3 + 1], t[1], atol=1e-6)


def test_scene_statistics():
    W, H, fx, fy = syn.CAMERAS["tum"]
    cam = syn.make_camera(W, H, fx, fy)
    sc = syn.make_scene(20000, cam, seed=0)
    assert sc["features_rest"].shape == (20000, 15, 3) and sc["features_dc"].shape == (20000, 1, 3)
    z = sc["xyz"][:, 2]
    assert 0.03 < np.mean(z <= 0.2) < 0.07                                  # ~5 % behind the near plane
    sig_px = np.exp(sc["scaling"]).mean(1) * fx / np.maximum(np.abs(z), 0.5)
    assert 2.0 < np.median(sig_px) < 3.5


def test_model_snapshot_restore_and_ply_roundtrip_on_cpu(tmp_path):
    """Host-side GaussianModel logic that needs no kernel: training-state snapshot / restore (used by bench.py to time every leg on
    the same iterations), LR schedule, savePly / loadPly (reference gaussian_model.cpp:838-1056) on CPU tensors."""
    import torch
    from photo_slam_b200 import trainer
    W, H, fx, fy = syn.CAMERAS["tum"]
    cam = syn.make_camera(W, H, fx, fy)
    sc = syn.make_scene(300, cam, seed=1)
    m = trainer.GaussianModel.from_numpy(sc, "cpu")
    opt = trainer.GaussianOptimizationParams()
    m.trainingSetup(opt)
    snap = m.snapshot()
    ptrs = [t.data_ptr() for t in m.tensors()]
    for t in m.tensors() + m.exp_avg_ + m.exp_avg_sq_:
        t.add_(1.0)
    m.step_, m.lr_[0] = 7, 123.0
    m.restore(snap)
    assert [t.data_ptr() for t in m.tensors()] == ptrs                      # restored in place: device pointers stay valid
    assert m.step_ == 0 and m.lr_[0] == opt.position_lr_init * m.spatial_lr_scale_
    for t, k in zip(m.tensors(), ("xyz", "features_dc", "features_rest", "opacity", "scaling", "rotation")):
        assert np.array_equal(t.numpy(), sc[k]), k
    assert not any(t.any() for t in m.exp_avg_ + m.exp_avg_sq_)
    # log-linear position learning rate (gaussian_model.cpp:1118-1131)
    assert math.isclose(m.exponLrFunc(0), m.lr_init_, rel_tol=1e-6) and math.isclose(m.exponLrFunc(m.max_steps_), m.lr_final_, rel_tol=1e-6)
    mid = m.exponLrFunc(m.max_steps_ // 2)
    assert math.isclose(mid, math.sqrt(m.lr_init_ * m.lr_final_), rel_tol=1e-3)
    path = str(tmp_path / "pc.ply")
    m.savePly(path)
    m2 = trainer.GaussianModel(3, "cpu")
    m2.loadPly(path)
    assert m2.active_sh_degree_ == 3
    for a, b in zip(m.tensors(), m2.tensors()):
        assert a.shape == b.shape and torch.equal(a, b)


def test_libtorch_cpu_sh_baseline_matches_the_host_sh_utils():
    """oracle/ref_sh_loss_cpu.py (reference include/sh_utils.h:64-136 as ATen CPU ops — the reported LibTorch-CPU baseline of bench.py)
    and photo_slam_b200/sh_utils.py (numpy) evaluate the same polynomial."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import numpy as np
    import torch
    import ref_sh_loss_cpu
    import photo_slam_b200.sh_utils as su
    rng = np.random.default_rng(0)
    sh = rng.normal(size=(200, 3, 16)).astype(np.float32)
    d = rng.normal(size=(200, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    for deg in range(4):
        a = ref_sh_loss_cpu.eval_sh(deg, torch.from_numpy(sh), torch.from_numpy(d)).numpy()
        b = su.eval_sh(deg, sh, d)
        assert np.allclose(a, b, rtol=1e-5, atol=1e-6), deg


def test_renderer_covariance_activation_matches_the_closed_form():
    """renderer.get_covariance_activation (GaussianModel::getCovarianceActivation, reference src/gaussian_model.cpp:73-101) on CPU tensors:
    Sigma = R diag(s)^2 R^T with R from the normalised quaternion, upper triangle in the rasterizer's order [xx, xy, xz, yy, yz, zz]."""
    import numpy as np
    import torch
    from photo_slam_b200 import renderer

    class M:
        pass
    rng = np.random.default_rng(0)
    m = M()
    q = rng.normal(size=(50, 4)).astype(np.float32)
    s = rng.normal(0, 0.3, size=(50, 3)).astype(np.float32)
    m.rotation_, m.scaling_ = torch.from_numpy(q), torch.from_numpy(s)
    got = renderer.get_covariance_activation(m, 1.3).numpy()
    qn = q / np.linalg.norm(q, axis=1, keepdims=True)
    w, x, y, z = qn.T
    R = np.stack([np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)], 1),
                  np.stack([2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)], 1),
                  np.stack([2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)], 1)], 1)
    S2 = (1.3 * np.exp(s)) ** 2
    cov = np.einsum("nij,nj,nkj->nik", R, S2, R)
    exp = np.stack([cov[:, 0, 0], cov[:, 0, 1], cov[:, 0, 2], cov[:, 1, 1], cov[:, 1, 2], cov[:, 2, 2]], 1)
    assert np.allclose(got, exp, rtol=1e-5, atol=1e-6)
    f = M()
    f.features_dc_, f.features_rest_ = torch.zeros(4, 1, 3), torch.ones(4, 8, 3)
    assert renderer.get_features(f).shape == (4, 9, 3)
