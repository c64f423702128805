"""-m gpu: simple-knn and operate_points replacements vs the reference build / the CPU oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 5, 1000, 1025, 60_000])
def test_dist_cuda2_matches_reference_simple_knn(cuda, n):
    import ref_gpu
    from photo_slam_b200 import points
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not built")
    gen = torch.Generator(device=cuda).manual_seed(n)
    pts = torch.randn((n, 3), device=cuda, generator=gen) * torch.tensor([3.0, 1.0, 0.3], device=cuda)
    if n >= 1000:
        pts[::7] = pts[1::7][: pts[::7].shape[0]]  # exact duplicates: zero distances
    a = points.distCUDA2(pts)
    b = ref_gpu.dist_cuda2(pts)
    torch.cuda.synchronize()
    assert torch.equal(a, b) or torch.allclose(a, b, rtol=1e-6, atol=0, equal_nan=True)


def test_dist_cuda2_matches_bruteforce_oracle(cuda):
    import oracle_c
    from photo_slam_b200 import points
    pts = torch.randn((3000, 3), device=cuda)
    a = points.distCUDA2(pts).cpu().numpy()
    assert np.allclose(a, oracle_c.knn_mean_dist2(pts.cpu().numpy()), rtol=1e-6)


def test_transform_points_and_scale_transform(cuda):
    import oracle_c
    import photo_slam_b200.synthetic as syn
    from photo_slam_b200 import points
    rng = np.random.default_rng(0)
    P = 5000
    R, t = syn.random_pose(rng, 1.0, 2.0)
    M = np.eye(4)
    M[:3, :3], M[:3, 3] = 1.7 * R, t
    m_flat = np.ascontiguousarray(M.T.reshape(-1), np.float32)  # column-major memory
    pts = rng.normal(size=(P, 3)).astype(np.float32)
    rots = rng.normal(size=(P, 4)).astype(np.float32)
    rots /= np.linalg.norm(rots, axis=1, keepdims=True)
    T = lambda a: torch.from_numpy(a).to(cuda)
    out = points.transformPoints(T(pts), T(m_flat)).cpu().numpy()
    L = oracle_c.lib()
    exp = np.zeros_like(pts)
    L.orc_transform_points(C.c_int(P), oracle_c._p(pts), oracle_c._p(m_flat), oracle_c._p(exp))
    assert np.allclose(out, exp, rtol=1e-6, atol=1e-6)
    # scale + transform + quaternion, masked by visibility in a camera
    cam = syn.make_camera(640, 480, 500.0, 500.0)
    vis = pts[:, 2] > 0.2
    ntm = rng.random(P) > 0.3
    unstable = rng.random(P) > 0.2
    for fix in (False, True):
        p_t, r_t, m_t = T(pts.copy()), T(rots.copy()), T(ntm.copy())
        n = points.scaleAndTransformThenMarkVisiblePoints(p_t, r_t, m_t, T(unstable), T(m_flat), T(cam["viewmatrix"]), T(cam["projmatrix"]), 0,
                                                          scale=1.3, fix_quaternion_write=fix)
        mask = ntm & unstable & vis
        assert n == int(mask.sum())
        ep, er = pts.copy(), rots.copy()
        op, orr = np.zeros_like(pts), np.zeros_like(rots)
        L.orc_scale_transform_points(C.c_int(P), C.c_float(1.3), oracle_c._p(pts), oracle_c._p(rots), oracle_c._p(m_flat),
                                     oracle_c._p(mask.astype(np.uint8), np.uint8), oracle_c._p(op), oracle_c._p(orr), C.c_int(0 if fix else 1))
        ep[mask], er[mask] = op[mask], orr[mask]
        assert np.allclose(p_t.cpu().numpy(), ep, rtol=1e-5, atol=1e-5)
        assert np.allclose(r_t.cpu().numpy(), er, rtol=1e-4, atol=1e-5)
        assert np.array_equal(m_t.cpu().numpy(), ntm & ~mask)
        if fix:  # the corrected write yields the rotation-composed unit quaternion (up to sign)
            q = r_t.cpu().numpy()[mask]
            assert np.all(np.isfinite(q))
