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


def test_stereo_vision_kernels_through_the_libtorch_shim(cuda):
    """reprojectDepthPinhole / monocularPinhole...NeighborhoodKeypoints / distCUDA2 / transformPoints exported by
    libcuda_rasterizer.so with the reference's C++ signatures, against numpy restatements of
    reference src/stereo_vision.cu:39-136."""
    import os
    from photo_slam_b200 import _lib, points
    shim = os.path.join(os.path.dirname(_lib.LIB_PATH), "libcuda_rasterizer.so")
    if not os.path.exists(shim):
        pytest.skip("libcuda_rasterizer.so not built")
    torch.ops.load_library(shim)
    rng = np.random.default_rng(3)
    W, H = 64, 48
    fx, fy, cx, cy = 60.0, 61.0, 31.5, 23.5
    depth = rng.uniform(0.5, 5.0, W * H).astype(np.float32)
    mask = rng.random(W * H) > 0.4
    out = torch.ops.psb200.reproject_depth_pinhole(torch.from_numpy(depth).to(cuda), torch.from_numpy(mask).to(cuda), fx, fy, cx, cy, W).cpu().numpy()
    v, u = np.divmod(np.arange(W * H), W)
    exp = np.stack([(u - cx) * depth / fx, (v - cy) * depth / fy, depth], 1).astype(np.float32) * mask[:, None]
    assert np.allclose(out, exp, rtol=1e-6, atol=1e-6)

    N = 700
    px = np.stack([rng.integers(0, W, N), rng.integers(0, H, N)], 1).astype(np.float32)
    has3d = rng.random(N) > 0.5
    pl = rng.uniform(0.5, 4.0, (N, 3)).astype(np.float32)
    colors = rng.random(W * H * 3 + 8).astype(np.float32)
    maxd = 40.0
    rp, rc = torch.ops.psb200.neighbour_depth_pinhole(torch.from_numpy(px).to(cuda), torch.from_numpy(has3d).to(cuda), torch.from_numpy(pl).to(cuda),
                                                      torch.from_numpy(colors).to(cuda), maxd, fx, fy, cx, cy, W)
    e_pt, e_col = [], []
    for i in range(N):
        pix = int(px[i, 1] * W + px[i, 0])
        if has3d[i]:
            p = pl[i]
        else:
            best, dep = np.float32(3.4e38), -1.0
            for j in range(N):
                if not has3d[j] or j == i:
                    continue
                d = np.float32((px[i, 0] - px[j, 0]) ** 2 + (px[i, 1] - px[j, 1]) ** 2)
                if d > maxd or d >= best:
                    continue
                best, dep = d, pl[j, 2]
            if dep <= 0:
                continue
            p = np.array([(int(px[i, 0]) - cx) * dep / fx, (int(px[i, 1]) - cy) * dep / fy, dep], np.float32)
        if p[2] > 0:
            e_pt.append(p)
            e_col.append(colors[pix:pix + 3])
    assert rp.shape[0] == len(e_pt)
    assert np.allclose(rp.cpu().numpy(), np.array(e_pt), rtol=1e-5, atol=1e-6) and np.allclose(rc.cpu().numpy(), np.array(e_col))

    pts = torch.randn((5000, 3), device=cuda)
    assert torch.equal(torch.ops.psb200.dist_cuda2(pts), points.distCUDA2(pts))
    m = torch.eye(4, device=cuda).flatten().contiguous()
    m[12:15] = torch.tensor([1.0, 2.0, 3.0], device=cuda)
    assert torch.allclose(torch.ops.psb200.transform_points(pts.clone(), m), pts + torch.tensor([1.0, 2.0, 3.0], device=cuda))


def _ref_points_ops():
    """torch.ops.psbref.*: the reference's OWN src/operate_points.cu / src/stereo_vision.cu kernels (compiled unmodified, oracle/Makefile
    target refpoints; façade oracle/ref_points_shim.cpp)."""
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = os.path.join(root, "oracle", "_ref", "libref_points.so")
    if not os.path.exists(lib):
        pytest.skip("oracle/_ref/libref_points.so not built (make -C oracle refpoints where /root/reference exists)")
    torch.ops.load_library(lib)
    return torch.ops.psbref


def test_point_operators_pinned_to_the_reference_kernels(cuda):
    """psb_transform_points / psb_scale_transform_points / markVisible against the reference's own kernels on the same inputs:
    transformPoints and scaleAndTransformThenMarkVisiblePoints (reference src/operate_points.cu:38-143, incl. the quaternion write
    quirk of cuda_rasterizer/operate_points.h:170-178, which fix_quaternion_write=False reproduces)."""
    import photo_slam_b200.synthetic as syn
    from photo_slam_b200 import points
    ref = _ref_points_ops()
    rng = np.random.default_rng(5)
    P = 40_003
    R, t = syn.random_pose(rng, 1.0, 2.0)
    M = np.eye(4)
    M[:3, :3], M[:3, 3] = R, t
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    m_flat = T(M.T.reshape(-1).astype(np.float32))
    pts = T(rng.normal(size=(P, 3)).astype(np.float32) * 3)
    rots = rng.normal(size=(P, 4)).astype(np.float32)
    rots = T(rots / np.linalg.norm(rots, axis=1, keepdims=True))
    a = points.transformPoints(pts.clone(), m_flat)
    b = ref.transform_points(pts.clone(), m_flat.view(4, 4))
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-6) and (a - b).abs().max().item() < 1e-5
    cam = syn.make_camera(640, 480, 500.0, 500.0)
    ntm, unstable = T(rng.random(P) > 0.3), T(rng.random(P) > 0.2)
    view, proj = T(cam["viewmatrix"]), T(cam["projmatrix"])
    pa, ra, ma = pts.clone(), rots.clone(), ntm.clone()
    na = points.scaleAndTransformThenMarkVisiblePoints(pa, ra, ma, unstable, m_flat, view, proj, 3, scale=1.3, fix_quaternion_write=False)
    pb, rb, mb, nb = ref.scale_transform_mark_visible(pts.clone(), rots.clone(), ntm.clone(), unstable, m_flat.view(4, 4), view.view(4, 4), proj.view(4, 4), 3, 1.3)
    torch.cuda.synchronize()
    assert na == nb and torch.equal(ma, mb)
    assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-5)
    assert torch.allclose(ra, rb, rtol=1e-5, atol=1e-6), (ra - rb).abs().max().item()


def test_stereo_vision_kernels_pinned_to_the_reference_kernels(cuda):
    """psb_reproject_depth_pinhole / psb_neighbour_depth_pinhole against reprojectDepthPinhole /
    monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints of reference src/stereo_vision.cu:39-215 (own kernels, compiled unmodified)."""
    import os
    from photo_slam_b200 import _lib
    ref = _ref_points_ops()
    shim = os.path.join(os.path.dirname(_lib.LIB_PATH), "libcuda_rasterizer.so")
    if not os.path.exists(shim):
        pytest.skip("libcuda_rasterizer.so not built")
    torch.ops.load_library(shim)
    rng = np.random.default_rng(11)
    W, H = 160, 120
    fx, fy, cx, cy = 150.0, 151.0, 79.5, 59.5
    depth = torch.from_numpy(rng.uniform(0.5, 5.0, W * H).astype(np.float32)).to(cuda)
    mask = torch.from_numpy(rng.random(W * H) > 0.4).to(cuda)
    a = torch.ops.psb200.reproject_depth_pinhole(depth, mask, fx, fy, cx, cy, W)
    b = ref.reproject_depth_pinhole(depth, mask, [fx, fy, cx, cy], W)
    assert torch.allclose(a, b, rtol=1e-6, atol=1e-6)
    N = 3000
    px = torch.from_numpy(np.stack([rng.integers(0, W, N), rng.integers(0, H, N)], 1).astype(np.float32)).to(cuda)
    has3d = torch.from_numpy(rng.random(N) > 0.5).to(cuda)
    pl = torch.from_numpy(rng.uniform(0.5, 4.0, (N, 3)).astype(np.float32)).to(cuda)
    colors = torch.from_numpy(rng.random(W * H * 3 + 8).astype(np.float32)).to(cuda)
    pa, ca = torch.ops.psb200.neighbour_depth_pinhole(px, has3d, pl, colors, 40.0, fx, fy, cx, cy, W)
    pb, cb = ref.neighbour_depth_pinhole(px, has3d, pl, colors, 40.0, [fx, fy, cx, cy], W)
    assert pa.shape == pb.shape and pa.shape[0] > N // 2
    assert torch.allclose(pa, pb, rtol=1e-6, atol=1e-6) and torch.equal(ca, cb)


def test_dist_cuda2_large_and_clustered(cuda):
    """The block-cooperative 3-NN search on a large, strongly non-uniform cloud (tight clusters + sparse background + duplicates) against
    the reference simple-knn build."""
    import time
    import ref_gpu
    from photo_slam_b200 import points
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not built")
    gen = torch.Generator(device=cuda).manual_seed(3)
    centres = torch.randn((40, 3), device=cuda, generator=gen) * 5
    pts = torch.cat([centres[torch.randint(0, 40, (200_000,), device=cuda, generator=gen)] + 0.01 * torch.randn((200_000, 3), device=cuda, generator=gen),
                     torch.randn((50_001, 3), device=cuda, generator=gen) * 20], dim=0)
    pts[::11] = pts[1::11][: pts[::11].shape[0]]
    a = points.distCUDA2(pts)
    torch.cuda.synchronize()
    t0 = time.time(); a = points.distCUDA2(pts); torch.cuda.synchronize(); t1 = time.time()
    b = ref_gpu.dist_cuda2(pts); torch.cuda.synchronize(); t2 = time.time()
    print(f"distCUDA2 over {pts.shape[0]} points: psb200 {1e3 * (t1 - t0):.2f} ms, reference simple-knn {1e3 * (t2 - t1):.2f} ms")
    assert torch.allclose(a, b, rtol=1e-6, atol=0)
