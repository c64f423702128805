"""CPU: the oracle's own building blocks against independent statements of the same math.

The reference's loss, Adam and activations ARE LibTorch ops (arithmetic living in LibTorch), so
the oracle's C restatements are pinned here against the same ops run by torch on CPU."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle_c
import photo_slam_b200.synthetic as syn


def _torch_loss(img, gt, lam=0.2):
    """loss_utils::l1_loss / ssim restated with torch ops exactly as reference include/loss_utils.h:28-124."""
    x = torch.from_numpy(img).clone().requires_grad_(True)
    y = torch.from_numpy(gt)
    gauss = torch.tensor([math.exp(-(i - 5) ** 2 / (2 * 1.5 ** 2)) for i in range(11)], dtype=torch.float32)
    gauss = (gauss / gauss.sum()).unsqueeze(1)
    win = gauss.mm(gauss.t()).float().unsqueeze(0).unsqueeze(0).expand(3, 1, 11, 11).contiguous()
    a, b = x.unsqueeze(0), y.unsqueeze(0)
    mu1, mu2 = F.conv2d(a, win, padding=5, groups=3), F.conv2d(b, win, padding=5, groups=3)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.conv2d(a * a, win, padding=5, groups=3) - mu1_sq
    s2 = F.conv2d(b * b, win, padding=5, groups=3) - mu2_sq
    s12 = F.conv2d(a * b, win, padding=5, groups=3) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean()
    l1 = (x - y).abs().mean()
    loss = (1.0 - lam) * l1 + lam * (1.0 - ssim)
    loss.backward()
    return loss.item(), l1.item(), ssim.item(), x.grad.numpy()


@pytest.mark.parametrize("H,W", [(37, 53), (16, 16), (8, 40)])
def test_loss_and_gradient_match_torch(H, W):
    rng = np.random.default_rng(H * W)
    img = rng.uniform(0, 1, (3, H, W)).astype(np.float32)
    gt = np.clip(img + rng.normal(0, 0.2, (3, H, W)), 0, 1).astype(np.float32)
    v, l1, ss, g = oracle_c.loss(img, gt)
    tv, tl1, tss, tg = _torch_loss(img, gt)
    assert abs(v - tv) < 1e-5 and abs(l1 - tl1) < 1e-6 and abs(ss - tss) < 1e-5
    assert np.linalg.norm(g - tg) / np.linalg.norm(tg) < 1e-4


def test_adam_matches_torch_single_tensor_adam():
    rng = np.random.default_rng(0)
    p0 = rng.normal(size=1000).astype(np.float32)
    p = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([p], lr=0.0, eps=1e-15, foreach=False, fused=False)
    opt.param_groups[0]["lr"] = 2.5e-3
    pn, m, v = p0.copy(), np.zeros_like(p0), np.zeros_like(p0)
    for t in range(1, 6):
        g = rng.normal(size=1000).astype(np.float32) * (0.0 if t == 3 else 1e-3)   # a zero-gradient step still moves p (momentum)
        p.grad = torch.from_numpy(g.copy())
        opt.step()
        pn, m, v = oracle_c.adam(pn, g, m, v, 2.5e-3, t)
        assert np.allclose(pn, p.detach().numpy(), rtol=2e-6, atol=1e-9), t


def test_activations_backward_match_torch_autograd():
    import ctypes as C
    rng = np.random.default_rng(1)
    P = 257
    o_raw = rng.normal(0, 1.5, (P, 1)).astype(np.float32)
    s_raw = rng.normal(-4, 0.5, (P, 3)).astype(np.float32)
    r_raw = rng.normal(size=(P, 4)).astype(np.float32)
    go, gs, gr = (rng.normal(size=s).astype(np.float32) for s in ((P, 1), (P, 3), (P, 4)))
    to, ts, tr = (torch.from_numpy(a.copy()).requires_grad_(True) for a in (o_raw, s_raw, r_raw))
    (torch.sigmoid(to) * torch.from_numpy(go)).sum().backward()
    (torch.exp(ts) * torch.from_numpy(gs)).sum().backward()
    (F.normalize(tr) * torch.from_numpy(gr)).sum().backward()
    L = oracle_c.lib()
    out = [np.zeros_like(o_raw), np.zeros_like(s_raw), np.zeros_like(r_raw)]
    L.orc_activations_backward(C.c_int(P), oracle_c._p(o_raw), oracle_c._p(s_raw), oracle_c._p(r_raw), oracle_c._p(go), oracle_c._p(gs),
                               oracle_c._p(gr), oracle_c._p(out[0]), oracle_c._p(out[1]), oracle_c._p(out[2]))
    for a, t in zip(out, (to, ts, tr)):
        assert np.allclose(a, t.grad.numpy(), rtol=1e-4, atol=1e-6)


def test_forward_structural_properties():
    rng = np.random.default_rng(4)
    R, t = syn.random_pose(rng)
    cam = syn.make_camera(203, 117, 165.0, 165.0, R, t)
    sc = syn.make_scene(4000, cam, seed=2, scale_px=5.0)
    act = syn.activate(sc)
    f = oracle_c.forward(cam, act, D=3, bg=(0.0, 0.0, 0.0))
    N = f["num_rendered"]
    assert N == int(f["tiles_touched"].sum()) and N > 0
    assert np.all(np.diff(f["keys"].astype(np.uint64)) >= 0) or np.all(f["keys"][1:] >= f["keys"][:-1])      # sortedness
    assert sorted(f["values"].tolist()) == sorted(np.repeat(np.arange(4000), f["tiles_touched"]).tolist())  # a permutation of the instances
    r = f["ranges"]
    touched = r[:, 1] > r[:, 0]
    assert (r[touched, 1] - r[touched, 0]).sum() == N                                                        # ranges partition the list
    assert np.all((f["radii"] > 0) == (f["tiles_touched"] > 0))
    assert f["n_contrib"].max() <= (r[:, 1] - r[:, 0]).max()
    assert np.all((f["final_T"] > 0) & (f["final_T"] <= 1))
    # culled Gaussians: behind the near plane <=> not visible
    vis = oracle_c.mark_visible(act["means3D"], cam)
    assert np.all(vis[f["radii"] > 0])
    # backward is linear in dL/dpix
    d1 = rng.normal(size=(3, 117, 203)).astype(np.float32) * 1e-3
    b1 = oracle_c.backward(cam, act, f, d1)
    b2 = oracle_c.backward(cam, act, f, 2 * d1)
    for k in b1:
        assert np.allclose(2 * b1[k], b2[k], rtol=2e-4, atol=1e-9), k
    # rows of invisible Gaussians stay zero
    inv = f["radii"] <= 0
    for k in b1:
        assert not np.any(b1[k][inv]), k


def test_knn_matches_bruteforce_numpy():
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(300, 3)).astype(np.float32)
    d = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    np.fill_diagonal(d, np.inf)
    ref = np.sort(d, axis=1)[:, :3].mean(1)
    assert np.allclose(oracle_c.knn_mean_dist2(pts), ref, rtol=1e-5)


def _densify_state(P, seed, extent):
    """A model with every densification case present: never-seen rows (0/0), clone and split candidates, transparent and oversized rows."""
    import photo_slam_b200.synthetic as syn
    rng = np.random.default_rng(seed)
    sc = syn.make_scene(P, syn.make_camera(160, 120, 130.0, 130.0), seed=seed)
    p = [sc["xyz"], sc["features_dc"], sc["features_rest"], sc["opacity"], sc["scaling"].copy(), sc["rotation"]]
    p[4][rng.choice(P, P // 50, replace=False)] = np.log(0.2 * extent)        # > 0.1 extent: world-size prune
    m = [rng.uniform(-1, 1, a.shape).astype(np.float32) for a in p]
    v = [rng.uniform(0, 1, a.shape).astype(np.float32) for a in p]
    accum = (rng.uniform(0, 0.004, (P, 1))).astype(np.float32)
    denom = np.full((P, 1), 2.0, np.float32)
    denom[::10] = 0.0
    accum[::10] = 0.0                                                          # 0/0 -> nan -> 0
    maxr = rng.uniform(0, 50, P).astype(np.float32)
    return p, m, v, accum, denom, maxr


@pytest.mark.parametrize("max_screen_size", [0, 20])
def test_densify_oracle_matches_aten_restatement(max_screen_size):
    """oracle/gs_oracle.c:orc_densify_and_prune (step-by-step C restatement of gaussian_model.cpp:588-815) pinned to the ATen ops the
    reference calls (oracle/ref_densify.py on CPU tensors): identical selection, compaction order and counts; values to 1e-6."""
    import torch
    import ref_densify
    P, extent, pd, tau, min_op = 3000, 5.0, 0.01, 0.001, 0.3
    p, m, v, accum, denom, maxr = _densify_state(P, 4, extent)
    T = lambda a: torch.from_numpy(a.copy())
    st = dict(p=[T(a) for a in p], m=[T(a) for a in m], v=[T(a) for a in v], accum=T(accum), denom=T(denom), max_radii=T(maxr))
    ns = ref_densify.split_count(dict(st, accum=T(accum), denom=T(denom)), tau, extent, pd)
    assert ns == oracle_c.densify_split_count(p[4], accum, denom, tau, extent, pd) and ns > 20
    z = np.random.default_rng(1).normal(size=(2 * ns, 3)).astype(np.float32)
    ref_densify.densify_and_prune(st, tau, min_op, extent, max_screen_size, pd, torch.from_numpy(z))
    op, om, ov = oracle_c.densify_and_prune(p, m, v, accum, denom, maxr, tau, min_op, extent, max_screen_size, pd, z)
    n = st["p"][0].shape[0]
    assert op[0].shape[0] == n and n != P
    for a, b in zip(op + om + ov, st["p"] + st["m"] + st["v"]):
        b = b.numpy()
        assert a.shape == b.shape
        assert np.allclose(a, b, rtol=1e-6, atol=1e-6), np.abs(a - b).max()
    # untouched survivors and clones are bit-exact copies; new rows have zero moments
    assert np.array_equal(op[2], st["p"][2].numpy()) and np.array_equal(om[2], st["m"][2].numpy())
    assert not st["accum"].any() and not st["max_radii"].any()


def test_reset_opacity_oracle_matches_aten():
    import torch
    import ref_densify
    o = np.random.default_rng(0).normal(0, 2, (500, 1)).astype(np.float32)
    st = ref_densify.reset_opacity(dict(p=[None, None, None, torch.from_numpy(o.copy())], m=[None] * 6, v=[None] * 6))
    assert np.allclose(oracle_c.reset_opacity(o), st["p"][3].numpy(), rtol=1e-6, atol=1e-6)


def test_philox4x32_10_known_answer_vectors():
    """oracle/gs_oracle.c:orc_philox4x32_10 — the restatement of csrc/psb_densify.cu:philox4x32_10 (same rounds, same constants) — against the
    Random123 known-answer vectors of Philox4x32-10 (Salmon et al., SC'11; kat_vectors of the Random123 distribution)."""
    import ctypes as C
    L = oracle_c.lib()

    def ph(ctr, key):
        c, k, o = (C.c_uint32 * 4)(*ctr), (C.c_uint32 * 2)(*key), (C.c_uint32 * 4)()
        L.orc_philox4x32_10(c, k, o)
        return list(o)
    assert ph([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert ph([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert ph([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_tile_backward_moment_identities():
    """The algebra behind render_bwd_kernel's round-2 formulation, checked in float64 against the reference's per-pixel expressions
    (cuda_rasterizer/backward.cu:497-547): (1) the nine per-Gaussian sums from six moments of w = G dL/dG + three colour sums;
    (2) dL/dalpha from the scalar recursion on dot(accum_rec, dL/dpixel) instead of the 3-vector recursion."""
    rng = np.random.default_rng(0)
    n = 257                                    # pixels touched by one Gaussian in one tile
    A, B, Cc, o = 0.31, -0.07, 0.22, 0.6       # conic, opacity
    W, H = 1200, 680
    ddelx_dx, ddely_dy = 0.5 * W, 0.5 * H
    dx, dy = rng.normal(0, 4, n), rng.normal(0, 4, n)
    G = np.exp(-0.5 * (A * dx * dx + Cc * dy * dy) - B * dx * dy)
    dL_dalpha = rng.normal(0, 1, n)
    dL_dG = o * dL_dalpha
    # reference, per pixel (backward.cu:518-547)
    gdx, gdy = G * dx, G * dy
    dG_ddelx = -gdx * A - gdy * B
    dG_ddely = -gdy * Cc - gdx * B
    ref = [np.sum(dL_dG * dG_ddelx * ddelx_dx), np.sum(dL_dG * dG_ddely * ddely_dy), np.sum(-0.5 * gdx * dx * dL_dG),
           np.sum(-0.5 * gdx * dy * dL_dG), np.sum(-0.5 * gdy * dy * dL_dG), np.sum(G * dL_dalpha)]
    # moments of w
    w = dL_dG * G
    M0, Mx, My, Mxx, Mxy, Myy = w.sum(), (w * dx).sum(), (w * dy).sum(), (w * dx * dx).sum(), (w * dx * dy).sum(), (w * dy * dy).sum()
    mine = [-(A * Mx + B * My) * ddelx_dx, -(Cc * My + B * Mx) * ddely_dy, -0.5 * Mxx, -0.5 * Mxy, -0.5 * Myy, M0 / o]
    assert np.allclose(mine, ref, rtol=1e-12, atol=1e-12)
    # (2) back-to-front sweep of one pixel over k contributors: vector recursion vs its dot product with dL/dpixel
    k = 40
    col, alpha = rng.uniform(0, 1, (k, 3)), rng.uniform(0.01, 0.6, k)
    dLdp = rng.normal(0, 1, 3)
    accum, last_alpha, last_color = np.zeros(3), 0.0, np.zeros(3)
    acc_dot, last_cdot = 0.0, 0.0
    for i in range(k):
        accum = last_alpha * last_color + (1.0 - last_alpha) * accum                 # backward.cu:503
        ref_dalpha = np.sum((col[i] - accum) * dLdp)                                 # :505-508 (before the * T factor)
        cdot = float(np.dot(col[i], dLdp))
        acc_dot = last_alpha * last_cdot + (1.0 - last_alpha) * acc_dot
        assert abs((cdot - acc_dot) - ref_dalpha) < 1e-12
        last_color, last_alpha, last_cdot = col[i], alpha[i], cdot
