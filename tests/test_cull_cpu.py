"""Specification test of the exact conservative reach test behind the tight instance lists (csrc/psb_common.cuh: TileCull):
the same float32 arithmetic restated in numpy must never drop a tile on which the oracle's blend would accept a pixel
(alpha >= 1/255 with power <= 0, reference forward.cu:330-339), on the oracle's own conics / opacities / rectangles."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import oracle_c  # noqa: E402
from photo_slam_b200 import synthetic as syn  # noqa: E402

f32 = np.float32


def _q(A, B, C, dx, dy):  # splat_q with the kernel's rounding order
    return f32(f32(f32(B * dx) * dy) + f32(f32(0.5) * f32(f32(f32(A * dx) * dx) + f32(f32(C * dy) * dy))))


def _rect_mask(mx, my, A, B, C, op, x0, y0, x1, y1, W, H):
    """TileCull::rect_mask: facing-edge minimum of the conic quadratic per tile, one rounding pad for the rectangle."""
    area = (x1 - x0) * (y1 - y0)
    if op < f32(1.0 / 255.0):
        return [False] * area
    if not (A > 0 and C > 0 and f32(A * C) > f32(B * B)):
        return [True] * area
    thr = f32(np.log(f32(255.0) * op)) + f32(1e-3)          # = -(pmin)
    nbc, nba = f32(-B / C), f32(-B / A)
    DX = max(abs(f32(mx - f32(x0 * 16))), abs(f32(mx - f32(min(x1 * 16, W) - 1))))
    DY = max(abs(f32(my - f32(y0 * 16))), abs(f32(my - f32(min(y1 * 16, H) - 1))))
    thrp = f32(thr + f32(f32(1e-5) * _q(A, abs(B), C, DX, DY) + f32(1e-4)))
    out = []
    for ty in range(y0, y1):
        py0 = ty * 16
        dylo, dyhi = f32(my - f32(min(py0 + 16, H) - 1)), f32(my - f32(py0))
        yin = dylo <= 0 and dyhi >= 0
        ey = dylo if dylo > 0 else dyhi
        for tx in range(x0, x1):
            px0 = tx * 16
            dxlo, dxhi = f32(mx - f32(min(px0 + 16, W) - 1)), f32(mx - f32(px0))
            xin = dxlo <= 0 and dxhi >= 0
            qmin = f32(0) if (xin and yin) else f32(3e38)
            if not xin:
                ex = dxlo if dxlo > 0 else dxhi
                qmin = _q(A, B, C, ex, min(max(f32(nbc * ex), dylo), dyhi))
            if not yin:
                s2 = min(max(f32(nba * ey), dxlo), dxhi)
                qmin = min(qmin, _q(A, B, C, s2, ey))
            out.append(not (qmin > thrp))
    return out


def _check(P, wh, scale_px, seed):
    W, H = wh
    Wc, Hc, fx, fy = syn.CAMERAS["tum"]
    cam = syn.make_camera(W, H, fx * W / Wc, fy * H / Hc)
    f = oracle_c.forward(cam, syn.activate(syn.make_scene(P, cam, seed=seed, scale_px=scale_px)))
    gx, gy = (W + 15) // 16, (H + 15) // 16
    tiles = culled = 0
    for i in np.nonzero(f["radii"] > 0)[0]:
        mx, my = f32(f["means2D"][i, 0]), f32(f["means2D"][i, 1])
        A, B, C, op = [f32(v) for v in f["conic_opacity"][i]]
        r = int(f["radii"][i])
        x0, y0 = min(gx, max(0, int((mx - r) / 16))), min(gy, max(0, int((my - r) / 16)))
        x1, y1 = min(gx, max(0, int((mx + r + 15) / 16))), min(gy, max(0, int((my + r + 15) / 16)))
        if not 0 < (x1 - x0) * (y1 - y0) <= 32:
            continue
        hits = _rect_mask(mx, my, A, B, C, op, x0, y0, x1, y1, W, H)
        k = 0
        for ty in range(y0, y1):
            for tx in range(x0, x1):
                xs = np.arange(tx * 16, min(tx * 16 + 16, W), dtype=np.float64)
                ys = np.arange(ty * 16, min(ty * 16 + 16, H), dtype=np.float64)
                dx, dy = np.float64(mx) - xs[None, :], np.float64(my) - ys[:, None]
                power = -0.5 * (np.float64(A) * dx * dx + np.float64(C) * dy * dy) - np.float64(B) * dx * dy
                alpha = np.where(power > 0, 0.0, np.minimum(0.99, np.float64(op) * np.exp(power)))
                assert hits[k] or not (alpha >= 1.0 / 255.0).any(), (i, tx, ty, float(alpha.max()))
                tiles += 1
                culled += 0 if hits[k] else 1
                k += 1
    return tiles, culled


def test_reach_test_never_drops_a_contributing_tile():
    total = dropped = 0
    for P, wh, scale_px, seed in ((1500, (320, 240), 4.0, 0), (1500, (320, 240), 12.0, 1), (400, (640, 480), 40.0, 2)):
        t, c = _check(P, wh, scale_px, seed)
        total, dropped = total + t, dropped + c
    # the test is also useful: about a third of the rectangle tiles cannot be reached
    assert total > 5000 and 0.2 < dropped / total < 0.6, (total, dropped)
