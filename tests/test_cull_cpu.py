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


def _row_mask(mx, my, A, B, C, op, x0, y0, x1, y1, W, H):
    """TileCull::rect_mask (preprocess, tight lists): per tile ROW the reachable dx-interval of the ellipse {q <= t} in closed form,
    kept tiles = the run of tile columns meeting it. Same float32 operation order as the kernel."""
    w, area = x1 - x0, (x1 - x0) * (y1 - y0)
    if op < f32(1.0 / 255.0):
        return [False] * area
    det = f32(f32(A * C) - f32(B * B))
    if not (A > 0 and C > 0 and f32(A * C) > f32(B * B)) or not det > 0:
        return [True] * area
    thr = f32(np.log(f32(255.0) * op)) + f32(1e-3)
    DX = max(abs(f32(mx - f32(x0 * 16))), abs(f32(mx - f32(min(x1 * 16, W) - 1))))
    DY = max(abs(f32(my - f32(y0 * 16))), abs(f32(my - f32(min(y1 * 16, H) - 1))))
    t = f32(thr + f32(f32(1e-5) * _q(A, abs(B), C, DX, DY) + f32(1e-4)))
    t2 = f32(f32(2) * t)
    Y = f32(np.sqrt(f32(f32(t2 * A) / det)))
    dyr = f32(-B * f32(np.sqrt(f32(t2 / f32(det * C)))))
    invA, t2A = f32(f32(1) / A), f32(t2 * A)
    out = [False] * area
    for ty in range(y0, y1):
        py0 = ty * 16
        dylo, dyhi = f32(my - f32(min(py0 + 16, H) - 1)), f32(my - f32(py0))
        lo, hi = max(dylo, -Y), min(dyhi, Y)
        if lo > hi:
            continue
        ya, yb = min(max(dyr, lo), hi), min(max(-dyr, lo), hi)
        Da = max(f32(t2A - f32(f32(det * ya) * ya)), f32(0))
        Db = max(f32(t2A - f32(f32(det * yb) * yb)), f32(0))
        xr = f32(f32(f32(-B * ya) + f32(np.sqrt(Da))) * invA)
        xl = f32(f32(f32(-B * yb) - f32(np.sqrt(Db))) * invA)
        e = f32(f32(1e-4) * f32(max(abs(xl), abs(xr)) + f32(1)))
        xr, xl = f32(xr + e), f32(xl - e)
        tlo = max(x0, int(np.ceil(f32(f32(f32(mx - xr) - f32(15)) * f32(0.0625)))))
        thi = min(x1 - 1, int(np.floor(f32(f32(mx - xl) * f32(0.0625)))))
        for tx in range(tlo, thi + 1):
            out[(ty - y0) * w + (tx - x0)] = True
    return out


def _rect_mask(mx, my, A, B, C, op, x0, y0, x1, y1, W, H):
    """TileCull::reaches evaluated per tile (the tile kernels' tile- and warp-level culls): facing-edge minimum of the conic quadratic,
    one rounding pad for the rectangle."""
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
    tiles = culled = culled_rows = 0
    for i in np.nonzero(f["radii"] > 0)[0]:
        mx, my = f32(f["means2D"][i, 0]), f32(f["means2D"][i, 1])
        A, B, C, op = [f32(v) for v in f["conic_opacity"][i]]
        r = int(f["radii"][i])
        x0, y0 = min(gx, max(0, int((mx - r) / 16))), min(gy, max(0, int((my - r) / 16)))
        x1, y1 = min(gx, max(0, int((mx + r + 15) / 16))), min(gy, max(0, int((my + r + 15) / 16)))
        if not 0 < (x1 - x0) * (y1 - y0) <= 32:
            continue
        hits = _rect_mask(mx, my, A, B, C, op, x0, y0, x1, y1, W, H)
        rows = _row_mask(mx, my, A, B, C, op, x0, y0, x1, y1, W, H)
        k = 0
        for ty in range(y0, y1):
            for tx in range(x0, x1):
                xs = np.arange(tx * 16, min(tx * 16 + 16, W), dtype=np.float64)
                ys = np.arange(ty * 16, min(ty * 16 + 16, H), dtype=np.float64)
                dx, dy = np.float64(mx) - xs[None, :], np.float64(my) - ys[:, None]
                power = -0.5 * (np.float64(A) * dx * dx + np.float64(C) * dy * dy) - np.float64(B) * dx * dy
                alpha = np.where(power > 0, 0.0, np.minimum(0.99, np.float64(op) * np.exp(power)))
                assert hits[k] or not (alpha >= 1.0 / 255.0).any(), (i, tx, ty, float(alpha.max()))
                assert rows[k] or not (alpha >= 1.0 / 255.0).any(), ("row mask", i, tx, ty, float(alpha.max()))
                tiles += 1
                culled += 0 if hits[k] else 1
                culled_rows += 0 if rows[k] else 1
                k += 1
    return tiles, culled, culled_rows


def test_reach_test_never_drops_a_contributing_tile():
    total = dropped = dropped_rows = 0
    for P, wh, scale_px, seed in ((1500, (320, 240), 4.0, 0), (1500, (320, 240), 12.0, 1), (400, (640, 480), 40.0, 2), (800, (203, 117), 6.0, 3)):
        t, c, cr = _check(P, wh, scale_px, seed)
        total, dropped, dropped_rows = total + t, dropped + c, dropped_rows + cr
    # the tests are also useful: about a third of the rectangle tiles cannot be reached; the O(rows) form culls as well as the per-tile one
    assert total > 5000 and 0.2 < dropped / total < 0.6, (total, dropped)
    assert abs(dropped_rows - dropped) <= 0.02 * dropped, (dropped_rows, dropped)
