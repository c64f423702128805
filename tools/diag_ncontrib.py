"""Diagnose n_contrib mismatches at full size."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ref_gpu
from helpers import scene_tensors
from photo_slam_b200 import rasterizer
from test_parity_ref_gpu import _export_mine
dev = torch.device("cuda:0")
P = 3_000_000
cam, sc, act, g, c = scene_tensors(P, "replica", seed=P % 97, pose_seed=None, dev=dev)
bg = torch.zeros(3, device=dev); e = torch.empty(0, device=dev)
args = (bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e, c["viewmatrix"], c["projmatrix"], c["tanfovx"], c["tanfovy"], c["H"], c["W"], g["shs"], 3, c["campos"], False)
m = rasterizer.RasterizeGaussiansCUDA(*args); r = ref_gpu.rasterize_forward(*args); torch.cuda.synchronize()
W, H = c["W"], c["H"]
mi = _export_mine(P, m[0], W, H, m[3], m[4], m[5]); ri = ref_gpu.intermediates(P, r[0], W, H, r[3], r[4], r[5])
bad = (mi["n_contrib"] != ri["n_contrib"]).nonzero().flatten()
print("env", os.environ.get("PSB_FWD_PPT"), os.environ.get("PSB_FWD_DEBUG"), "mismatching pixels:", bad.numel(), "of", W * H)
print("conic/opac equal:", torch.equal(mi["conic_opacity"][r[2] > 0], ri["conic_opacity"][r[2] > 0]), "final_T bitdiff frac", (mi["final_T"] != ri["final_T"]).float().mean().item())
for pid in bad[:6].tolist():
    py, px = divmod(pid, W)
    tile = (py // 16) * ((W + 15) // 16) + px // 16
    r0, r1 = ri["ranges"][tile].tolist()
    print(f"pixel ({px},{py}) tile {tile} list {r1 - r0}: n_contrib mine {mi['n_contrib'][pid].item()} ref {ri['n_contrib'][pid].item()} final_T mine {mi['final_T'][pid].item():.9g} ref {ri['final_T'][pid].item():.9g}")
    # replay the blend for this pixel in float64 to see what is near a threshold
    ids = ri["values_sorted"][r0:r1].long()
    xy = ri["means2D"][ids].double().cpu(); co = ri["conic_opacity"][ids].double().cpu()
    T = 1.0
    for k in range(ids.numel()):
        dx, dy = xy[k, 0].item() - px, xy[k, 1].item() - py
        power = -0.5 * (co[k, 0].item() * dx * dx + co[k, 2].item() * dy * dy) - co[k, 1].item() * dx * dy
        if power > 0: continue
        alpha = min(0.99, co[k, 3].item() * np.exp(power))
        if alpha < 1 / 255: 
            if abs(alpha - 1 / 255) < 1e-6: print(f"   k={k+1} alpha {alpha:.9f} ~ 1/255 (skip)")
            continue
        tT = T * (1 - alpha)
        if abs(alpha - 1 / 255) < 1e-6 or abs(tT - 1e-4) < 2e-9: print(f"   k={k+1} alpha {alpha:.9f} test_T {tT:.9g} T {T:.9g}")
        if tT < 1e-4: print(f"   stop at k={k+1} test_T {tT:.9g}"); break
        T = tT
