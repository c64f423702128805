"""Generates tests/golden/ref_small.npz ON THE GPU BOX by running the reference's own kernels
(oracle/_ref/libref_rasterizer.so = /root/reference cuda_rasterizer compiled unmodified for sm_100a).

    gpurun -- 'python tests/golden/make_golden.py && cp tests/golden/ref_small.npz gpurun_out/'

The fixture pins the CPU oracle (tests/test_oracle_golden.py) and is the committed record of what the
reference computes for these seeded inputs. Inputs are regenerated from the seeds; they are stored too so
the file is self-contained.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

import ref_gpu  # noqa: E402
from helpers import scene_tensors  # noqa: E402

CASES = {"a": dict(P=1500, wh=(128, 80), pose_seed=3, D=3, bg=(0.1, 0.2, 0.3), scale_px=4.0, seed=7),
         "b": dict(P=600, wh=(75, 41), pose_seed=8, D=1, bg=(0.0, 0.0, 0.0), scale_px=7.0, seed=9)}


def main():
    dev = torch.device("cuda:0")
    out = {}
    for name, k in CASES.items():
        cam, sc, act, g, c = scene_tensors(k["P"], "tum", seed=k["seed"], pose_seed=k["pose_seed"], dev=dev, wh=k["wh"], scale_px=k["scale_px"])
        bg = torch.tensor(k["bg"], device=dev)
        e = torch.empty(0, device=dev)
        R, color, radii, gb, bb, ib = ref_gpu.rasterize_forward(bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e,
                                                                c["viewmatrix"], c["projmatrix"], c["tanfovx"], c["tanfovy"], c["H"], c["W"],
                                                                g["shs"], k["D"], c["campos"])
        inter = ref_gpu.intermediates(k["P"], R, c["W"], c["H"], gb, bb, ib)
        gen = torch.Generator(device=dev).manual_seed(1)
        dL = torch.randn((3, c["H"], c["W"]), device=dev, generator=gen) / (3 * c["H"] * c["W"])
        grads = ref_gpu.rasterize_backward(bg, g["means3D"], radii, e, g["scales"], g["rotations"], 1.0, e, c["viewmatrix"], c["projmatrix"],
                                           c["tanfovx"], c["tanfovy"], dL, g["shs"], k["D"], c["campos"], gb, R, bb, ib)
        torch.cuda.synchronize()
        vis = (radii > 0)
        d = {"num_rendered": np.int64(R), "radii": radii, "out_color": color, "dL_dpix": dL, "D": np.int64(k["D"]), "bg": bg,
             "W": np.int64(c["W"]), "H": np.int64(c["H"])}
        for kk in ("viewmatrix", "projmatrix", "campos"):
            d[kk] = c[kk]
        d["tanfovx"], d["tanfovy"] = np.float32(c["tanfovx"]), np.float32(c["tanfovy"])
        for kk, v in g.items():
            d["in_" + kk] = v
        for kk, v in inter.items():
            if kk in ("depths", "means2D", "cov3D", "conic_opacity", "rgb", "clamped"):
                v = v.clone()
                v[~vis] = 0  # uninitialised in the reference for culled Gaussians
            d[kk] = v
        for nm, t in zip(["dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dsh", "dL_dscales", "dL_drotations", "dL_dconic"], grads):
            d[nm] = t
        for kk, v in d.items():
            out[f"{name}_{kk}"] = v.detach().cpu().numpy() if torch.is_tensor(v) else v
    path = os.path.join(HERE, "ref_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
