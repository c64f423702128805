"""Times the B1-level forward / backward of psb200 and of the reference build on one synthetic scene
(CUDA events, default stream). Usage: python tools/stage_times.py [P] [camera] [iters]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import ref_gpu  # noqa: E402
from helpers import scene_tensors  # noqa: E402
from photo_slam_b200 import rasterizer  # noqa: E402


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), float(np.min(ts))


def main():
    P = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    camname = sys.argv[2] if len(sys.argv) > 2 else "replica"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    which = sys.argv[4] if len(sys.argv) > 4 else "both"
    dev = torch.device("cuda:0")
    cam, sc, act, g, c = scene_tensors(P, camname, seed=0, pose_seed=None, dev=dev)
    bg = torch.zeros(3, device=dev)
    e = torch.empty(0, device=dev)
    args = (bg, g["means3D"], e, g["opacities"], g["scales"], g["rotations"], 1.0, e, c["viewmatrix"], c["projmatrix"], c["tanfovx"],
            c["tanfovy"], c["H"], c["W"], g["shs"], 3, c["campos"], False)
    dL = torch.randn((3, c["H"], c["W"]), device=dev) / (3 * c["H"] * c["W"])
    impls = []
    if which in ("both", "mine"):
        impls.append(("psb200", rasterizer.RasterizeGaussiansCUDA, rasterizer.RasterizeGaussiansBackwardCUDA))
    if which in ("both", "ref"):
        impls.append(("reference", ref_gpu.rasterize_forward, ref_gpu.rasterize_backward))
    for name, fwd, bwd in impls:
        out = fwd(*args)
        torch.cuda.synchronize()
        R, radii = out[0], out[2]
        vis = int((radii > 0).sum())
        bargs = (bg, g["means3D"], radii, e, g["scales"], g["rotations"], 1.0, e, c["viewmatrix"], c["projmatrix"], c["tanfovx"],
                 c["tanfovy"], dL, g["shs"], 3, c["campos"], out[3], R, out[4], out[5])
        tf = timeit(lambda: fwd(*args), iters)
        tb = timeit(lambda: bwd(*bargs), iters)
        print(f"{name}: P={P} visible={vis} num_rendered={R} ({R / max(vis, 1):.2f} tiles/visible) fwd median {tf[0]:.3f} ms (min {tf[1]:.3f}) "
              f"bwd median {tb[0]:.3f} ms (min {tb[1]:.3f})", flush=True)


if __name__ == "__main__":
    main()
