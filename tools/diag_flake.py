"""Debug aid: two identical trainers stepped side by side; reports the first iteration whose images differ."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
from helpers import scene_tensors
from photo_slam_b200 import trainer

def main():
    dev = torch.device("cuda:0")
    P, wh = 40_000, (320, 240)
    cam, sc, act, g, c = scene_tensors(P, "tum", seed=0, pose_seed=None, dev=dev, wh=wh, scale_px=4.0)
    rng = np.random.default_rng(0)
    sc2 = {k: v.copy() for k, v in sc.items()}
    sc2["features_dc"] = (sc2["features_dc"] + rng.normal(0, 0.5, sc2["features_dc"].shape)).astype(np.float32)
    sc2["opacity"] = (sc2["opacity"] + rng.normal(0, 0.5, sc2["opacity"].shape)).astype(np.float32)
    gt = trainer.GaussianTrainer(trainer.GaussianModel.from_numpy(sc2, dev)).render(c).clamp(0, 1).clone()
    ms = [trainer.GaussianModel.from_numpy(sc, dev) for _ in range(2)]
    for m in ms:
        m.trainingSetup(trainer.GaussianOptimizationParams())
    ts = [trainer.GaussianTrainer(m) for m in ms]
    imgs = [torch.empty((3, wh[1], wh[0]), device=dev) for _ in range(2)]
    rad = [torch.zeros(P, dtype=torch.int32, device=dev) for _ in range(2)]
    for it in range(20):
        res = []
        for k in range(2):
            ts[k].trainForOneIteration(c, gt, out_color=imgs[k], radii=rad[k])
            res.append(ts[k].result())
        torch.cuda.synchronize()
        d = (imgs[0] - imgs[1]).abs()
        mx = d.max().item()
        if mx > 1e-3 or abs(res[0][0] - res[1][0]) > 2e-5:
            ch, y, x = np.unravel_index(int(d.argmax().item()), d.shape)
            bad = (d.max(0).values > 1e-3)
            ys, xs = torch.nonzero(bad, as_tuple=True)
            print(f"it {it}: loss {res[0][0]:.7f} vs {res[1][0]:.7f}; n {res[0][3]} vs {res[1][3]}; max diff {mx:.4f} at (x={x}, y={y}); "
                  f"{int(bad.sum())} px differ, bbox x[{int(xs.min())},{int(xs.max())}] y[{int(ys.min())},{int(ys.max())}]; "
                  f"radii differ on {int((rad[0] != rad[1]).sum())} Gaussians; max |dxyz| {float((ms[0].xyz_ - ms[1].xyz_).abs().max()):.3e}")
            dr = torch.nonzero(rad[0] != rad[1]).flatten()[:5]
            for gi in dr.tolist():
                print("   gaussian", gi, "radii", int(rad[0][gi]), int(rad[1][gi]), "xyz", ms[0].xyz_[gi].tolist(), ms[1].xyz_[gi].tolist())
            return
    print("no mismatch in 20 iterations; losses", res[0][0], res[1][0])

if __name__ == "__main__":
    main()
