"""Debug aid: render + first training losses of the test scene (run once per PSB_TIGHT / PSB_LIB setting, compare the printed numbers)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle")]
from helpers import scene_tensors
from photo_slam_b200 import trainer
dev = torch.device("cuda:0")
P, wh = 40_000, (320, 240)
cam, sc, act, g, c = scene_tensors(P, "tum", seed=0, pose_seed=None, dev=dev, wh=wh, scale_px=4.0)
m = trainer.GaussianModel.from_numpy(sc, dev)
m.trainingSetup(trainer.GaussianOptimizationParams())
t = trainer.GaussianTrainer(m)
img = t.render(c).clone()
n0 = t.result()[3]
gt = torch.from_numpy(np.random.default_rng(5).random((3, wh[1], wh[0]), dtype=np.float32)).to(dev)
out = []
for it in range(18):
    t.trainForOneIteration(c, gt)
    r = t.result()
    out.append("%.6f/%d" % (r[0], r[3]))
print("render sum %.4f n %d | " % (img.double().sum().item(), n0) + " ".join(out))
