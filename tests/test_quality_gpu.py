"""-m gpu: the quality gate of BASELINE.json — "PSNR within 0.1 dB of reference after equal iterations" — on a run that exercises the whole
trainer surface: 600 iterations over 6 keyframes of a 200 k-Gaussian scene at 640x480, with densification (clone / split / prune every
100 iterations) and the SH-degree schedule, psb200 (fused step + fused densify) vs the reference chain (oracle/ref_trainer.py: the
reference's own rasterizer kernels + ATen ops + ATen densification restated from src/gaussian_model.cpp:556-815)."""
import math

import numpy as np
import pytest
import torch

import photo_slam_b200.synthetic as syn

pytestmark = pytest.mark.gpu
LRS = [0.00016, 0.0025, 0.0025 / 20, 0.05, 0.005, 0.001]


def test_psnr_within_a_tenth_of_a_db_after_600_iterations_with_densification(cuda):
    import ref_gpu
    import ref_trainer
    from photo_slam_b200 import trainer
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not built")
    P, W, H, fx = 200_000, 640, 480, 520.9
    cam0 = syn.make_camera(W, H, fx, fx)
    gt_scene = syn.make_scene(P, cam0, seed=21, scale_px=3.0)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)
    views = []
    for k in range(6):
        R, t = syn.random_pose(np.random.default_rng(50 + k), max_angle=0.06, max_trans=0.12)
        cam = syn.make_camera(W, H, fx, fx, R, t)
        views.append(dict(viewmatrix=T(cam["viewmatrix"]), projmatrix=T(cam["projmatrix"]), campos=T(cam["campos"]), tanfovx=float(cam["tanfovx"]),
                          tanfovy=float(cam["tanfovy"]), W=W, H=H))
    gt_model = trainer.GaussianModel.from_numpy(gt_scene, cuda)
    gt_tr = trainer.GaussianTrainer(gt_model)
    gts = [gt_tr.render(c).clamp(0, 1).clone() for c in views]
    # the trainee: a perturbed, thinned copy of the scene (so densification has something to do)
    rng = np.random.default_rng(3)
    keep = rng.random(P) < 0.8
    start = {k: v[keep].copy() for k, v in gt_scene.items()}
    n0 = int(keep.sum())
    start["xyz"] += rng.normal(0, 0.01, start["xyz"].shape).astype(np.float32)
    start["features_dc"] += rng.normal(0, 0.6, start["features_dc"].shape).astype(np.float32)
    start["features_rest"] *= 0
    start["opacity"] += rng.normal(0, 0.7, start["opacity"].shape).astype(np.float32)
    start["scaling"] += np.float32(0.15)

    opt = trainer.GaussianOptimizationParams()
    model = trainer.GaussianModel.from_numpy(start, cuda)
    model.trainingSetup(opt)
    model.lr_ = list(LRS)
    model.setShDegree(0)
    tr = trainer.GaussianTrainer(model, opt)
    ref = ref_trainer.RefTrainer(start, cuda, LRS, sh_degree=0)

    def psnr_of(render):
        return float(np.mean([10.0 * math.log10(1.0 / ((render(c) - g) ** 2).mean().item()) for c, g in zip(views, gts)]))

    p_start = psnr_of(lambda c: tr.render(c))
    extent, min_op, tau = 1.0, 0.005, None     # percent_dense * extent = 0.01: the scene's scales straddle it, so both clone and split fire
    gen = torch.Generator(device=cuda).manual_seed(5)
    counts = []
    for it in range(1, 601):
        c, g = views[it % 6], gts[it % 6]
        if it % 150 == 0:                                   # SH degree schedule (oneUpShDegree, gaussian_mapper.cpp:672-674)
            model.oneUpShDegree()
            ref.sh_degree = model.active_sh_degree_
        tr.trainForOneIteration(c, g)
        tr.result()
        ref.train_for_one_iteration(c, g)
        if it % 100 == 0 and it <= 500:
            if tau is None:                                 # one threshold for both sides and all rounds: top 5 % of the mean screen gradients
                grads = (model.xyz_gradient_accum_ / model.denom_).nan_to_num(0.0).squeeze()
                tau = float(torch.quantile(grads[grads > 0][:1_000_000], 0.95))
            ns_ref = __import__("ref_densify").split_count(dict(p=[t.detach() for t in ref.tensors()], accum=ref.xyz_gradient_accum.clone(),
                                                                 denom=ref.denom.clone()), tau, extent, 0.01)
            z = torch.randn((2 * max(ns_ref, model.densifySplitCount(tau, extent)), 3), device=cuda, generator=gen)
            cm = model.densifyAndPrune(tau, min_op, extent, 20, samples=z[:2 * model.densifySplitCount(tau, extent)])
            ref.densify_and_prune(tau, min_op, extent, 20, 0.01, z[:2 * ns_ref])
            counts.append((it, cm, ref.xyz.size(0)))
    torch.cuda.synchronize()
    p_psb = psnr_of(lambda c: tr.render(c))
    p_ref = psnr_of(lambda c: ref.render(c)[0].detach())
    print(f"PSNR over 6 views: start {p_start:.3f} dB -> psb200 {p_psb:.3f} dB, reference {p_ref:.3f} dB; Gaussians {n0} -> psb {model.num_points()} / ref {ref.xyz.size(0)}")
    for it, cm, nr in counts:
        print(f"  densify @{it}: psb (P_new, kept, clones, children/copy, split) = {cm}; reference P_new = {nr}")
    assert sum(cm[2] for _, cm, _ in counts) > 100 and sum(cm[3] for _, cm, _ in counts) > 100, "the run must exercise both clone and split"
    assert p_psb > p_start + 2.0 and p_ref > p_start + 2.0, "training must improve the picture"
    assert abs(p_psb - p_ref) < 0.1, (p_psb, p_ref)
    assert abs(model.num_points() - ref.xyz.size(0)) <= 0.02 * ref.xyz.size(0)
