"""-m gpu: the fused trainer step (psb_trainer_step & friends, through the C-ABI) against the reference's
iteration restated with the reference's own rasterizer kernels + the LibTorch/ATen ops (oracle/ref_trainer.py)."""
import math

import numpy as np
import pytest
import torch

import photo_slam_b200.synthetic as syn
from helpers import scene_tensors

pytestmark = pytest.mark.gpu

LRS = [0.00032, 0.0025, 0.0025 / 20, 0.05, 0.005, 0.001]


def _setup(P, wh, dev, seed=3, pose_seed=4, scale_px=5.0):
    import ref_gpu
    import ref_trainer
    from photo_slam_b200 import trainer
    if not ref_gpu.available():
        pytest.skip("oracle/_ref not built")
    cam, sc, act, g, c = scene_tensors(P, "tum", seed=seed, pose_seed=pose_seed, dev=dev, wh=wh, scale_px=scale_px)
    return cam, sc, c, trainer, ref_trainer


def _relnorm(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


@pytest.mark.parametrize("H,W,use_mask", [(96, 160, False), (117, 203, True), (16, 16, False), (480, 640, False)])
def test_fused_loss_matches_torch_ops(cuda, H, W, use_mask):
    import ref_trainer
    from photo_slam_b200 import trainer
    gen = torch.Generator(device=cuda).manual_seed(H * W)
    img = torch.rand((3, H, W), device=cuda, generator=gen)
    gt = (img + 0.2 * torch.randn((3, H, W), device=cuda, generator=gen)).clamp(0, 1)
    mask = (torch.rand((3, H, W), device=cuda, generator=gen) > 0.1).float() if use_mask else None
    torch.backends.cudnn.allow_tf32 = False  # compare against full-fp32 convolutions
    x = img.clone().requires_grad_(True)
    m = x * mask if use_mask else x
    l1 = (m - gt).abs().mean()
    ss = ref_trainer.ssim(m, gt)
    loss = 0.8 * l1 + 0.2 * (1 - ss)
    loss.backward()
    v, vl1, vss, grad = trainer.fused_loss(img, gt, mask, 0.2)
    assert abs(v - loss.item()) < 2e-6 and abs(vl1 - l1.item()) < 2e-6 and abs(vss - ss.item()) < 2e-6
    assert _relnorm(grad, x.grad) < 2e-5


def test_backward_gradients_match_reference_chain(cuda):
    """psb_trainer_backward: gradients w.r.t. the RAW parameters == autograd of the reference chain."""
    P, wh = 30_000, (320, 240)
    cam, sc, c, trainer, ref_trainer = _setup(P, wh, cuda)
    gt = torch.rand((3, wh[1], wh[0]), device=cuda)
    ref = ref_trainer.RefTrainer(sc, cuda, LRS)
    image, viewspace, vis, radii = ref.render(c)
    loss = 0.8 * (image - gt).abs().mean() + 0.2 * (1 - ref_trainer.ssim(image, gt))
    loss.backward()
    model = trainer.GaussianModel.from_numpy(sc, cuda)
    model.trainingSetup(trainer.GaussianOptimizationParams())
    before = [t.clone() for t in model.tensors()]
    dp = trainer.DataParallelTrainer(model)  # world == 1: backward -> (no all-reduce) -> Adam
    my_radii = torch.zeros(P, dtype=torch.int32, device=cuda)
    my_img = torch.zeros_like(image)
    dp.trainForOneIteration(c, gt, out_color=my_img, radii=my_radii)
    l, l1, ss, n = dp.result()
    assert abs(l - loss.item()) < 1e-5 * max(1.0, abs(loss.item()))
    assert (my_radii != radii).sum().item() <= P // 5000, "radii (activations computed in-kernel vs ATen)"
    assert _relnorm(my_img, image) < 1e-5
    for name, seg, t, b in zip(trainer.GROUPS, dp.segs, ref.tensors(), before):
        g = seg.view_as(t.grad)
        rn = _relnorm(g, t.grad)
        print(f"grad {name}: rel-norm error {rn:.2e}")
        # the reference chain itself is not run-to-run deterministic (float atomics); 2e-4 in norm is far below the
        # 1e-4 *per-element* gate applied to the rasterizer outputs in test_parity_ref_gpu.py
        assert rn < 2e-4, f"grad {name}: {rn}"
    # densification statistics
    assert (model.max_radii2D_ != torch.where(vis, radii.float(), torch.zeros_like(radii, dtype=torch.float32))).sum().item() <= P // 5000
    assert (model.denom_.flatten() != vis.float()).sum().item() <= P // 5000
    acc = torch.norm(viewspace.grad[:, :2], dim=-1) * vis
    assert _relnorm(model.xyz_gradient_accum_.flatten(), acc) < 2e-4


def test_adam_update_matches_torch_adam(cuda):
    from photo_slam_b200 import trainer
    import ctypes as C
    from photo_slam_b200 import _lib
    P = 12_345
    sc = syn.make_scene(P, syn.make_camera(64, 48, 50.0, 50.0), seed=1)
    model = trainer.GaussianModel.from_numpy(sc, cuda)
    model.trainingSetup(trainer.GaussianOptimizationParams())
    tr = trainer.GaussianTrainer(model)
    params = [t.clone().requires_grad_(True) for t in model.tensors()]
    opt = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(params, LRS)], lr=0.0, eps=1e-15, foreach=False, fused=False)
    gen = torch.Generator(device=cuda).manual_seed(0)
    for it in range(1, 5):
        grads = [torch.randn(t.shape, device=cuda, generator=gen) * (0.0 if it == 3 else 1e-3) for t in params]
        for p, g in zip(params, grads):
            p.grad = g.clone()
        opt.step()
        cm, cs = model._cmodel(), tr._cstep()
        ptrs = (C.c_void_p * 6)(*[g.data_ptr() for g in grads])
        _lib.check(tr.L.psb_adam_update(P, 16, C.byref(cm), ptrs, C.byref(cs), 1.0, None), "adam")
        model.step_ += 1
        torch.cuda.synchronize()
        # the update direction m / (sqrt(v) + eps) is O(1): agreement is measured in units of one step (lr);
        # psb200 evaluates sqrt / division on the MUFU unit (~2 ulp), torch with IEEE sequences
        for p, t, lrate in zip(params, model.tensors(), LRS):
            assert (t - p.detach()).abs().max().item() <= 2e-5 * lrate + 1e-6 * p.detach().abs().max().item(), it


def test_fused_step_equals_split_path_and_tracks_reference(cuda):
    """(a) fused backward+Adam == backward -> Adam split path; (b) several iterations stay on the reference's
    loss trajectory and end within 0.1 dB PSNR of it."""
    P, wh = 40_000, (320, 240)
    cam, sc, c, trainer, ref_trainer = _setup(P, wh, cuda, scale_px=4.0)
    # target = render of a perturbed copy of the scene, so the loss actually decreases
    rng = np.random.default_rng(0)
    sc2 = {k: v.copy() for k, v in sc.items()}
    sc2["features_dc"] = (sc2["features_dc"] + rng.normal(0, 0.5, sc2["features_dc"].shape)).astype(np.float32)
    sc2["opacity"] = (sc2["opacity"] + rng.normal(0, 0.5, sc2["opacity"].shape)).astype(np.float32)
    tgt_model = trainer.GaussianModel.from_numpy(sc2, cuda)
    gt = trainer.GaussianTrainer(tgt_model).render(c).clamp(0, 1).clone()
    torch.cuda.synchronize()

    ref = ref_trainer.RefTrainer(sc, cuda, LRS)
    fused_model = trainer.GaussianModel.from_numpy(sc, cuda)
    split_model = trainer.GaussianModel.from_numpy(sc, cuda)
    for m in (fused_model, split_model):
        m.trainingSetup(trainer.GaussianOptimizationParams())
    # split path in its pipelined (slab-wise) form: the code path bench.py --gpus N runs, minus the NCCL calls
    fused, split = trainer.GaussianTrainer(fused_model), trainer.DataParallelTrainer(split_model, pipeline=True, nslabs=3)
    # The rasterizer model itself is discontinuous (a splat's 3-sigma tile rectangle gains or loses a whole tile column when
    # ceil(radius) or a tile boundary is crossed: reference auxiliary.h:46-56, forward.cu:215-221), so trajectories that differ
    # in the last bits (float atomics, fused vs split Adam) part by ~1e-4 in loss at some iteration — in the reference too.
    # Tight agreement is therefore asserted on the first iterations, the trajectory as a whole within 1 %.
    TIGHT_ITERS = 5
    losses = []
    for it in range(25):
        lr, img_r, _ = ref.train_for_one_iteration(c, gt)
        fused.trainForOneIteration(c, gt)
        lf = fused.result()[0]
        split.trainForOneIteration(c, gt)
        ls = split.result()[0]
        losses.append((lr, lf, ls))
        if it < TIGHT_ITERS:
            assert abs(lf - ls) <= 2e-5 * max(1.0, abs(ls)), (it, lf, ls)
            assert abs(lf - lr) <= 2e-3 * abs(lr), (it, lf, lr)
        else:
            assert abs(lf - ls) <= 1e-2 * abs(ls) and abs(lf - lr) <= 1e-2 * abs(lr), (it, lr, lf, ls)
        if it == TIGHT_ITERS - 1:
            # (a) the two psb paths agree on the parameters far below the size of one Adam step
            for a, b, lrate, name in zip(fused_model.tensors(), split_model.tensors(), LRS, trainer.GROUPS):
                frac = ((a - b).abs() > 0.5 * lrate).float().mean().item()
                assert frac < 2e-3, f"{name}: {frac} of entries differ by more than half a step"
    assert losses[-1][1] < losses[0][1], "loss must decrease"
    # (b) final PSNR vs the reference's
    def psnr(img):
        return 10.0 * math.log10(1.0 / ((img - gt) ** 2).mean().item())
    p_ref = psnr(ref.render(c)[0].detach())
    p_psb = psnr(fused.render(c))
    print("PSNR reference %.3f dB, psb200 %.3f dB; first/last loss ref %.5f/%.5f psb %.5f/%.5f" %
          (p_ref, p_psb, losses[0][0], losses[-1][0], losses[0][1], losses[-1][1]))
    assert abs(p_ref - p_psb) < 0.1


def test_arena_overflow_is_a_noop_then_retried(cuda):
    from photo_slam_b200 import trainer
    P, wh = 2_000, (640, 480)
    cam, sc, act, g, c = scene_tensors(P, "tum", seed=2, pose_seed=None, dev=cuda, wh=wh, scale_px=150.0)  # ~100s of tiles per Gaussian
    gt = torch.rand((3, wh[1], wh[0]), device=cuda)
    a = trainer.GaussianModel.from_numpy(sc, cuda)
    a.trainingSetup(trainer.GaussianOptimizationParams())
    ta = trainer.GaussianTrainer(a)
    ta.trainForOneIteration(c, gt)
    la, _, _, n = ta.result()          # first attempt overflows 6*P + 65536 instances, is repeated internally
    assert n > 6 * P + 65536, "test scene must overflow the initial arena"
    assert a.step_ == 1
    b = trainer.GaussianModel.from_numpy(sc, cuda)
    b.trainingSetup(trainer.GaussianOptimizationParams())
    tb = trainer.GaussianTrainer(b)
    tb.render(c)                       # same view: grows the arena before the training step
    tb.result()
    tb.render(c)
    tb.result()
    tb.trainForOneIteration(c, gt)
    lb = tb.result()[0]
    assert abs(la - lb) < 1e-6
    for x, y, lrate in zip(a.tensors(), b.tensors(), LRS):
        assert ((x - y).abs() > 0.5 * lrate).float().mean().item() < 2e-3


def test_host_front_end_matches_device_path(cuda):
    """GaussianTrainer.trainHost (pinned host inputs, copy overlapped, loss read through the early read-back event) == the plain path."""
    from photo_slam_b200 import trainer
    P, wh = 20_000, (320, 240)
    cam, sc, act, g, c = scene_tensors(P, "tum", seed=6, pose_seed=7, dev=cuda, wh=wh, scale_px=4.0)
    a = trainer.GaussianModel.from_numpy(sc, cuda)
    b = trainer.GaussianModel.from_numpy(sc, cuda)
    for m in (a, b):
        m.trainingSetup(trainer.GaussianOptimizationParams())
    ta, tb = trainer.GaussianTrainer(a), trainer.GaussianTrainer(b)
    gts = [torch.rand((3, wh[1], wh[0])).pin_memory() for _ in range(3)]
    hostcam = dict(c, viewmatrix=c["viewmatrix"].cpu().pin_memory(), projmatrix=c["projmatrix"].cpu().pin_memory(),
                   campos=c["campos"].cpu().pin_memory())
    la, lb = [], []
    for it in range(6):
        la.append(ta.trainHost(hostcam, gts[it % 3]))
        tb.trainForOneIteration(c, gts[it % 3].to(cuda))
        lb.append(tb.result()[0])
    ta.flushHost()
    assert len(la) == 6
    for it, (x, y) in enumerate(zip(la, lb)):
        # tight while the two trajectories are still bit-close; later a tile-rectangle flip (see the trajectory test) may
        # separate them by ~1e-4
        assert abs(x - y) <= (2e-5 if it < 3 else 1e-3) * max(1.0, abs(y)), (it, x, y)
    for x, y, lrate in zip(a.tensors(), b.tensors(), LRS):
        assert ((x - y).abs() > 0.5 * lrate).float().mean().item() < 2e-3
