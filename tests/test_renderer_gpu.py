"""-m gpu: photo_slam_b200.renderer.render — the host mirror of GaussianRenderer::render (reference src/gaussian_renderer.cpp:23-149) with every
pipeline branch (compute_cov3D, convert_SHs, override_color, SH storage other than degree 3) — against the fused trainer render and across
branches: all of them must draw the same picture of the same model, and gradients must flow to the model tensors."""
import numpy as np
import pytest
import torch

import photo_slam_b200.synthetic as syn
from helpers import scene_tensors

pytestmark = pytest.mark.gpu


def _rn(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def test_renderer_branches_agree_with_the_fused_render(cuda):
    from photo_slam_b200 import renderer, trainer
    P, wh = 30_000, (320, 240)
    cam, sc, act, g, c = scene_tensors(P, "tum", seed=12, pose_seed=13, dev=cuda, wh=wh, scale_px=4.0)
    model = trainer.GaussianModel.from_numpy(sc, cuda)
    fused = trainer.GaussianTrainer(model).render(c).clone()
    W, H = wh
    with torch.no_grad():
        img0, _, vis0, radii0 = renderer.render(c, H, W, model)
        img_sh, _, _, _ = renderer.render(c, H, W, model, renderer.GaussianPipelineParams(convert_SHs=True))
        img_cov, _, _, radii_cov = renderer.render(c, H, W, model, renderer.GaussianPipelineParams(compute_cov3D=True))
        img_both, _, _, _ = renderer.render(c, H, W, model, renderer.GaussianPipelineParams(convert_SHs=True, compute_cov3D=True))
    assert _rn(img0, fused) < 1e-5, _rn(img0, fused)                 # operator path (ATen activations) vs fused in-kernel activations
    assert _rn(img_sh, img0) < 1e-5 and _rn(img_both, img_cov) < 1e-5   # SH evaluated by the host expression vs by the kernel
    assert _rn(img_cov, img0) < 1e-4 and (radii_cov != radii0).sum().item() <= P // 2000   # covariance built by matmul vs in the kernel
    assert vis0.sum().item() > P // 2
    # scaling_modifier goes through both covariance routes identically
    with torch.no_grad():
        a = renderer.render(c, H, W, model, scaling_modifier=0.7)[0]
        b = renderer.render(c, H, W, model, renderer.GaussianPipelineParams(compute_cov3D=True), scaling_modifier=0.7)[0]
    assert _rn(b, a) < 1e-4 and _rn(a, img0) > 1e-2
    # override_color == a degree-0 model whose DC term encodes that colour
    col = torch.rand((P, 3), device=cuda)
    flat = trainer.GaussianModel.from_numpy(sc, cuda)
    flat.features_dc_ = ((col - 0.5) / 0.28209479177387814).unsqueeze(1).contiguous()
    flat.features_rest_.zero_()
    flat.setShDegree(0)
    with torch.no_grad():
        o = renderer.render(c, H, W, model, override_color=col, use_override_color=True)[0]
        d = renderer.render(c, H, W, flat)[0]
    assert _rn(o, d) < 1e-5
    # SH storage of another size (max degree 2: M = 9) takes the same path
    sc2 = dict(sc, features_rest=np.ascontiguousarray(sc["features_rest"][:, :8]))
    m2 = trainer.GaussianModel.from_numpy(sc2, cuda, sh_degree=2)
    with torch.no_grad():
        x = renderer.render(c, H, W, m2)[0]
        y = renderer.render(c, H, W, m2, renderer.GaussianPipelineParams(convert_SHs=True))[0]
    assert _rn(x, y) < 1e-5 and torch.isfinite(x).all()
    # gradients: through the operator into the raw model tensors, same for the default and the convert_SHs branch
    grads = []
    for pipe in (renderer.GaussianPipelineParams(), renderer.GaussianPipelineParams(convert_SHs=True)):
        for t in model.tensors():
            t.requires_grad_(True)
            t.grad = None
        img, viewspace, vis, radii = renderer.render(c, H, W, model, pipe)
        gt = torch.rand((3, H, W), device=cuda, generator=torch.Generator(device=cuda).manual_seed(1))
        ((img - gt) ** 2).mean().backward()
        assert viewspace.grad is not None and viewspace.grad[:, :2].abs().sum().item() > 0
        grads.append([t.grad.clone() for t in model.tensors()])
    for a, b, name in zip(grads[0], grads[1], trainer.GROUPS):
        assert a.abs().sum().item() > 0, name
        assert _rn(b, a) < 1e-3, (name, _rn(b, a))
