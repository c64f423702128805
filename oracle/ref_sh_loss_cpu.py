"""TEST INFRASTRUCTURE / REPORTED BASELINE ONLY — the reference's LibTorch SH evaluation and loss, restated with the same ATen
ops on CPU tensors (the op library LibTorch's CPU backend dispatches to):

  eval_sh      reference include/sh_utils.h:64-136 (the `Pipeline.convert_SHs` path of gaussian_renderer.cpp:106-113)
  l1 + ssim    reference include/loss_utils.h:28-124 (ref_trainer.ssim is device-agnostic)

north_star asks for this path "timed on the host cores, core count stated, as a reported baseline": bench.py's
cpu_baseline leg calls time_sh_and_loss(). Also cross-checks photo_slam_b200/sh_utils.py in tests/test_host_cpu.py."""
import time

import torch

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def eval_sh(deg, sh, dirs):
    """sh [..., C, (deg+1)^2], dirs [..., 3] -> [..., C]; op-for-op as sh_utils.h:64-136 (deg <= 3)"""
    result = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        result = result - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            result = (result + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] +
                      C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                result = (result + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10] +
                          C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] + C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] +
                          C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] + C3[5] * z * (xx - yy) * sh[..., 14] + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return result


def render_colors(xyz, f_dc, f_rest, campos, deg=3):
    """gaussian_renderer.cpp:106-113: shs_view = features.transpose(1,2).view(-1,3,(D+1)^2); dir = normalize(xyz - campos);
    colors = clamp_min(eval_sh + 0.5, 0)"""
    feats = torch.cat((f_dc, f_rest), dim=1)
    shs_view = feats.transpose(1, 2).reshape(-1, 3, feats.size(1))
    dirs = xyz - campos.reshape(1, 3)
    dirs = dirs / dirs.norm(dim=1, keepdim=True)
    return torch.clamp_min(eval_sh(deg, shs_view, dirs) + 0.5, 0.0)


def time_sh_and_loss(P, H, W, threads, budget_s=12.0):
    """-> dict(sh_eval_ms, loss_fwd_bwd_ms, cores, sample). CPU tensors, torch.set_num_threads(threads)."""
    import ref_trainer
    old = torch.get_num_threads()
    torch.set_num_threads(threads)
    try:
        g = torch.Generator().manual_seed(0)
        xyz, f_dc, f_rest = torch.randn((P, 3), generator=g), torch.randn((P, 1, 3), generator=g), 0.05 * torch.randn((P, 15, 3), generator=g)
        campos = torch.zeros(3)
        render_colors(xyz, f_dc, f_rest, campos)
        n, t0 = 0, time.time()
        while n < 2 or (time.time() - t0 < budget_s / 2 and n < 20):
            render_colors(xyz, f_dc, f_rest, campos)
            n += 1
        sh_ms = (time.time() - t0) / n * 1e3
        img = torch.rand((3, H, W), generator=g, requires_grad=True)
        gt = torch.rand((3, H, W), generator=g)

        def loss_once():
            img.grad = None
            loss = 0.8 * (img - gt).abs().mean() + 0.2 * (1.0 - ref_trainer.ssim(img, gt))
            loss.backward()
        loss_once()
        m, t0 = 0, time.time()
        while m < 2 or (time.time() - t0 < budget_s / 2 and m < 20):
            loss_once()
            m += 1
        loss_ms = (time.time() - t0) / m * 1e3
    finally:
        torch.set_num_threads(old)
    return {"sh_eval_ms": sh_ms, "sh_eval_gaussians": P, "loss_fwd_bwd_ms": loss_ms, "image": f"{W}x{H}", "cores": threads,
            "sample": f"{n} x eval_sh(deg 3) over {P} Gaussians + {m} x (L1 + SSIM forward + autograd backward) on a {W}x{H} image, ATen CPU ops"}
