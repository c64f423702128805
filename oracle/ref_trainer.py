"""TEST INFRASTRUCTURE / BASELINE ARM ONLY — the reference's training iteration, restated op for op.

The reference's host code for this path is LibTorch (C++); its arithmetic is the LibTorch op library, i.e. the
same ATen CUDA kernels PyTorch dispatches to. This module restates GaussianMapper::trainForOneIteration
(reference src/gaussian_mapper.cpp:677-772) with those ops, around the reference's OWN rasterizer kernels
(oracle/_ref/libref_rasterizer.so, compiled unmodified from /root/reference):

  activations            torch.sigmoid / exp / F.normalize / cat(clone, clone)    src/gaussian_model.cpp:48-71
  rasterizer             RasterizeGaussiansCUDA / BackwardCUDA via autograd.Function src/gaussian_rasterizer.cpp:28-180
  loss                   l1_loss + ssim (5 grouped conv2d)                          include/loss_utils.h:28-124
  densification stats    masked index_put_ / max                                    gaussian_mapper.cpp:714-719, gaussian_model.cpp:817-831
  optimizer              torch.optim.Adam(eps=1e-15), single-tensor (non-foreach, non-fused) loop like LibTorch's
                         Adam::step, 6 param groups, then zero_grad(set_to_none=True)  gaussian_model.cpp:477-503, gaussian_mapper.cpp:769-772

Used by tests (parity of the fused psb200 step) and by `bench.py --impl reference` (the timed baseline).
"""
import math

import torch
import torch.nn.functional as F

import ref_gpu


class _RefRasterize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, rs):
        R, color, radii, gb, bb, ib = ref_gpu.rasterize_forward(rs["bg"], means3D, colors_precomp, opacities, scales, rotations, 1.0,
                                                                cov3Ds_precomp, rs["viewmatrix"], rs["projmatrix"], rs["tanfovx"],
                                                                rs["tanfovy"], rs["H"], rs["W"], sh, rs["sh_degree"], rs["campos"])
        ctx.rs, ctx.R = rs, R
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, gb, bb, ib)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out, _):
        rs = ctx.rs
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, gb, bb, ib = ctx.saved_tensors
        g = ref_gpu.rasterize_backward(rs["bg"], means3D, radii, colors_precomp, scales, rotations, 1.0, cov3Ds_precomp, rs["viewmatrix"],
                                       rs["projmatrix"], rs["tanfovx"], rs["tanfovy"], grad_out, sh, rs["sh_degree"], rs["campos"], gb,
                                       ctx.R, bb, ib)
        dm2, dcol, dop, dm3, dcov, dsh, dsc, drot, _ = g
        return dm3, dm2, dsh, None, dop, dsc, drot, None, None


def create_window(device):
    gauss = torch.tensor([math.exp(-(x - 5) ** 2 / (2.0 * 1.5 * 1.5)) for x in range(11)], dtype=torch.float32, device=device)
    gauss = (gauss / gauss.sum()).unsqueeze(1)
    return gauss.mm(gauss.t()).float().unsqueeze(0).unsqueeze(0).expand(3, 1, 11, 11).contiguous()


def ssim(img1, img2):
    window = create_window(img1.device)  # the reference rebuilds the window every call (loss_utils.h:113-124)
    a, b = img1.unsqueeze(0), img2.unsqueeze(0)
    mu1 = F.conv2d(a, window, padding=5, groups=3)
    mu2 = F.conv2d(b, window, padding=5, groups=3)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(a * a, window, padding=5, groups=3) - mu1_sq
    sigma2_sq = F.conv2d(b * b, window, padding=5, groups=3) - mu2_sq
    sigma12 = F.conv2d(a * b, window, padding=5, groups=3) - mu1_mu2
    C1, C2 = 0.01 * 0.01, 0.03 * 0.03
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


class RefTrainer:
    """Reference-semantics trainer on raw parameter tensors (same layouts as photo_slam_b200.trainer.GaussianModel)."""

    def __init__(self, scene_np, device, lrs, lambda_dssim=0.2, sh_degree=3):
        t = lambda a: torch.from_numpy(a).to(device).contiguous().requires_grad_(True)
        self.xyz, self.f_dc, self.f_rest = t(scene_np["xyz"]), t(scene_np["features_dc"]), t(scene_np["features_rest"])
        self.opacity, self.scaling, self.rotation = t(scene_np["opacity"]), t(scene_np["scaling"]), t(scene_np["rotation"])
        params = [self.xyz, self.f_dc, self.f_rest, self.opacity, self.scaling, self.rotation]
        self.optimizer = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(params, lrs)], lr=0.0, eps=1e-15,
                                          foreach=False, fused=False)
        P = self.xyz.size(0)
        self.max_radii2D = torch.zeros(P, device=device)
        self.xyz_gradient_accum = torch.zeros((P, 1), device=device)
        self.denom = torch.zeros((P, 1), device=device)
        self.lambda_dssim, self.sh_degree = lambda_dssim, sh_degree
        self.bg = torch.zeros(3, device=device)

    def tensors(self):
        return [self.xyz, self.f_dc, self.f_rest, self.opacity, self.scaling, self.rotation]

    def render(self, cam):
        screenspace_points = torch.zeros_like(self.xyz, requires_grad=True)
        opacity = torch.sigmoid(self.opacity)
        scales = torch.exp(self.scaling)
        rotations = F.normalize(self.rotation)
        shs = torch.cat((self.f_dc.clone(), self.f_rest.clone()), dim=1)
        rs = dict(bg=self.bg, viewmatrix=cam["viewmatrix"], projmatrix=cam["projmatrix"], campos=cam["campos"], tanfovx=cam["tanfovx"],
                  tanfovy=cam["tanfovy"], H=cam["H"], W=cam["W"], sh_degree=self.sh_degree)
        e = torch.empty(0, device=self.xyz.device)
        color, radii = _RefRasterize.apply(self.xyz, screenspace_points, shs, e, opacity, scales, rotations, e, rs)
        return color, screenspace_points, radii > 0, radii

    def densify_and_prune(self, max_grad, min_opacity, extent, max_screen_size, percent_dense=0.01, z=None):
        """GaussianModel::densifyAndPrune (reference src/gaussian_model.cpp:795-815) on this trainer's tensors and torch.optim.Adam state,
        through the ATen restatement oracle/ref_densify.py (tensor surgery + optimizer-state surgery like :588-714). z: optional injected
        standard-normal draw for the split (else torch.randn)."""
        import ref_densify
        ps = self.tensors()
        zeros = lambda t: torch.zeros_like(t)
        stt = [self.optimizer.state.get(p, {}) for p in ps]
        st = dict(p=[p.detach() for p in ps], m=[s.get("exp_avg", zeros(p)).detach() for s, p in zip(stt, ps)],
                  v=[s.get("exp_avg_sq", zeros(p)).detach() for s, p in zip(stt, ps)], accum=self.xyz_gradient_accum, denom=self.denom,
                  max_radii=self.max_radii2D)
        steps = [s.get("step", torch.tensor(0.0)) for s in stt]
        lrs = [g["lr"] for g in self.optimizer.param_groups]
        ref_densify.densify_and_prune(st, max_grad, min_opacity, extent, max_screen_size, percent_dense, z)
        new = [t.contiguous().clone().requires_grad_(True) for t in st["p"]]
        self.xyz, self.f_dc, self.f_rest, self.opacity, self.scaling, self.rotation = new
        self.optimizer = torch.optim.Adam([{"params": [p], "lr": lr} for p, lr in zip(new, lrs)], lr=0.0, eps=1e-15, foreach=False, fused=False)
        for p, m, v, stp in zip(new, st["m"], st["v"], steps):
            self.optimizer.state[p] = {"step": stp.clone() if torch.is_tensor(stp) else torch.tensor(float(stp)), "exp_avg": m.contiguous().clone(),
                                       "exp_avg_sq": v.contiguous().clone()}
        self.max_radii2D, self.xyz_gradient_accum, self.denom = st["max_radii"], st["accum"], st["denom"]

    def train_for_one_iteration(self, cam, gt_image, mask=None, densify_stats=True, sync=True):
        image, viewspace, visibility_filter, radii = self.render(cam)
        masked = image * mask if mask is not None else image
        Ll1 = torch.abs(masked - gt_image).mean()
        loss = (1.0 - self.lambda_dssim) * Ll1 + self.lambda_dssim * (1.0 - ssim(masked, gt_image))
        loss.backward()
        if sync:
            torch.cuda.synchronize()          # gaussian_mapper.cpp:701
        loss_value = loss.item() if sync else loss.detach()   # gaussian_mapper.cpp:705
        with torch.no_grad():
            if densify_stats:
                self.max_radii2D[visibility_filter] = torch.max(self.max_radii2D[visibility_filter], radii[visibility_filter].float())
                self.xyz_gradient_accum[visibility_filter] += torch.norm(viewspace.grad[visibility_filter, :2], dim=-1, keepdim=True)
                self.denom[visibility_filter] += 1
            self.optimizer.step()
            self.optimizer.zero_grad(set_to_none=True)
        return loss_value, image.detach(), radii
