"""Host-side mirror of the reference's rasterizer operator interface, calling the psb200 C-ABI.

Mirrors, name for name and argument for argument:
  RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible
      reference include/rasterize_points.h:18-65, src/rasterize_points.cu:36-214
  GaussianRasterizationSettings, GaussianRasterizerFunction (here _RasterizeGaussians),
  rasterizeGaussians, GaussianRasterizer
      reference include/gaussian_rasterizer.h:25-127, src/gaussian_rasterizer.cpp:28-234

PyTorch is used only for device memory, streams and autograd bookkeeping.
"""
from dataclasses import dataclass

import torch

from . import _lib


def _ptr(t):
    """None / empty tensor -> NULL (the reference's "None" convention, gaussian_rasterizer.cpp:209-219)."""
    if t is None or t.numel() == 0:
        return None
    return t.data_ptr()


def _f32c(t):
    if t is None:
        return None
    if t.dtype != torch.float32:
        raise TypeError("psb200 rasterizer expects float32 tensors")
    return t.contiguous()


def _stream():
    return torch.cuda.current_stream().cuda_stream


class _Scratch:
    """Allocator-callback target: resizes a byte tensor on demand (reference rasterize_points.cu:28-34)."""

    def __init__(self, device):
        self.device = device
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.cb = _lib.ALLOC_FN(self._alloc)

    def _alloc(self, nbytes, _user):
        self.tensor = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self.tensor.data_ptr()


def RasterizeGaussiansCUDA(background, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp,
                           viewmatrix, projmatrix, tan_fovx, tan_fovy, image_height, image_width, sh, degree, campos,
                           prefiltered):
    """-> (num_rendered, out_color [3,H,W], radii [P] int32, geomBuffer, binningBuffer, imgBuffer)."""
    if means3D.dim() != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")  # AT_ERROR in the reference
    L = _lib.lib()
    dev = means3D.device
    P, H, W = means3D.size(0), int(image_height), int(image_width)
    out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    geom, binning, img = _Scratch(dev), _Scratch(dev), _Scratch(dev)
    rendered = 0
    if P != 0:
        M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
        tensors = [_f32c(t) for t in (background, means3D, sh, colors, opacity, scales, rotations, cov3D_precomp,
                                      viewmatrix, projmatrix, campos)]
        bg, m3, shc, col, opa, sca, rot, cov, vm, pm, cp = tensors
        rendered = _lib.check(L.psb_rasterize_forward(
            geom.cb, None, binning.cb, None, img.cb, None, P, int(degree), int(M), _ptr(bg), W, H,
            _ptr(m3), _ptr(shc), _ptr(col), _ptr(opa), _ptr(sca), float(scale_modifier), _ptr(rot), _ptr(cov),
            _ptr(vm), _ptr(pm), _ptr(cp), float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
            out_color.data_ptr(), radii.data_ptr(), _stream()), "psb_rasterize_forward")
    return rendered, out_color, radii, geom.tensor, binning.tensor, img.tensor


def RasterizeGaussiansBackwardCUDA(background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp,
                                   viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos,
                                   geomBuffer, R, binningBuffer, imageBuffer):
    """-> (dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations)."""
    L = _lib.lib()
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if (sh is not None and sh.numel() != 0) else 0
    o = dict(dtype=torch.float32, device=means3D.device)
    dL_dmeans3D = torch.zeros((P, 3), **o)
    dL_dmeans2D = torch.zeros((P, 3), **o)
    dL_dcolors = torch.zeros((P, 3), **o)
    dL_dconic = torch.zeros((P, 2, 2), **o)
    dL_dopacity = torch.zeros((P, 1), **o)
    dL_dcov3D = torch.zeros((P, 6), **o)
    dL_dsh = torch.zeros((P, M, 3), **o)
    dL_dscales = torch.zeros((P, 3), **o)
    dL_drotations = torch.zeros((P, 4), **o)
    if P != 0:
        tensors = [_f32c(t) for t in (background, means3D, sh, colors, scales, rotations, cov3D_precomp, viewmatrix,
                                      projmatrix, campos, dL_dout_color)]
        bg, m3, shc, col, sca, rot, cov, vm, pm, cp, dpix = tensors
        _lib.check(L.psb_rasterize_backward(
            P, int(degree), int(M), int(R), _ptr(bg), W, H, _ptr(m3), _ptr(shc), _ptr(col), _ptr(sca),
            float(scale_modifier), _ptr(rot), _ptr(cov), _ptr(vm), _ptr(pm), _ptr(cp), float(tan_fovx), float(tan_fovy),
            _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer), _ptr(dpix),
            dL_dmeans2D.data_ptr(), dL_dconic.data_ptr(), dL_dopacity.data_ptr(), dL_dcolors.data_ptr(),
            dL_dmeans3D.data_ptr(), dL_dcov3D.data_ptr(), _ptr(dL_dsh), dL_dscales.data_ptr(), dL_drotations.data_ptr(),
            _stream()), "psb_rasterize_backward")
    return dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations


def markVisible(means3D, viewmatrix, projmatrix):
    L = _lib.lib()
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P != 0:
        _lib.check(L.psb_mark_visible(P, _f32c(means3D).data_ptr(), _f32c(viewmatrix).data_ptr(),
                                      _f32c(projmatrix).data_ptr(), present.data_ptr(), _stream()), "psb_mark_visible")
    return present


@dataclass
class GaussianRasterizationSettings:
    """Field for field the reference struct (include/gaussian_rasterizer.h:25-55), without trailing underscores."""
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool = False


class _RasterizeGaussians(torch.autograd.Function):
    """reference GaussianRasterizerFunction (src/gaussian_rasterizer.cpp:28-180)."""

    @staticmethod
    def forward(ctx, means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
        rs = raster_settings
        num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = RasterizeGaussiansCUDA(
            rs.bg, means3D, colors_precomp, opacities, scales, rotations, rs.scale_modifier, cov3Ds_precomp,
            rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree,
            rs.campos, rs.prefiltered)
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        rs = ctx.raster_settings
        colors_precomp, means3D, scales, rotations, cov3Ds_precomp, radii, sh, geomBuffer, binningBuffer, imgBuffer = ctx.saved_tensors
        (grad_means2D, grad_colors_precomp, grad_opacities, grad_means3D, grad_cov3Ds_precomp, grad_sh, grad_scales,
         grad_rotations) = RasterizeGaussiansBackwardCUDA(
            rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier, cov3Ds_precomp, rs.viewmatrix,
            rs.projmatrix, rs.tanfovx, rs.tanfovy, grad_out_color, sh, rs.sh_degree, rs.campos, geomBuffer,
            ctx.num_rendered, binningBuffer, imgBuffer)
        # order of reference gaussian_rasterizer.cpp:159-179
        def g(t, ref):
            return t if (ref is not None and ref.numel() != 0) else None
        return (grad_means3D, grad_means2D, g(grad_sh, sh), g(grad_colors_precomp, colors_precomp), grad_opacities,
                g(grad_scales, scales), g(grad_rotations, rotations), g(grad_cov3Ds_precomp, cov3Ds_precomp), None)


def rasterize_gaussians(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp, raster_settings):
    """reference rasterizeGaussians (include/gaussian_rasterizer.h:77-99)."""
    return _RasterizeGaussians.apply(means3D, means2D, sh, colors_precomp, opacities, scales, rotations, cov3Ds_precomp,
                                     raster_settings)


class GaussianRasterizer:
    """reference GaussianRasterizer (include/gaussian_rasterizer.h:101-127, src/gaussian_rasterizer.cpp:182-234)."""

    def __init__(self, raster_settings: GaussianRasterizationSettings):
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        with torch.no_grad():
            return markVisible(positions, self.raster_settings.viewmatrix, self.raster_settings.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None):
        has = lambda t: t is not None and t.numel() != 0
        if has(shs) == has(colors_precomp):
            raise RuntimeError("Please provide excatly one of either SHs or precomputed colors!")
        if (has(scales) or has(rotations)) == has(cov3D_precomp) or (has(scales) != has(rotations)):
            raise RuntimeError("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        empty = torch.empty(0, dtype=torch.float32, device=means3D.device)
        e = lambda t: t if has(t) else empty
        return rasterize_gaussians(means3D, means2D, e(shs), e(colors_precomp), opacities, e(scales), e(rotations),
                                   e(cov3D_precomp), self.raster_settings)

    __call__ = forward
