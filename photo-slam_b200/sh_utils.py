"""Host-side (numpy) mirror of the reference's include/sh_utils.h:64-141 and loss_utils::psnr (include/loss_utils.h:33-47):
SH evaluation as a tensor expression (the `Pipeline.convert_SHs` path, off in every shipped config), RGB<->SH DC conversion
and PSNR. The rasterizer's own SH evaluation lives in csrc/psb_geom.cuh; tests check the two against each other."""
import numpy as np

C0 = 0.28209479177387814
C1 = 0.4886025119029199
C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154, -0.4570457994644658, 1.445305721320277,
      -0.5900435899266435]


def eval_sh(deg, sh, dirs):
    """sh [..., C, (deg+1)^2], dirs [..., 3] unit vectors -> [..., C] (sh_utils.h:64-136, degrees 0-3 as the rasterizer supports)."""
    assert 0 <= deg <= 3 and sh.shape[-1] >= (deg + 1) ** 2
    r = C0 * sh[..., 0]
    if deg > 0:
        x, y, z = dirs[..., 0:1], dirs[..., 1:2], dirs[..., 2:3]
        r = r - C1 * y * sh[..., 1] + C1 * z * sh[..., 2] - C1 * x * sh[..., 3]
        if deg > 1:
            xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
            r = (r + C2[0] * xy * sh[..., 4] + C2[1] * yz * sh[..., 5] + C2[2] * (2.0 * zz - xx - yy) * sh[..., 6] +
                 C2[3] * xz * sh[..., 7] + C2[4] * (xx - yy) * sh[..., 8])
            if deg > 2:
                r = (r + C3[0] * y * (3 * xx - yy) * sh[..., 9] + C3[1] * xy * z * sh[..., 10] + C3[2] * y * (4 * zz - xx - yy) * sh[..., 11] +
                     C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[..., 12] + C3[4] * x * (4 * zz - xx - yy) * sh[..., 13] +
                     C3[5] * z * (xx - yy) * sh[..., 14] + C3[6] * x * (xx - 3 * yy) * sh[..., 15])
    return r


def RGB2SH(rgb):
    """sh_utils.h:138-141"""
    return (np.asarray(rgb) - 0.5) / C0


def SH2RGB(sh):
    return np.asarray(sh) * C0 + 0.5


def psnr(img1, img2):
    """loss_utils.h:33-37: 20 log10(1 / sqrt(mse)), mse over all elements."""
    mse = float(np.mean((np.asarray(img1, np.float64) - np.asarray(img2, np.float64)) ** 2))
    return 20.0 * np.log10(1.0 / np.sqrt(mse))
