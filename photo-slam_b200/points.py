"""Host-side mirror of the reference's point-cloud operators that ship in the same shared objects as the rasterizer:

  distCUDA2                                  reference third_party/simple-knn/spatial.h:14, spatial.cu:15-26
  transformPoints                            reference include/operate_points.h:27-29, src/operate_points.cu:73-93
  scaleAndTransformThenMarkVisiblePoints     reference include/operate_points.h:31-40, src/operate_points.cu:95-143
"""
import ctypes as C

import torch

from . import _lib
from .rasterizer import markVisible


def _L():
    L = _lib.lib()
    if not getattr(L, "_points_bound", False):
        vp = C.c_void_p
        L.psb_dist_cuda2.argtypes = [C.c_int, vp, vp, vp]
        L.psb_transform_points.argtypes = [C.c_int, vp, vp, vp, vp]
        L.psb_scale_transform_points.argtypes = [C.c_int, C.c_float, vp, vp, vp, vp, vp, vp, C.c_int, vp]
        for n in ("psb_dist_cuda2", "psb_transform_points", "psb_scale_transform_points"):
            getattr(L, n).restype = C.c_int
        L._points_bound = True
    return L


def _stream():
    return torch.cuda.current_stream().cuda_stream


def distCUDA2(points):
    """-> [P] mean squared distance to the 3 nearest neighbours."""
    P = points.size(0)
    pts = points.contiguous()
    means = torch.zeros((P,), dtype=torch.float32, device=points.device)
    _lib.check(_L().psb_dist_cuda2(P, pts.data_ptr() if P else None, means.data_ptr() if P else None, _stream()), "psb_dist_cuda2")
    return means


def transformPoints(points, transformmatrix):
    """Returns the transformed points (the reference rebinds its `points` reference argument to the new tensor)."""
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    P = points.size(0)
    out = torch.zeros_like(points)
    if P != 0:
        _lib.check(_L().psb_transform_points(P, points.contiguous().data_ptr(), transformmatrix.contiguous().data_ptr(), out.data_ptr(),
                                             _stream()), "psb_transform_points")
        return out
    return points


def scaleAndTransformThenMarkVisiblePoints(points, rots, point_not_transformed_mask, point_unstable_mask, transformmatrix, viewmatrix,
                                           projmatrix, num_transformed, scale=1.0, fix_quaternion_write=False):
    """In-place on points / rots / point_not_transformed_mask like the reference; returns the updated num_transformed."""
    if points.dim() != 2 or points.size(1) != 3:
        raise RuntimeError("points must have dimensions (num_points, 3)")
    present = markVisible(points, viewmatrix, projmatrix)
    n = present.size(0)
    if point_not_transformed_mask.size(0) != n or point_unstable_mask.size(0) != n:
        raise RuntimeError("points_mask must have dimensions (num_points)")
    final_mask = torch.logical_and(torch.logical_and(point_not_transformed_mask, point_unstable_mask), present)
    num_transformed += int(final_mask.sum().item())
    P = points.size(0)
    if P != 0:
        tp, tr = torch.zeros_like(points), torch.zeros_like(rots)
        _lib.check(_L().psb_scale_transform_points(P, float(scale), points.contiguous().data_ptr(), rots.contiguous().data_ptr(),
                                                   transformmatrix.contiguous().data_ptr(), final_mask.contiguous().data_ptr(), tp.data_ptr(),
                                                   tr.data_ptr(), int(fix_quaternion_write), _stream()), "psb_scale_transform_points")
        points[final_mask] = tp[final_mask]
        rots[final_mask] = tr[final_mask]
        point_not_transformed_mask[final_mask] = False
    return num_transformed
