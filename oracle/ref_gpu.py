"""TEST INFRASTRUCTURE ONLY — torch/ctypes driver of oracle/_ref/libref_rasterizer.so: the reference's own
CUDA kernels (cuda_rasterizer/*.cu, simple-knn), compiled UNMODIFIED for sm_100a by oracle/Makefile.

Used by the -m gpu parity tests (my kernels vs the real reference on the same B200), by
tests/golden/make_golden.py (pins the CPU oracle) and by `bench.py --impl reference` (the timed baseline).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libref_rasterizer.so")
ALLOC_FN = C.CFUNCTYPE(C.c_void_p, C.c_size_t, C.c_void_p)
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"{LIB_PATH} missing: run `make -C oracle ref` where /root/reference exists")
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        L.ref_forward.restype = C.c_int
        L.ref_forward.argtypes = [ALLOC_FN, vp, ALLOC_FN, vp, ALLOC_FN, vp, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int,
                                  vp, vp, vp, vp, vp, C.c_float, vp, vp, vp, vp, vp, C.c_float, C.c_float, C.c_int, vp, vp]
        L.ref_backward.restype = None
        L.ref_backward.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_float, vp, vp,
                                   vp, vp, vp, C.c_float, C.c_float, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        L.ref_mark_visible.argtypes = [C.c_int, vp, vp, vp, vp]
        L.ref_simple_knn.argtypes = [C.c_int, vp, vp]
        for n in ("ref_geom_pointers", "ref_binning_pointers", "ref_image_pointers"):
            getattr(L, n).argtypes = [vp, C.c_int, C.POINTER(C.c_void_p)]
        _lib = L
    return _lib


def _ptr(t):
    return None if (t is None or t.numel() == 0) else t.data_ptr()


class _Scratch:
    def __init__(self, device):
        self.device = device
        self.tensor = torch.empty(0, dtype=torch.uint8, device=device)
        self.cb = ALLOC_FN(self._alloc)

    def _alloc(self, n, _u):
        self.tensor = torch.empty(int(n), dtype=torch.uint8, device=self.device)
        return self.tensor.data_ptr()


def rasterize_forward(bg, means3D, colors, opacity, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                      tan_fovx, tan_fovy, H, W, sh, degree, campos, prefiltered=False):
    """Same contract as reference RasterizeGaussiansCUDA (src/rasterize_points.cu:36-114)."""
    L = lib()
    dev = means3D.device
    P = means3D.size(0)
    out_color = torch.zeros((3, H, W), dtype=torch.float32, device=dev)
    radii = torch.zeros((P,), dtype=torch.int32, device=dev)
    g, b, i = _Scratch(dev), _Scratch(dev), _Scratch(dev)
    M = sh.size(1) if (sh is not None and sh.numel()) else 0
    rendered = 0
    if P:
        rendered = L.ref_forward(g.cb, None, b.cb, None, i.cb, None, P, degree, M, _ptr(bg), W, H, _ptr(means3D), _ptr(sh),
                                 _ptr(colors), _ptr(opacity), _ptr(scales), float(scale_modifier), _ptr(rotations),
                                 _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos), float(tan_fovx),
                                 float(tan_fovy), int(prefiltered), out_color.data_ptr(), radii.data_ptr())
    return rendered, out_color, radii, g.tensor, b.tensor, i.tensor


def rasterize_backward(bg, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, viewmatrix, projmatrix,
                       tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos, geomBuffer, R, binningBuffer, imageBuffer):
    """Same contract as reference RasterizeGaussiansBackwardCUDA (src/rasterize_points.cu:116-193)."""
    L = lib()
    P = means3D.size(0)
    H, W = dL_dout_color.size(1), dL_dout_color.size(2)
    M = sh.size(1) if (sh is not None and sh.numel()) else 0
    o = dict(dtype=torch.float32, device=means3D.device)
    d = dict(dL_dmeans3D=torch.zeros((P, 3), **o), dL_dmeans2D=torch.zeros((P, 3), **o), dL_dcolors=torch.zeros((P, 3), **o),
             dL_dconic=torch.zeros((P, 2, 2), **o), dL_dopacity=torch.zeros((P, 1), **o), dL_dcov3D=torch.zeros((P, 6), **o),
             dL_dsh=torch.zeros((P, M, 3), **o), dL_dscales=torch.zeros((P, 3), **o), dL_drotations=torch.zeros((P, 4), **o))
    if P:
        L.ref_backward(P, degree, M, R, _ptr(bg), W, H, _ptr(means3D), _ptr(sh), _ptr(colors), _ptr(scales),
                       float(scale_modifier), _ptr(rotations), _ptr(cov3D_precomp), _ptr(viewmatrix), _ptr(projmatrix),
                       _ptr(campos), float(tan_fovx), float(tan_fovy), _ptr(radii), _ptr(geomBuffer), _ptr(binningBuffer),
                       _ptr(imageBuffer), dL_dout_color.contiguous().data_ptr(), d["dL_dmeans2D"].data_ptr(),
                       d["dL_dconic"].data_ptr(), d["dL_dopacity"].data_ptr(), d["dL_dcolors"].data_ptr(),
                       d["dL_dmeans3D"].data_ptr(), d["dL_dcov3D"].data_ptr(), _ptr(d["dL_dsh"]), d["dL_dscales"].data_ptr(),
                       d["dL_drotations"].data_ptr())
    return (d["dL_dmeans2D"], d["dL_dcolors"], d["dL_dopacity"], d["dL_dmeans3D"], d["dL_dcov3D"], d["dL_dsh"], d["dL_dscales"],
            d["dL_drotations"], d["dL_dconic"])


def _view(ptr, n, dtype, device):
    """Zero-copy-free snapshot: copy n elements of `dtype` starting at device address `ptr` into a new tensor."""
    out = torch.empty(n, dtype=dtype, device=device)
    if n:
        L = lib()
        L.ref_memcpy_d2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        rc = L.ref_memcpy_d2d(out.data_ptr(), ptr, n * out.element_size())
        assert rc == 0, rc
    return out


def intermediates(P, R, W, H, geomBuffer, binningBuffer, imageBuffer):
    """Reference GeometryState / BinningState / ImageState contents as tensors."""
    L = lib()
    dev = geomBuffer.device
    torch.cuda.synchronize()
    gp = (C.c_void_p * 10)()
    L.ref_geom_pointers(geomBuffer.data_ptr(), P, gp)
    o = dict(depths=_view(gp[0], P, torch.float32, dev), clamped=_view(gp[1], 3 * P, torch.uint8, dev).view(P, 3),
             means2D=_view(gp[3], 2 * P, torch.float32, dev).view(P, 2), cov3D=_view(gp[4], 6 * P, torch.float32, dev).view(P, 6),
             conic_opacity=_view(gp[5], 4 * P, torch.float32, dev).view(P, 4), rgb=_view(gp[6], 3 * P, torch.float32, dev).view(P, 3),
             point_offsets=_view(gp[7], P, torch.int32, dev), tiles_touched=_view(gp[8], P, torch.int32, dev))
    if R > 0:
        bp = (C.c_void_p * 4)()
        L.ref_binning_pointers(binningBuffer.data_ptr(), R, bp)
        o["keys_sorted"] = _view(bp[1], R, torch.int64, dev)
        o["values_sorted"] = _view(bp[3], R, torch.int32, dev)
    ip = (C.c_void_p * 3)()
    L.ref_image_pointers(imageBuffer.data_ptr(), W * H, ip)
    T = ((W + 15) // 16) * ((H + 15) // 16)
    o["ranges"] = _view(ip[0], 2 * T, torch.int32, dev).view(T, 2)
    o["n_contrib"] = _view(ip[1], W * H, torch.int32, dev)
    o["final_T"] = _view(ip[2], W * H, torch.float32, dev)
    return o


def mark_visible(means3D, viewmatrix, projmatrix):
    L = lib()
    P = means3D.size(0)
    present = torch.zeros((P,), dtype=torch.bool, device=means3D.device)
    if P:
        L.ref_mark_visible(P, means3D.data_ptr(), viewmatrix.data_ptr(), projmatrix.data_ptr(), present.data_ptr())
    return present


def dist_cuda2(points):
    """reference distCUDA2 (third_party/simple-knn/spatial.cu:15-26)."""
    L = lib()
    P = points.size(0)
    out = torch.zeros((P,), dtype=torch.float32, device=points.device)
    L.ref_simple_knn(P, points.contiguous().data_ptr(), out.data_ptr())
    return out
