"""TEST INFRASTRUCTURE ONLY — numpy/ctypes driver of oracle/liboracle.so (the CPU restatement, gs_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_lib = None

f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def build():
    src = os.path.join(_HERE, "gs_oracle.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "liboracle.so"], check=True, capture_output=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_bin.restype = C.c_int
        _lib.orc_loss.restype = C.c_float
    return _lib


def _p(a, dtype=np.float32):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=dtype)
    return a.ctypes.data_as(C.c_void_p)


def forward(cam, g, D=3, bg=(0.0, 0.0, 0.0), scale_modifier=1.0, colors_precomp=None, cov3D_precomp=None):
    """g: dict(means3D, shs [P,M,3], opacities, scales, rotations) float32 numpy (activated values).
    Returns dict with every intermediate of the reference forward."""
    L = lib()
    means3D = np.ascontiguousarray(g["means3D"], np.float32)
    P = means3D.shape[0]
    shs = None if colors_precomp is not None else np.ascontiguousarray(g["shs"], np.float32)
    M = 0 if shs is None else shs.shape[1]
    W, H = cam["W"], cam["H"]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    o = dict(radii=np.zeros(P, np.int32), means2D=np.zeros((P, 2), np.float32), depths=np.zeros(P, np.float32),
             cov3D=np.zeros((P, 6), np.float32), rgb=np.zeros((P, 3), np.float32),
             conic_opacity=np.zeros((P, 4), np.float32), tiles_touched=np.zeros(P, np.uint32),
             clamped=np.zeros((P, 3), np.uint8))
    keep = [means3D, shs]
    args = [C.c_int(P), C.c_int(D), C.c_int(M), _p(means3D), _p(g.get("scales")), C.c_float(scale_modifier),
            _p(g.get("rotations")), _p(g["opacities"]), _p(shs), _p(cov3D_precomp), _p(colors_precomp),
            _p(cam["viewmatrix"]), _p(cam["projmatrix"]), _p(cam["campos"]), C.c_int(W), C.c_int(H),
            C.c_float(cam["tanfovx"]), C.c_float(cam["tanfovy"]),
            _p(o["radii"], np.int32), _p(o["means2D"]), _p(o["depths"]), _p(o["cov3D"]), _p(o["rgb"]),
            _p(o["conic_opacity"]), _p(o["tiles_touched"], np.uint32), _p(o["clamped"], np.uint8)]
    L.orc_preprocess(*args)
    if cov3D_precomp is not None:
        o["cov3D"] = np.ascontiguousarray(cov3D_precomp, np.float32)
    colors = o["rgb"] if colors_precomp is None else np.ascontiguousarray(colors_precomp, np.float32)
    o["colors"] = colors
    o["point_offsets"] = np.zeros(P, np.uint32)
    N = int(o["tiles_touched"].sum())
    o["keys"] = np.zeros(max(N, 1), np.uint64)
    o["values"] = np.zeros(max(N, 1), np.uint32)
    o["ranges"] = np.zeros((gx * gy, 2), np.uint32)
    n2 = L.orc_bin(C.c_int(P), C.c_int(W), C.c_int(H), _p(o["radii"], np.int32), _p(o["means2D"]), _p(o["depths"]),
                   _p(o["tiles_touched"], np.uint32), _p(o["point_offsets"], np.uint32), _p(o["keys"], np.uint64),
                   _p(o["values"], np.uint32), _p(o["ranges"], np.uint32))
    assert n2 == N
    o["num_rendered"] = N
    o["keys"], o["values"] = o["keys"][:N], o["values"][:N]
    o["out_color"] = np.zeros((3, H, W), np.float32)
    o["final_T"] = np.zeros(H * W, np.float32)
    o["n_contrib"] = np.zeros(H * W, np.uint32)
    bg = np.asarray(bg, np.float32)
    L.orc_render_forward(C.c_int(W), C.c_int(H), _p(o["ranges"], np.uint32), _p(o["values"], np.uint32), _p(o["means2D"]),
                         _p(colors), _p(o["conic_opacity"]), _p(bg), _p(o["out_color"]), _p(o["final_T"]),
                         _p(o["n_contrib"], np.uint32))
    o["bg"] = bg
    del keep
    return o


def backward(cam, g, fwd, dL_dpix, D=3, scale_modifier=1.0, colors_precomp=None, cov3D_precomp=None):
    L = lib()
    means3D = np.ascontiguousarray(g["means3D"], np.float32)
    P = means3D.shape[0]
    shs = None if colors_precomp is not None else np.ascontiguousarray(g["shs"], np.float32)
    M = 0 if shs is None else shs.shape[1]
    W, H = cam["W"], cam["H"]
    o = dict(dL_dmean2D=np.zeros((P, 3), np.float32), dL_dconic=np.zeros((P, 4), np.float32),
             dL_dopacity=np.zeros((P, 1), np.float32), dL_dcolor=np.zeros((P, 3), np.float32),
             dL_dmean3D=np.zeros((P, 3), np.float32), dL_dcov3D=np.zeros((P, 6), np.float32),
             dL_dsh=np.zeros((P, max(M, 1), 3), np.float32), dL_dscale=np.zeros((P, 3), np.float32),
             dL_drot=np.zeros((P, 4), np.float32))
    dL_dpix = np.ascontiguousarray(dL_dpix, np.float32)
    L.orc_render_backward(C.c_int(P), C.c_int(W), C.c_int(H), _p(fwd["ranges"], np.uint32), _p(fwd["values"], np.uint32),
                          _p(fwd["bg"]), _p(fwd["means2D"]), _p(fwd["conic_opacity"]), _p(fwd["colors"]),
                          _p(fwd["final_T"]), _p(fwd["n_contrib"], np.uint32), _p(dL_dpix), _p(o["dL_dmean2D"]),
                          _p(o["dL_dconic"]), _p(o["dL_dopacity"]), _p(o["dL_dcolor"]))
    L.orc_preprocess_backward(
        C.c_int(P), C.c_int(D), C.c_int(M), _p(means3D), _p(fwd["radii"], np.int32), _p(shs), _p(fwd["clamped"], np.uint8),
        _p(g.get("scales") if cov3D_precomp is None else None), _p(g.get("rotations") if cov3D_precomp is None else None),
        C.c_float(scale_modifier), _p(fwd["cov3D"]), _p(cam["viewmatrix"]), _p(cam["projmatrix"]), C.c_int(W), C.c_int(H),
        C.c_float(cam["tanfovx"]), C.c_float(cam["tanfovy"]), _p(cam["campos"]), _p(o["dL_dmean2D"]), _p(o["dL_dconic"]),
        _p(o["dL_dmean3D"]), _p(o["dL_dcolor"]), _p(o["dL_dcov3D"]), _p(o["dL_dsh"]), _p(o["dL_dscale"]), _p(o["dL_drot"]))
    if M == 0:
        o["dL_dsh"] = np.zeros((P, 0, 3), np.float32)
    return o


def loss(img, gt, lambda_dssim=0.2, want_grad=True):
    L = lib()
    img = np.ascontiguousarray(img, np.float32)
    gt = np.ascontiguousarray(gt, np.float32)
    _, H, W = img.shape
    l1, ss = C.c_float(), C.c_float()
    grad = np.zeros_like(img) if want_grad else None
    val = L.orc_loss(C.c_int(H), C.c_int(W), _p(img), _p(gt), C.c_float(lambda_dssim), C.byref(l1), C.byref(ss), _p(grad))
    return float(val), float(l1.value), float(ss.value), grad


def adam(p, g, m, v, lr, step_t, beta1=0.9, beta2=0.999, eps=1e-15):
    L = lib()
    p, m, v = (np.ascontiguousarray(a, np.float32).copy() for a in (p, m, v))
    g = np.ascontiguousarray(g, np.float32)
    L.orc_adam(C.c_size_t(p.size), _p(p), _p(g), _p(m), _p(v), C.c_float(lr), C.c_float(beta1), C.c_float(beta2),
               C.c_float(eps), C.c_int(step_t))
    return p, m, v


def mark_visible(means3D, cam):
    L = lib()
    means3D = np.ascontiguousarray(means3D, np.float32)
    out = np.zeros(means3D.shape[0], np.uint8)
    L.orc_mark_visible(C.c_int(means3D.shape[0]), _p(means3D), _p(cam["viewmatrix"]), _p(cam["projmatrix"]), _p(out, np.uint8))
    return out.astype(bool)


def knn_mean_dist2(pts):
    L = lib()
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.zeros(pts.shape[0], np.float32)
    L.orc_knn_mean_dist2(C.c_int(pts.shape[0]), _p(pts), _p(out))
    return out


DN_K = (3, 3, 45, 1, 3, 4)


def densify_split_count(scaling, accum, denom, max_grad, extent, percent_dense):
    L = lib()
    L.orc_densify_split_count.restype = C.c_int
    P = scaling.shape[0]
    return L.orc_densify_split_count(C.c_int(P), _p(scaling), _p(accum), _p(denom), C.c_float(max_grad), C.c_float(extent), C.c_float(percent_dense))


def densify_and_prune(p, m, v, accum, denom, max_radii, max_grad, min_opacity, extent, max_screen_size, percent_dense, z):
    """p/m/v: lists of 6 float32 arrays [P, ...] (reference shapes). Returns (p, m, v) lists of the densified model."""
    L = lib()
    L.orc_densify_and_prune.restype = C.c_int
    P = p[0].shape[0]
    ins = [np.ascontiguousarray(a, np.float32) for a in list(p) + list(m) + list(v)]
    outs = [np.zeros((2 * P + 1) * DN_K[i % 6], np.float32) for i in range(18)]
    in_ptrs = (C.c_void_p * 18)(*[a.ctypes.data for a in ins])
    out_ptrs = (C.c_void_p * 18)(*[a.ctypes.data for a in outs])
    zz = np.ascontiguousarray(z, np.float32) if z is not None and len(z) else np.zeros((1, 3), np.float32)
    n = L.orc_densify_and_prune(C.c_int(P), in_ptrs, _p(accum), _p(denom), _p(max_radii), C.c_float(max_grad), C.c_float(min_opacity),
                                C.c_float(extent), C.c_float(percent_dense), C.c_int(int(max_screen_size)), _p(zz), out_ptrs)
    shapes = [(n, 3), (n, 1, 3), (n, 15, 3), (n, 1), (n, 3), (n, 4)]
    res = [o[:n * DN_K[i % 6]].reshape(shapes[i % 6]).copy() for i, o in enumerate(outs)]
    return res[0:6], res[6:12], res[12:18]


def reset_opacity(opacity):
    L = lib()
    o = np.ascontiguousarray(opacity, np.float32).copy()
    L.orc_reset_opacity(C.c_int(o.size), _p(o))
    return o
