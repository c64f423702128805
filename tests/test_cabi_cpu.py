"""CPU: the C-ABI library builds/loads without a GPU, exports every symbol include/psb200.h declares, its
size functions are pure, and argument validation fails loudly before any CUDA call."""
import ctypes as C

import numpy as np
import pytest

from photo_slam_b200 import _lib


def test_exports_every_declared_symbol():
    L = _lib.lib()
    names = _lib.exported_symbols()
    assert {"psb_rasterize_forward", "psb_rasterize_backward", "psb_mark_visible", "psb_version", "psb_last_error"} <= set(names)
    for n in names:
        assert hasattr(L, n), f"libpsb200.so does not export {n}"
    assert L.psb_version() >= 100


def test_scratch_sizes_are_pure_and_monotonic():
    L = _lib.lib()
    for fn in (L.psb_geometry_bytes, L.psb_binning_bytes, L.psb_image_bytes):
        a, b, c = fn(0), fn(1000), fn(100000)
        assert 0 < a <= b <= c and fn(1000) == b
    # per-Gaussian state: 48 B record + 8 B rect + 4 B tiles + 4 B tile mask + 2x(4+4) B sort ping-pong + 4 B offsets (+ sort status)
    per = (L.psb_geometry_bytes(2_000_000) - L.psb_geometry_bytes(1_000_000)) / 1e6
    assert 84 <= per <= 88, per


def test_argument_validation_before_any_cuda_call():
    L = _lib.lib()
    cb = _lib.ALLOC_FN(lambda n, u: 0)
    one = 4096  # fake non-null pointer; never dereferenced because validation fails first
    # both shs and colors_precomp given
    rc = L.psb_rasterize_forward(cb, None, cb, None, cb, None, 10, 3, 16, one, 64, 64, one, one, one, one, one, 1.0, one, None, one, one,
                                 one, 1.0, 1.0, 0, one, None, None)
    assert rc == -1 and b"exactly one of shs" in L.psb_last_error()
    # neither scales/rotations nor cov3D
    rc = L.psb_rasterize_forward(cb, None, cb, None, cb, None, 10, 3, 16, one, 64, 64, one, one, None, one, None, 1.0, None, None, one,
                                 one, one, 1.0, 1.0, 0, one, None, None)
    assert rc == -1 and b"scales+rotations" in L.psb_last_error()
    # SH degree larger than the coefficient count allows
    rc = L.psb_rasterize_forward(cb, None, cb, None, cb, None, 10, 3, 4, one, 64, 64, one, one, None, one, one, 1.0, one, None, one, one,
                                 one, 1.0, 1.0, 0, one, None, None)
    assert rc == -1
    rc = L.psb_rasterize_forward(cb, None, cb, None, cb, None, -1, 3, 16, one, 64, 64, one, one, None, one, one, 1.0, one, None, one,
                                 one, one, 1.0, 1.0, 0, one, None, None)
    assert rc == -1
    assert L.psb_mark_visible(5, None, None, None, None, None) == -1


def test_rasterizer_frontend_validation_is_host_side():
    import torch
    from photo_slam_b200 import rasterizer
    P = 4
    z = torch.zeros
    rs = rasterizer.GaussianRasterizationSettings(32, 32, 1.0, 1.0, z(3), 1.0, torch.eye(4).flatten(), torch.eye(4).flatten(), 3, z(3))
    r = rasterizer.GaussianRasterizer(rs)
    with pytest.raises(RuntimeError, match="SHs or precomputed colors"):
        r(z(P, 3), z(P, 3), z(P, 1), shs=None, colors_precomp=None, scales=z(P, 3), rotations=z(P, 4))
    with pytest.raises(RuntimeError, match="scale/rotation pair or precomputed 3D covariance"):
        r(z(P, 3), z(P, 3), z(P, 1), shs=z(P, 16, 3), scales=z(P, 3), rotations=z(P, 4), cov3D_precomp=z(P, 6))
    with pytest.raises(RuntimeError, match="num_points, 3"):
        rasterizer.RasterizeGaussiansCUDA(z(3), z(P, 2), None, None, None, None, 1.0, None, None, None, 1.0, 1.0, 8, 8, None, 0, None, False)


def test_libtorch_shim_exports_the_reference_symbols():
    """libcuda_rasterizer.so must export the exact Itanium-mangled names that reference include/rasterize_points.h:18-65
    and cuda_rasterizer/rasterizer.h:24-82 declare (what libgaussian_mapper.so links against). The names below were
    obtained by compiling a translation unit against the reference's own headers (`nm -u`)."""
    import os
    import subprocess
    lib = os.path.join(os.path.dirname(_lib.LIB_PATH), "libcuda_rasterizer.so")
    assert os.path.exists(lib), "build it with `python photo-slam_b200/build.py --torch`"
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True).stdout
    for sym in ("_Z11markVisibleRN2at6TensorES1_S1_",
                "_Z22RasterizeGaussiansCUDARKN2at6TensorES2_S2_S2_S2_S2_fS2_S2_S2_ffiiS2_iS2_b",
                "_Z30RasterizeGaussiansBackwardCUDARKN2at6TensorES2_S2_S2_S2_S2_fS2_S2_S2_ffS2_S2_iS2_S2_iS2_S2_",
                "_ZN14CudaRasterizer10Rasterizer11markVisibleEiPfS1_S1_Pb",
                "_ZN14CudaRasterizer10Rasterizer7forwardESt8functionIFPcmEES4_S4_iiiPKfiiS6_S6_S6_S6_S6_fS6_S6_S6_S6_S6_ffbPfPi",
                "_ZN14CudaRasterizer10Rasterizer8backwardEiiiiPKfiiS2_S2_S2_S2_fS2_S2_S2_S2_S2_ffPKiPcS5_S5_S2_PfS6_S6_S6_S6_S6_S6_S6_S6_",
                # include/operate_points.h:27-40, include/stereo_vision.h:26-40, third_party/simple-knn/spatial.h:14
                "_Z15transformPointsRN2at6TensorES1_",
                "_Z38scaleAndTransformThenMarkVisiblePointsRN2at6TensorES1_S1_S1_S1_S1_S1_Rif",
                "_Z21reprojectDepthPinholeRN2at6TensorES1_RSt6vectorIfSaIfEEi",
                "_Z66monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypointsRN2at6TensorES1_S1_S1_fRSt6vectorIfSaIfEEi",
                "_Z9distCUDA2RKN2at6TensorE"):
        assert sym in out, sym
