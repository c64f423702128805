"""CPU: pins the C oracle (oracle/gs_oracle.c) against the committed outputs of the REFERENCE's own kernels
(tests/golden/ref_small.npz, generated on a B200 by tests/golden/make_golden.py).

Tolerances: integers exact except for a stated handful of float-threshold cases (the oracle header explains
why gcc/libm and nvcc/device-libm cannot be bit-identical); floats 1e-4 relative."""
import os

import numpy as np
import pytest

import oracle_c

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_small.npz")


def _case(z, name):
    g = {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "_")}
    cam = dict(viewmatrix=g["viewmatrix"], projmatrix=g["projmatrix"], campos=g["campos"], tanfovx=np.float32(g["tanfovx"]),
               tanfovy=np.float32(g["tanfovy"]), W=int(g["W"]), H=int(g["H"]))
    act = {k[3:]: g[k] for k in g if k.startswith("in_")}
    return g, cam, act


def _relbad(a, b, rtol=1e-4, atol=1e-7):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(np.mean(np.abs(a - b) > rtol * np.maximum(np.abs(a), np.abs(b)) + atol))


@pytest.mark.parametrize("name", ["a", "b"])
def test_oracle_matches_reference_golden(name):
    if not os.path.exists(GOLD):
        pytest.fail("tests/golden/ref_small.npz missing (generate it on the GPU box with tests/golden/make_golden.py)")
    z = np.load(GOLD)
    g, cam, act = _case(z, name)
    D = int(g["D"])
    f = oracle_c.forward(cam, act, D=D, bg=g["bg"])
    P = act["means3D"].shape[0]
    vis = g["radii"] > 0
    # --- integer / index outputs
    assert (f["radii"] != g["radii"]).sum() <= max(1, P // 1000), "radii"
    same = f["radii"] == g["radii"]
    assert (f["tiles_touched"][same] != g["tiles_touched"].astype(np.uint32)[same]).sum() <= max(1, P // 1000)
    if np.array_equal(f["radii"], g["radii"]) and np.array_equal(f["tiles_touched"], g["tiles_touched"].astype(np.uint32)):
        assert f["num_rendered"] == int(g["num_rendered"])
        assert np.array_equal(f["values"], g["values_sorted"].astype(np.uint32)), "sorted Gaussian ids"
        assert np.array_equal(f["ranges"], g["ranges"].astype(np.uint32)), "tile ranges"
        kdiff = (f["keys"] != g["keys_sorted"].astype(np.uint64)).mean()
        assert kdiff <= 0.02, f"sorted keys: {kdiff} differ (depth last-ulp)"
        nc = (f["n_contrib"] != g["n_contrib"].astype(np.uint32)).mean()
        assert nc <= 2e-3, f"n_contrib differs on {nc} of pixels"
    # --- float outputs
    for k in ("depths", "means2D", "conic_opacity", "rgb", "cov3D"):
        assert _relbad(f[k][vis & same], g[k][vis & same]) <= 1e-3, k
    assert _relbad(f["out_color"], g["out_color"], atol=1e-5) <= 2e-3, "out_color"
    # --- backward, driven by the reference's own forward state so the comparison isolates the backward math
    fwd_ref = dict(f)
    fwd_ref.update(radii=g["radii"], means2D=g["means2D"], conic_opacity=g["conic_opacity"], colors=g["rgb"], cov3D=g["cov3D"],
                   clamped=g["clamped"].astype(np.uint8), ranges=g["ranges"].astype(np.uint32), values=g["values_sorted"].astype(np.uint32),
                   final_T=g["final_T"], n_contrib=g["n_contrib"].astype(np.uint32))
    b = oracle_c.backward(cam, act, fwd_ref, g["dL_dpix"], D=D)
    for mine, ref in [("dL_dmean2D", "dL_dmeans2D"), ("dL_dconic", "dL_dconic"), ("dL_dopacity", "dL_dopacity"), ("dL_dcolor", "dL_dcolors"),
                      ("dL_dmean3D", "dL_dmeans3D"), ("dL_dcov3D", "dL_dcov3D"), ("dL_dsh", "dL_dsh"), ("dL_dscale", "dL_dscales"),
                      ("dL_drot", "dL_drotations")]:
        a, r = b[mine].reshape(P, -1), g[ref].reshape(P, -1)
        nrm = np.linalg.norm((a - r).astype(np.float64)) / (np.linalg.norm(r.astype(np.float64)) + 1e-30)
        assert nrm < 5e-5, f"{mine}: rel-norm {nrm}"
