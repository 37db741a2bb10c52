"""N3: in-air stereo triangulation VISION::NormalTriangulation (vision.cpp:395-466) + ComputeMarkerPose.
The reference's land log holds 3-D corners only (no 2-D land corners were logged), so there is no golden vector for the
DLT itself: three independent restatements (C++ one-sided Jacobi SVD, NumPy/LAPACK SVD, the device's Jacobi on A^T A)
are cross-checked, plus a synthetic pinhole round trip."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _inputs(cfg, n, seed, noise=2e-4, far=0.1):
    from fbus_ekf_b200 import synth
    rng = np.random.default_rng(seed)
    Rm, p = synth.random_marker_poses(n, rng, far_fraction=far)
    return Rm, p, synth.marker_corners_inair(cfg, Rm, p, noise=noise, rng=rng)


def test_oracles_and_host_math_agree(cfg, built):
    import fbus_oracle_np as onp
    import orc
    from fbus_ekf_b200 import capi
    hm = C.CDLL(os.path.join(ROOT, "tests", "_build_host_math.so"))
    Rm, p, c = _inputs(cfg, 150, 0)
    pose, c3, valid = orc.inair_solve(cfg, c)
    assert 0 < valid.sum() < len(valid)
    ocfg = onp.default_config()
    for i in range(c.shape[1]):
        Cn, ok = onp.normal_triangulation(ocfg, c[:, i])
        assert ok == bool(valid[i])
        c16, o3, po = np.ascontiguousarray(c[:, i]), np.zeros(12), np.zeros(7)
        assert hm.hm_inair(C.byref(cfg), c16.ctypes.data_as(capi.c_float_p), capi.dptr(o3), capi.dptr(po)) == valid[i]
        if ok:
            assert np.abs(Cn.ravel() - c3[:, i]).max() <= 1e-11
            assert np.abs(o3 - c3[:, i]).max() <= 1e-11 and np.abs(po - pose[:, i]).max() <= 1e-10
    # noise-free pinhole projections triangulate back to the truth (float32 corner rounding only)
    Rm, p, c0 = _inputs(cfg, 100, 1, noise=0.0, far=0.0)
    pose0, _, v0 = orc.inair_solve(cfg, c0)
    assert v0.all() and np.abs(pose0[:3].T - p).max() < 2e-6


@pytest.mark.gpu
def test_gpu_inair_solve(cfg):
    import orc
    from fbus_ekf_b200 import BatchFilter
    Rm, p, c = _inputs(cfg, 3000 + 11, 2)
    f = BatchFilter(cfg, batch=1)
    pose, c3, valid = f.InAirSolve(c)
    po, co, vo = orc.inair_solve(cfg, c)
    assert np.array_equal(valid, vo) and 0 < valid.sum() < len(valid)
    ok = valid == 1
    assert np.abs(c3[:, ok] - co[:, ok]).max() <= 1e-8
    assert np.abs(pose[:, ok] - po[:, ok]).max() <= 1e-8
    assert np.array_equal(pose[:, ~ok], po[:, ~ok])
