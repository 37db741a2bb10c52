"""Host logic: the structured device math (fbus_math.cuh / fbus_refract.cuh compiled for the host by
tests/host_math_harness.cpp) against the dense oracle on random states.  This checks the algebra the CUDA kernels use
(block-sparse F P F^T, reduced 6-dim update, two-sweep rank-6 downdate, closed-form eigenvector) on machines without
a GPU; the kernels themselves are checked on the GPU by the -m gpu tests."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import random_states

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm(built):
    from fbus_ekf_b200 import capi
    lib = C.CDLL(os.path.join(ROOT, "tests", "_build_host_math.so"))
    P = C.POINTER(capi.FbusConfig)
    lib.hm_propagate.argtypes = [P, capi.c_double_p, capi.c_double_p, capi.c_double_p, capi.c_double_p, C.c_double]
    lib.hm_update.argtypes = [P, capi.c_double_p, capi.c_double_p, C.c_int, capi.c_double_p, capi.c_double_p]
    lib.hm_refract.argtypes = [P, capi.c_float_p, capi.c_double_p, capi.c_double_p]
    lib.hm_marker_pose.argtypes = [P, capi.c_double_p, capi.c_double_p]
    return lib


def _pack(P):
    return np.array([P[i, j] for i in range(18) for j in range(i, 18)])


def _unpack(p):
    P = np.zeros((18, 18))
    k = 0
    for i in range(18):
        for j in range(i, 18):
            P[i, j] = P[j, i] = p[k]
            k += 1
    return P


def _nom(st, b=0):
    return np.concatenate([[st["t"][b]], st["q"][:, b], st["R"][:, b], st["p"][:, b], st["v"][:, b], st["ba"][:, b], st["bg"][:, b],
                           st["g"][:, b]]).copy()


@pytest.mark.parametrize("small_gyro", [False, True])
def test_propagate_and_update_vs_oracle(cfg, hm, small_gyro):
    import orc
    from fbus_ekf_b200 import capi
    rng = np.random.default_rng(11)
    worst = 0.0
    for trial in range(25):
        o = orc.Oracle(cfg, 1)
        st = random_states(1, rng)
        o.set_state(st)
        accel = rng.normal(size=3) + np.array([0, 9.8, -0.1])
        gyro = rng.normal(size=3) * (1e-6 if small_gyro else 0.05)
        dt = 0.005
        imu = capi.make_imu_stream(np.array([1.0 + dt]), np.concatenate([accel, gyro]).reshape(1, 6, 1).copy(), 1)
        o.propagate(imu, 0, 1, 10.0)
        so = o.get_state()
        nom, Pp = _nom(st), _pack(st["P"][:, 0].reshape(18, 18)).copy()
        assert hm.hm_propagate(C.byref(cfg), capi.dptr(Pp), capi.dptr(nom), capi.dptr(accel), capi.dptr(gyro), dt) == 0
        Pref = so["P"][:, 0].reshape(18, 18)
        worst = max(worst, np.abs(_unpack(Pp) - Pref).max() / np.abs(Pref).max(), np.abs(nom - _nom(so)).max())
        ids = np.zeros((1, 1, 1), dtype=np.int32)
        pose = np.zeros((1, 1, 7, 1))
        yP = rng.normal(size=3) * 0.2 + np.array([0, 0, 0.5])
        yQ = rng.normal(size=4)
        yQ /= np.linalg.norm(yQ)
        pose[0, 0, :3, 0], pose[0, 0, 3:, 0] = yP, yQ
        o.update(capi.make_det_frames(np.array([1.0 + dt]), ids, pose, 1, 1), 0)
        su = o.get_state()
        assert hm.hm_update(C.byref(cfg), capi.dptr(Pp), capi.dptr(nom), 0, capi.dptr(yP), capi.dptr(yQ)) == 0
        Pref = su["P"][:, 0].reshape(18, 18)
        worst = max(worst, np.abs(_unpack(Pp) - Pref).max() / np.abs(Pref).max(), np.abs(nom - _nom(su)).max())
    assert worst <= 1e-11, worst


def test_refraction_vs_oracle_and_logs(cfg, hm, golden):
    import orc
    from fbus_ekf_b200 import capi
    wc, wi = golden["water_corners"], golden["water_image"]
    sel = np.arange(0, len(wc), 5)
    corners = np.ascontiguousarray(wc[sel, 2:18].T.astype(np.float32))
    pose_o, c3_o, _ = orc.refract_solve(cfg, corners)
    for n, r in enumerate(sel):
        c16 = np.ascontiguousarray(corners[:, n])
        c3, po = np.zeros(12), np.zeros(7)
        assert hm.hm_refract(C.byref(cfg), c16.ctypes.data_as(capi.c_float_p), capi.dptr(c3), capi.dptr(po)) == 1
        assert np.abs(c3 - c3_o[:, n]).max() <= 1e-12 and np.abs(po - pose_o[:, n]).max() <= 1e-10
        assert np.abs(po[:3] - wi[r, 2:5]).max() <= 2e-5 and np.abs(po[3:] - wi[r, 5:9]).max() <= 5e-5
