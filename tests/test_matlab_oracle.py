"""MATLAB-semantics mode, CPU side: the NumPy restatement of matlab/*.m (oracle/fbus_oracle_matlab.py) against the host build
of the product's math inlines for that mode (tests/host_math_harness.cpp), and its relation to the C++-semantics oracle where
the two filters coincide.  The GPU path is checked in tests/test_gpu_matlab_mode.py."""
import ctypes as C
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hm(built):
    from fbus_ekf_b200 import capi
    lib = C.CDLL(os.path.join(ROOT, "tests", "_build_host_math.so"))
    lib.hm_rotmat_to_quat_eig.argtypes = [capi.c_double_p, capi.c_double_p]
    lib.hm_expm_rot_minus_I.argtypes = [capi.c_double_p, C.c_double, capi.c_double_p]
    lib.hm_propagate_nominal_matlab.argtypes = [capi.c_double_p, capi.c_double_p, capi.c_double_p, C.c_double]
    return lib


def test_eig_quaternion(hm):
    """rotmat_to_quaternion.m: Jacobi on the host == LAPACK eigh, for exact rotations and for the calibration's
    slightly non-orthonormal R_IL; for an exact rotation it is the unit quaternion of that rotation"""
    import fbus_oracle_matlab as om
    import fbus_oracle_np as onp
    from fbus_ekf_b200 import capi
    rng = np.random.default_rng(0)
    mats = [np.diag([-1.0, -1.0, 1.0]) @ onp.TSC_LEFT_1[:3, :3]] + [r for _, r in onp.default_markers().values()]
    for _ in range(50):
        A = np.linalg.qr(rng.normal(size=(3, 3)))[0]
        mats.append(A * np.sign(np.linalg.det(A)))
    for R in mats:
        q = np.zeros(4)
        hm.hm_rotmat_to_quat_eig(capi.dptr(np.ascontiguousarray(R.ravel())), capi.dptr(q))
        qo = om.rotmat_to_quaternion(R)
        assert np.abs(q - qo).max() <= 1e-14
        assert abs(np.linalg.norm(q) - 1) <= 1e-15
        if np.abs(R @ R.T - np.eye(3)).max() < 1e-12:
            assert np.abs(om.quaternion_to_rotmat(q) - R).max() <= 1e-14


def test_expm_block(hm):
    """Fx(7:9,7:9) = expm(-[w]x dt) (ImuUpdate.m:68): closed form vs scipy's expm, also for tiny and zero rates"""
    from scipy.linalg import expm
    import fbus_oracle_matlab as om
    from fbus_ekf_b200 import capi
    rng = np.random.default_rng(1)
    for scale in (0.0, 1e-12, 1e-6, 1e-2, 1.0, 30.0):
        for _ in range(5):
            w = rng.normal(size=3) * scale
            W = np.zeros(9)
            hm.hm_expm_rot_minus_I(capi.dptr(w), 0.005, capi.dptr(W))
            ref = expm(-om.vector_to_crossmat(w) * 0.005) - np.eye(3)
            assert np.abs(W.reshape(3, 3) - ref).max() <= 2e-16 + 1e-15 * np.abs(ref).max()


def test_nominal_step(hm):
    """ImuUpdate.m:37-60,76-79 (nominal part) on random states with a stale carried rotation matrix"""
    import fbus_oracle_matlab as om
    from fbus_ekf_b200 import capi
    rng = np.random.default_rng(2)
    cfg = om.default_config()
    for _ in range(40):
        S = om.State(cfg)
        q = rng.normal(size=4)
        S.quaternion = q / np.linalg.norm(q)
        q2 = S.quaternion + 1e-3 * rng.normal(size=4)
        S.rotateMat = om.quaternion_to_rotmat(q2)  # not the rotation of the current quaternion, and not orthonormal
        S.position, S.velocity = rng.normal(size=3), rng.normal(size=3) * 0.3
        S.accelBias, S.gyroBias = rng.normal(size=3) * 0.05, rng.normal(size=3) * 2e-3
        S.gravity = np.array([9.8, 0, 0]) + rng.normal(size=3) * 0.01
        nom = np.concatenate([[1.0], S.quaternion, S.rotateMat.ravel(), S.position, S.velocity, S.accelBias, S.gyroBias, S.gravity])
        accel, gyro = rng.normal(size=3) + [0, 9.8, 0], rng.normal(size=3) * 0.1
        hm.hm_propagate_nominal_matlab(capi.dptr(nom), capi.dptr(accel), capi.dptr(gyro), 0.005)
        om.ImuUpdate(S, accel, gyro, 0.005)
        ref = np.concatenate([S.quaternion, S.rotateMat.ravel(), S.position, S.velocity])
        assert np.abs(nom[1:20] - ref).max() <= 1e-13


def test_script_runs_and_stays_near_the_cpp_filter(golden):
    """the two published implementations of the same filter on the same log: finite, and within a few cm of each other
    (they differ in P0, R, the zeroed quaternion residual ... SURVEY A.4) -- a sanity check of the restatement, not a pin"""
    import fbus_oracle_matlab as om
    import fbus_oracle_np as onp
    imu, img = golden["land_imu"], golden["land_image"][:200]
    r = om.run_script(om.default_config(), imu, img)
    assert np.isfinite(r["rows"]).all() and len(r["rows"]) == 199
    c = onp.replay(onp.default_config(), imu[imu[:, 0] <= img[-1, 0] + 0.01], img)
    assert np.abs(r["rows"][5:, 1:4] - c["rows"][5:199, 1:4]).max() < 0.05
