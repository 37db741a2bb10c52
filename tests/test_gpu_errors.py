"""Error behaviour of the C ABI (include/fbus_ekf.h): every call returns 0 or a negative FBUS_E_* code and leaves a message in
fbus_last_error(); per-filter soft conditions do not abort the batch (the reference logs a warning and carries on,
filter.cpp:343-358,432-447,672-673)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_bad_arguments_return_codes(cfg):
    from fbus_ekf_b200 import BatchFilter, capi
    from fbus_ekf_b200.filter import FbusError
    lib = capi.lib()
    # creation: zero batch, non-existent device
    h = C.c_void_p()
    assert lib.fbus_create(C.byref(h), C.byref(cfg), 0, 0) == capi.FBUS_E_BADARG
    assert b"batch" in lib.fbus_last_error(None).lower() or lib.fbus_last_error(None)
    with pytest.raises(FbusError):
        BatchFilter(cfg, batch=4, device=99)
    f = BatchFilter(cfg, batch=8)
    # streams whose batch does not match the handle
    t = np.arange(4, dtype=np.float64) * 0.005
    imu = capi.make_imu_stream(t, np.zeros((4, 6, 4)), 4)
    rc = lib.fbus_propagate(f._h, C.byref(imu), 0, 4, 1.0)
    assert rc == capi.FBUS_E_BADARG and lib.fbus_last_error(f._h)
    # out-of-range sample window
    imu8 = capi.make_imu_stream(t, np.zeros((4, 6, 8)), 8)
    assert lib.fbus_propagate(f._h, C.byref(imu8), 2, 9, 1.0) == capi.FBUS_E_BADARG
    # null pointers
    assert lib.fbus_refract_solve(f._h, None, 4, None, None, None, capi.FBUS_MEM_HOST) == capi.FBUS_E_BADARG
    assert lib.fbus_get_state(f._h, None) == capi.FBUS_E_BADARG
    # the handle is still usable after the failed calls
    assert lib.fbus_propagate(f._h, C.byref(imu8), 0, 4, 1.0) == 0
    f.Synchronize()
    f.close()


def test_soft_conditions_are_status_bits(cfg):
    """unknown marker id / marker out of range / no detection: the frame is skipped for that filter only, the status word says why"""
    from fbus_ekf_b200 import BatchFilter, capi, synth
    B = 4
    traj = synth.truth_trajectory(cfg, duration=0.2)
    N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
    imu = np.ascontiguousarray(np.repeat(traj["base_imu"][:, :, None], B, axis=2))
    ids = np.zeros((W, 1, B), dtype=np.int32)
    pose = np.ascontiguousarray(np.repeat(traj["base_pose"][:, None, :, None], B, axis=3))
    ids[:, 0, 1] = 77              # filter 1: a marker that is not in the map
    pose[:, 0, 0:3, 2] *= 50.0     # filter 2: marker far beyond marker_max_dist
    ids[:, 0, 3] = -1              # filter 3: nothing detected
    f = BatchFilter(cfg, batch=B)
    f.StepWindows(capi.make_imu_stream(traj["t_imu"], imu, B), capi.make_det_frames(traj["t_frames"], ids, pose, B, 1), traj["win_off"], 0, W)
    st = f.GetState(with_cov=False)
    assert st["initialised"][0] == 1 and st["status"][0] & capi.FBUS_ST_NONFINITE == 0
    assert st["initialised"][1] == 0 and st["status"][1] & capi.FBUS_ST_INIT_FAILED
    assert st["initialised"][2] == 0 and st["status"][2] & capi.FBUS_ST_INIT_FAILED
    assert st["initialised"][3] == 0 and st["status"][3] & capi.FBUS_ST_NO_DETECTION
    assert np.isfinite(st["p"][:, 0]).all()
    f.close()
