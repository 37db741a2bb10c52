"""FBUS_FLAG_MATLAB: the GPU path with the numerics of matlab/*.m against oracle/fbus_oracle_matlab.py, the NumPy
restatement of the .m files, driven like matlab/FBUS_EKF.m:114-210 -- per frame state 1e-9, covariance 1e-9 relative,
vision-only poses, a reset gap (ResetState.m: propagation and update skipped), and the un-fused calls."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _gpu_run(imu, img, n_frames, batch=3, n_init=500):
    from fbus_ekf_b200 import BatchFilter, capi, replay
    cfg = capi.config_matlab()
    f = BatchFilter(cfg, batch=batch)
    t_imu = np.ascontiguousarray(imu[:, 0])
    data = np.ascontiguousarray(np.repeat(imu[:, 1:7, None], batch, axis=2))
    stream = capi.make_imu_stream(t_imu, data, batch)
    f.InitGravityAndGyrobias(stream, 0, n_init)
    t_frames, groups = replay.group_frames(img)
    ids, pose = replay.frames_to_soa(t_frames, groups, batch)
    det = capi.make_det_frames(t_frames, ids, pose, batch, ids.shape[1])
    off = replay.window_offsets(t_imu, t_frames, n_init)
    # two calls: the image time of the previous frame (preImgTime) and the IMU cursor carry over on the device
    half = n_frames // 2
    tr = np.concatenate([f.StepWindows(stream, det, off, 0, half, trace=True), f.StepWindows(stream, det, off, half, n_frames, trace=True)])
    return tr, f.GetState(), f


def _check(tr, st, ref, n_frames):
    rows = ref["rows"]
    for b in range(tr.shape[2]):
        got = tr[:n_frames, 1:, b]
        assert np.abs(got - rows[:n_frames, 1:]).max() <= 1e-9, np.abs(got - rows[:n_frames, 1:]).max()
    P = st["P"][:, 0].reshape(18, 18)
    Pref = ref["P"][n_frames - 1]
    assert np.abs(P - Pref).max() <= 1e-9 * np.abs(Pref).max()
    big = np.abs(Pref) >= 1e-6 * np.abs(Pref).max()
    assert (np.abs(P - Pref)[big] <= 1e-9 * np.abs(Pref)[big]).all()


@pytest.mark.parametrize("name", ["land", "water"])
def test_log_replay(golden, name):
    import fbus_oracle_matlab as om
    n = 160
    imu, img = golden[f"{name}_imu"], golden[f"{name}_image"][:n + 1]
    ref = om.run_script(om.default_config(), imu, img, trace_cov=True)
    assert len(ref["rows"]) == n and set(ref["kinds"]) == {"update"}
    tr, st, f = _gpu_run(imu, img, n)
    _check(tr, st, ref, n)
    # vision-only pose of the last frame (ComputeVisionOnlyResults.m)
    assert np.abs(st["pv"][:, 0] - ref["vision"][n - 1, 0:3]).max() <= 1e-12
    assert np.abs(st["qv"][:, 0] - ref["vision"][n - 1, 3:7]).max() <= 1e-12
    assert (st["status"] == 0).all()


def test_reset_gap(golden):
    """0.3 s without image rows: ResetState.m replaces p, q, rotateMat, zeroes v and b_a (b_g kept), and that frame neither
    propagates nor updates (FBUS_EKF.m:168-171); the next frame skips the IMU samples older than the reset frame"""
    import fbus_oracle_matlab as om
    from fbus_ekf_b200 import capi
    imu, img = golden["land_imu"], golden["land_image"]
    t0 = img[60, 0]
    img = np.concatenate([img[:61], img[(img[:, 0] > t0 + 0.3)][:80]])
    n = len(img) - 1
    ref = om.run_script(om.default_config(), imu, img, trace_cov=True)
    assert ref["kinds"].count("reset") == 1 and ref["kinds"][61] == "reset"
    tr, st, f = _gpu_run(imu, img, n)
    _check(tr, st, ref, n)
    assert (st["status"] & capi.ST_RESET_DONE).all()
    # the gyro bias survives the reset (ResetState.m zeroes velocity and accelBias only)
    assert np.abs(tr[61, 14:17, 0]).max() > 0 and np.abs(tr[61, 8:14, 0]).max() == 0


def test_unfused_calls(golden):
    """InitPositionAndQuaternion / ImuUpdate / MeasureUpdate / ResetState as separate C-ABI calls, the MATLAB function API"""
    import fbus_oracle_matlab as om
    from fbus_ekf_b200 import BatchFilter, capi, replay
    imu, img = golden["land_imu"], golden["land_image"][:12]
    cfg_o = om.default_config()
    S = om.State(cfg_o)
    S.gravity, S.gyroBias = om.InitGravityAndGyrobias(imu[:500])
    om.InitPositionAndQuaternion(S, img[0:1, 1:9], cfg_o)
    f = BatchFilter(capi.config_matlab(), batch=2)
    t_imu = np.ascontiguousarray(imu[:, 0])
    data = np.ascontiguousarray(np.repeat(imu[:, 1:7, None], 2, axis=2))
    stream = capi.make_imu_stream(t_imu, data, 2)
    f.InitGravityAndGyrobias(stream, 0, 500)
    t_frames, groups = replay.group_frames(img)
    ids, pose = replay.frames_to_soa(t_frames, groups, 2)
    det = capi.make_det_frames(t_frames, ids, pose, 2, 1)
    f.InitPositionAndQuaternion(det, 0, 1)
    st = f.GetState()
    assert np.abs(st["q"][:, 0] - S.quaternion).max() <= 1e-13 and np.abs(st["p"][:, 0] - S.position).max() <= 1e-13
    assert np.abs(st["R"][:, 0] - S.rotateMat.ravel()).max() <= 1e-13
    # a few IMU samples with explicit time steps, an update, a reset
    j0 = int(np.searchsorted(t_imu, t_frames[0], side="right"))
    st["t"][:] = t_imu[j0 - 1]
    f.SetState(st)
    for j in range(j0, j0 + 6):
        om.ImuUpdate(S, imu[j, 1:4], imu[j, 4:7], imu[j, 0] - imu[j - 1, 0])
    f.ImuUpdate(stream, j0, 6, float("inf"))
    om.MeasureUpdate(S, img[1:2, 1:9], cfg_o)
    f.MeasureUpdate(det, 1)
    st = f.GetState()
    for got, want in ((st["q"], S.quaternion), (st["p"], S.position), (st["v"], S.velocity), (st["ba"], S.accelBias),
                      (st["bg"], S.gyroBias), (st["g"], S.gravity), (st["R"], S.rotateMat.ravel())):
        assert np.abs(got[:, 0] - want).max() <= 1e-10
    P = st["P"][:, 0].reshape(18, 18)
    assert np.abs(P - S.covariance).max() <= 1e-9 * np.abs(S.covariance).max()
    om.ResetState(S, img[2:3, 1:9], cfg_o)
    f.ResetState(det, 2)
    st = f.GetState()
    assert np.abs(st["q"][:, 0] - S.quaternion).max() <= 1e-13 and np.abs(st["p"][:, 0] - S.position).max() <= 1e-13
    assert np.abs(st["v"]).max() == 0 and np.abs(st["ba"]).max() == 0 and np.abs(st["bg"][:, 0] - S.gyroBias).max() <= 1e-13


def test_flag_combination_refused():
    from fbus_ekf_b200 import BatchFilter, FbusError, capi
    cfg = capi.config_matlab()
    cfg.flags |= capi.FBUS_FLAG_JOSEPH
    with pytest.raises(FbusError):
        BatchFilter(cfg, batch=1)
