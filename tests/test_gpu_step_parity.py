"""GPU parity, per step: un-fused C-ABI calls (fbus_propagate / fbus_update / fbus_reset_state /
fbus_init_*) against the CPU oracle on identical seeded inputs.
Tolerance (north_star): per-step state and covariance within 1e-9 relative."""
import numpy as np
import pytest

from helpers import cov_close, random_states, state_close

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def _pair(cfg, batch):
    import orc
    from fbus_ekf_b200 import BatchFilter
    return BatchFilter(cfg, batch=batch), orc.Oracle(cfg, batch)


def _random_imu(rng, n, batch, t0, dt=0.005, small_gyro=False):
    from fbus_ekf_b200 import capi
    t = t0 + dt * (1 + np.arange(n))
    data = np.zeros((n, 6, batch))
    data[:, 0:3, :] = rng.normal(size=(n, 3, batch)) * 0.3 + np.array([0, 9.8, -0.1])[None, :, None]
    data[:, 3:6, :] = rng.normal(size=(n, 3, batch)) * (1e-6 if small_gyro else 0.05)
    return capi.make_imu_stream(t, np.ascontiguousarray(data), batch)


def _random_dets(rng, W, m, batch, t, ids_pool=(0, 1, 5, 16, 99), p_empty=0.2):
    from fbus_ekf_b200 import capi
    ids = rng.choice(ids_pool, size=(W, m, batch)).astype(np.int32)
    ids[rng.random(size=ids.shape) < p_empty] = -1
    pose = np.zeros((W, m, 7, batch))
    pose[:, :, 0:3, :] = rng.normal(size=(W, m, 3, batch)) * 0.4 + np.array([0, 0, 0.8])[None, None, :, None]
    q = rng.normal(size=(W, m, 4, batch))
    pose[:, :, 3:7, :] = q / np.linalg.norm(q, axis=2, keepdims=True)
    return capi.make_det_frames(np.asarray(t, dtype=np.float64), np.ascontiguousarray(ids), np.ascontiguousarray(pose), batch, m)


@pytest.mark.parametrize("small_gyro", [False, True])
def test_propagate_each_step(cfg, small_gyro):
    rng = np.random.default_rng(1)
    B = 96
    f, o = _pair(cfg, B)
    st = random_states(B, rng)
    f.SetState(st)
    o.set_state(st)
    imu = _random_imu(rng, 12, B, 1.0, small_gyro=small_gyro)
    for i in range(12):
        f.ImuUpdate(imu, i, 1, 10.0)
        o.propagate(imu, i, 1, 10.0)
        sg, so = f.GetState(), o.get_state()
        ok, w = state_close(sg, so, RTOL)
        assert ok, f"step {i}: nominal state off by {w}"
        ok, w = cov_close(sg["P"], so["P"], RTOL)
        assert ok, f"step {i}: covariance off by {w}"


def test_propagate_window_selection(cfg):
    """BatchImuProcessing skips samples before nominal.t and stops after t_end (filter.cpp:496-503)"""
    rng = np.random.default_rng(2)
    B = 40
    f, o = _pair(cfg, B)
    st = random_states(B, rng, t0=1.0123)
    f.SetState(st)
    o.set_state(st)
    imu = _random_imu(rng, 20, B, 1.0)  # samples at 1.005 .. 1.100; some < nominal.t
    f.ImuUpdate(imu, 0, 20, 1.0651)
    o.propagate(imu, 0, 20, 1.0651)
    sg, so = f.GetState(), o.get_state()
    assert np.array_equal(sg["t"], so["t"])
    assert state_close(sg, so, RTOL)[0]
    assert cov_close(sg["P"], so["P"], RTOL)[0]


def test_update_and_reset(cfg):
    rng = np.random.default_rng(3)
    B, m, W = 128, 3, 4
    f, o = _pair(cfg, B)
    st = random_states(B, rng)
    st["prev_marker_id"][:] = rng.choice([0, 1, 5], size=B)
    f.SetState(st)
    o.set_state(st)
    det = _random_dets(rng, W, m, B, 1.0 + 0.04 * np.arange(W) + 0.08)  # frame 2,3 exceed the 0.1 s gap
    for w in range(W):
        f.ResetState(det, w)
        o.reset_state(det, w)
        f.MeasureUpdate(det, w)
        o.update(det, w)
        sg, so = f.GetState(), o.get_state()
        ok, wv = state_close(sg, so, RTOL, fields=("t", "q", "R", "p", "v", "ba", "bg", "g", "pv", "qv"))
        assert ok, f"frame {w}: state off by {wv}"
        ok, wv = cov_close(sg["P"], so["P"], RTOL)
        assert ok, f"frame {w}: covariance off by {wv}"
        assert np.array_equal(sg["prev_marker_id"], so["prev_marker_id"])
        assert np.array_equal(sg["status"], so["status"])


def test_init_calls(cfg):
    rng = np.random.default_rng(4)
    B = 64
    f, o = _pair(cfg, B)
    imu = _random_imu(rng, 50, B, 0.0)
    f.InitGravityAndGyrobias(imu, 5, 40)
    o.init_gravity_gyrobias(imu, 5, 40)
    det = _random_dets(rng, 2, 2, B, [0.3, 0.34])
    for n_before in (0, 7):
        f.InitPositionAndQuaternion(det, 1, n_before)
        o.init_position_quaternion(det, 1, n_before)
        sg, so = f.GetState(), o.get_state()
        assert np.array_equal(sg["initialised"], so["initialised"])
        assert np.array_equal(sg["status"], so["status"])
        assert state_close(sg, so, 1e-12)[0]
    assert sg["initialised"].sum() > 0 and (sg["initialised"] == 0).sum() > 0  # both outcomes exercised


def test_known_answers(cfg):
    """analytic pins: dt = 0 => P' = P + diag(Qbar); symmetric in => symmetric out (packed storage)."""
    from fbus_ekf_b200 import BatchFilter, capi
    rng = np.random.default_rng(5)
    B = 8
    f = BatchFilter(cfg, batch=B)
    st = random_states(B, rng)
    f.SetState(st)
    t = np.array([1.0])
    data = np.ascontiguousarray(rng.normal(size=(1, 6, B)))
    f.ImuUpdate(capi.make_imu_stream(t, data, B), 0, 1, 10.0)
    sg = f.GetState()
    Q = np.zeros(18)
    Q[3:6], Q[6:9], Q[9:12], Q[12:15] = cfg.accel_n_cov, cfg.gyro_n_cov, cfg.accel_b_cov, cfg.gyro_b_cov
    P0 = st["P"].reshape(18, 18, B)
    P1 = sg["P"].reshape(18, 18, B)
    assert np.abs(P1 - (P0 + np.diag(Q)[:, :, None])).max() <= 1e-13 * np.abs(P0).max()
    assert np.array_equal(P1, P1.transpose(1, 0, 2))
    for k in ("p", "v", "q"):
        assert np.abs(sg[k] - st[k]).max() <= 1e-15


def test_joseph_flag(cfg):
    """FBUS_FLAG_JOSEPH (opt-in, north_star): Joseph-form covariance update.  Checked against the oracle's dense
    (I-KH)P(I-KH)^T + K R K^T and, being algebraically equal, against the default form at rounding level."""
    import copy
    import orc
    from fbus_ekf_b200 import BatchFilter, capi
    rng = np.random.default_rng(6)
    B = 64
    cj = copy.copy(cfg)
    cj.flags = capi.FLAG_JOSEPH
    fj, oj, fd = BatchFilter(cj, batch=B), orc.Oracle(cj, B), BatchFilter(cfg, batch=B)
    st = random_states(B, rng)
    for x in (fj, fd):
        x.SetState(st)
    oj.set_state(st)
    det = _random_dets(rng, 2, 2, B, [1.02, 1.06], ids_pool=(0, 1, 5), p_empty=0.0)
    for w in range(2):
        fj.MeasureUpdate(det, w)
        fd.MeasureUpdate(det, w)
        oj.update(det, w)
    sj, sd, so = fj.GetState(), fd.GetState(), oj.get_state()
    assert state_close(sj, so, RTOL)[0] and cov_close(sj["P"], so["P"], RTOL)[0]
    assert cov_close(sj["P"], sd["P"], 1e-9)[0]
    assert not np.array_equal(sj["P"], sd["P"])  # a different evaluation, not a no-op
