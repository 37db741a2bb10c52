"""Cross-checks the two independent oracle restatements (C++ dense/LDLT/Jacobi vs NumPy/LAPACK) on the EKF chain, plus
analytic known-answer tests (the EKF has no golden output better than cm level, so the parity chain is
oracle-C++ <-> oracle-NumPy <-> GPU)."""
import numpy as np


def _cpp_replay(cfg, imu, img, n_init=500):
    import orc
    from fbus_ekf_b200 import capi, replay
    o = orc.Oracle(cfg, 1)
    t_imu = np.ascontiguousarray(imu[:, 0])
    stream = capi.make_imu_stream(t_imu, np.ascontiguousarray(imu[:, 1:7, None]), 1)
    o.init_gravity_gyrobias(stream, 0, n_init)
    t_frames, groups = replay.group_frames(img)
    ids, pose = replay.frames_to_soa(t_frames, groups, 1)
    det = capi.make_det_frames(t_frames, ids, pose, 1, ids.shape[1])
    off = replay.window_offsets(t_imu, t_frames, n_init)
    trace = np.zeros((len(t_frames), 17, 1))
    o.step_windows(stream, det, off, 0, len(t_frames), trace)
    return trace[:, :, 0], o.get_state()


def test_replay_cpp_vs_numpy(cfg, golden):
    import fbus_oracle_np as onp
    for name, nf in (("land", 150), ("water", 120)):
        img = golden[f"{name}_image"][:nf]
        imu = golden[f"{name}_imu"]
        imu = imu[imu[:, 0] <= img[-1, 0] + 0.01]
        rows_c, st_c = _cpp_replay(cfg, imu, img)
        res = onp.replay(onp.default_config(), imu, img, trace_cov=True)
        rows_n = res["rows"]
        assert np.array_equal(rows_c[:, 0], rows_n[:, 0])
        assert np.abs(rows_c - rows_n).max() <= 1e-10, name
        Pn = res["P"][-1]
        Pc = st_c["P"][:, 0].reshape(18, 18)
        assert np.abs(Pc - Pn).max() <= 1e-9 * np.abs(Pn).max()
        big = np.abs(Pn) >= 1e-6 * np.abs(Pn).max()
        assert (np.abs(Pc - Pn)[big] <= 1e-9 * np.abs(Pn)[big]).all()


def test_known_answers_numpy_oracle():
    import fbus_oracle_np as onp
    cfg = onp.default_config()
    k = onp.Consts(cfg)
    rng = np.random.default_rng(0)
    f = onp.Filter(k)
    f.initialised = True
    f.q = np.array([0.6, 0.1, -0.3, 0.73])
    f.q /= np.linalg.norm(f.q)
    f.R = onp.q2R(f.q)
    f.g = np.array([9.8, 0, 0])
    A = rng.normal(size=(18, 18))
    f.P = A @ A.T * 1e-3 + np.eye(18) * 1e-2
    P0 = f.P.copy()
    # dt = 0  =>  P' = P + diag(Qbar)   (process noise is NOT scaled by dt)
    f.update_covariance(0.0, rng.normal(size=3), rng.normal(size=3))
    assert np.abs(f.P - (P0 + np.diag(k.Qbar))).max() < 1e-15
    assert np.array_equal(f.P, f.P.T)
    # at rest: R a = -g  =>  v and p stay put
    f.v[:] = 0
    p0 = f.p.copy()
    f.update_nominal(0.005, f.R.T @ (-f.g) + f.ba, f.bg.copy())
    assert np.abs(f.v).max() < 1e-15 and np.abs(f.p - p0).max() < 1e-15
    # H is the derivative of the predicted measurement w.r.t. the error state (finite differences)
    d = np.concatenate([[0], [0.05, -0.1, 0.6], [0.5, -0.5, 0.5, 0.5]])
    hP, hQ, H = f.measurement_model(d)
    eps = 1e-6
    for j in range(3):
        g = onp.Filter(k)
        g.__dict__.update({kk: (v.copy() if isinstance(v, np.ndarray) else v) for kk, v in f.__dict__.items()})
        g.p[j] += eps
        hP2, _, _ = g.measurement_model(d)
        assert np.abs((hP2 - hP) / eps - H[0:3, j]).max() < 1e-6
    for j in range(3):
        g = onp.Filter(k)
        g.__dict__.update({kk: (v.copy() if isinstance(v, np.ndarray) else v) for kk, v in f.__dict__.items()})
        dth = np.zeros(3)
        dth[j] = eps
        g.q = onp.qmul(f.q, np.concatenate([[1.0], 0.5 * dth]))
        g.R = onp.q2R(g.q / np.linalg.norm(g.q))
        hP2, hQ2, _ = g.measurement_model(d)
        assert np.abs((hQ2 - hQ) / eps - H[3:7, 6 + j]).max() < 1e-5
        assert np.abs((hP2 - hP) / eps - H[0:3, 6 + j]).max() < 1e-5


def test_update_is_noop_direction_for_zero_gain_rows(cfg):
    """C++ oracle: unknown marker id -> update skipped entirely, state untouched (filter.cpp:671-673)"""
    import orc
    from fbus_ekf_b200 import capi
    from helpers import random_states
    rng = np.random.default_rng(3)
    o = orc.Oracle(cfg, 4)
    st = random_states(4, rng)
    o.set_state(st)
    ids = np.full((1, 1, 4), 77, dtype=np.int32)
    pose = np.zeros((1, 1, 7, 4))
    pose[0, 0, 2] = 0.5
    pose[0, 0, 3] = 1.0
    o.update(capi.make_det_frames(np.array([1.0]), ids, pose, 4, 1), 0)
    s2 = o.get_state()
    for kf in ("q", "p", "v", "ba", "bg", "g", "P"):
        assert np.array_equal(s2[kf], st[kf])
    assert (s2["status"] & capi.ST_UPDATE_SKIPPED).all()
