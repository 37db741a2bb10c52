"""Pins the oracle (oracle/fbus_oracle.cpp) to the REFERENCE'S OWN filter.cpp, executed here: oracle/_ref is
/root/reference/C++/src/filter.cpp compiled unmodified against stand-in third-party headers (oracle/ref_build/).
Every function of the EKF chain is compared per call on random states and over both bundled log
replays, reset frames included.  Tolerance 1e-13 relative: the two differ only in the association order of a few sums.

/root/reference does not exist on the GPU box; these tests use the prebuilt oracle/_ref/*.so that travels with the
snapshot and skip when neither the library nor the reference tree is there."""
import os

import numpy as np
import pytest

from helpers import random_states

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RTOL = 1e-13


def _need_ref():
    import orc
    if not orc.ref_available():
        if not os.path.exists("/root/reference/C++/src/filter.cpp"):
            pytest.skip("oracle/_ref not built and /root/reference absent")
        orc.build_ref()
    return orc


def _rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b) / np.maximum(np.abs(b), 1.0)).max())


def _cov_rel(P, Pref):
    scale = np.abs(Pref).max(axis=0, keepdims=True)
    return float((np.abs(P - Pref) / scale).max())


def _assert_same(so, sr, fields=("t", "q", "R", "p", "v", "ba", "bg", "g"), tol=RTOL, cov=True, what=""):
    for f in fields:
        assert _rel(so[f], sr[f]) <= tol, (what, f, _rel(so[f], sr[f]))
    if cov:
        assert _cov_rel(so["P"], sr["P"]) <= tol, (what, "P", _cov_rel(so["P"], sr["P"]))
    assert np.array_equal(so["prev_marker_id"], sr["prev_marker_id"]), what
    assert np.array_equal(so["initialised"], sr["initialised"]), what


def _pair(cfg, B):
    orc = _need_ref()
    return orc.Oracle(cfg, B), orc.Ref(cfg, B)


def test_ctor_state(cfg):
    """FILTER::FILTER (filter.hpp:63-137): P0, q = identity, everything else zero, preUsedMarkerID_ = 0"""
    o, r = _pair(cfg, 3)
    so, sr = o.get_state(), r.get_state()
    _assert_same(so, sr, tol=0.0)
    assert np.array_equal(so["P"], sr["P"])


def test_propagate_each_step(cfg):
    """F1 UpdateCovariance + F2 UpdateNominalState through F3 BatchImuProcessing, one sample per call, incl. the
    small-gyro branch (filter.cpp:544-561), samples before the nominal time (skipped) and after t_end (left alone)"""
    from fbus_ekf_b200 import capi
    B, N = 16, 24
    rng = np.random.default_rng(11)
    o, r = _pair(cfg, B)
    st = random_states(B, rng, t0=1.0)
    o.set_state(st)
    r.set_state(st)
    t = 1.0 + 0.005 * np.arange(-2, N - 2)  # two samples older than the state
    data = np.zeros((N, 6, B))
    data[:, 0:3] = rng.normal(size=(N, 3, B)) * 0.5 + np.array([0.0, 9.8, -0.1])[None, :, None]
    data[:, 3:6] = rng.normal(size=(N, 3, B)) * 0.05
    data[5, 3:6, : B // 2] = st["bg"][:, : B // 2] + 1e-6  # |w| < 1e-4: first-order branch
    imu = capi.make_imu_stream(t, data, B)
    for i in range(N):
        t_end = t[-3]  # the last two samples are later than the frame
        o.propagate(imu, i, 1, t_end)
        r.propagate(imu, i, 1, t_end)
        _assert_same(o.get_state(), r.get_state(), what=f"sample {i}")
    assert np.all(o.get_state()["t"] == t[-3])


def test_update_reset_init_calls(cfg):
    """F4 ObservationUpdate (marker hysteresis, unknown ids, sign flip of the quaternion residual), F5 ResetSystemState
    (gap above / below 0.1 s, out-of-range marker) and F6 InitializePose (success, no IMU before, unknown marker)"""
    from fbus_ekf_b200 import capi
    import fbus_oracle_np as onp
    B, m = 24, 3
    rng = np.random.default_rng(5)
    o, r = _pair(cfg, B)
    st = random_states(B, rng, t0=2.0)
    st["prev_marker_id"][:] = rng.integers(0, 3, B)
    o.set_state(st)
    r.set_state(st)
    W = 4
    ids = rng.integers(0, 9, size=(W, m, B)).astype(np.int32)
    ids[:, 2, ::3] = -1          # ragged frames
    ids[1, :, 1] = 77            # unknown marker everywhere in one frame of one filter
    ids[2, 0, 2] = 77
    pose = np.zeros((W, m, 7, B))
    pose[:, :, 0:3] = rng.normal(size=(W, m, 3, B)) * 0.4 + np.array([0.0, 0.0, 0.9])[None, None, :, None]
    pose[3, :, 2, 5] = 7.0       # farther than marker_max_dist
    q = rng.normal(size=(W, m, 4, B))
    q /= np.linalg.norm(q, axis=2, keepdims=True)
    pose[:, :, 3:7] = q
    tdet = np.array([2.0, 2.04, 2.3, 2.34])
    det = capi.make_det_frames(tdet, ids, pose, B, m)
    for w in range(W):
        o.reset_state(det, w)
        r.reset_state(det, w)
        so, sr = o.get_state(), r.get_state()
        _assert_same(so, sr, what=f"reset {w}")
        o.update(det, w)
        r.update(det, w)
        so, sr = o.get_state(), r.get_state()
        _assert_same(so, sr, what=f"update {w}")
    # vision-only pose of the last frame (filter.cpp:454-459) wherever a reset computed it
    done = (so["status"] & capi.ST_RESET_SKIPPED) == 0
    assert _rel(so["pv"][:, done], sr["pv"][:, done]) <= RTOL and _rel(so["qv"][:, done], sr["qv"][:, done]) <= RTOL
    assert (so["status"] & capi.ST_RESET_DONE).any() and (so["status"] & capi.ST_UPDATE_SKIPPED).any()
    # InitializePose on fresh filters
    o, r = _pair(cfg, B)
    for n_before, w in ((0, 0), (7, 1), (7, 0)):
        o.init_position_quaternion(det, w, n_before)
        r.init_position_quaternion(det, w, n_before)
        so, sr = o.get_state(), r.get_state()
        _assert_same(so, sr, what=f"init {n_before} {w}")
        fail_o = (so["status"] & capi.ST_INIT_FAILED) != 0
        fail_r = (sr["status"] & capi.ST_INIT_FAILED) != 0
        assert np.array_equal(fail_o, fail_r)
    assert so["initialised"].any()


def test_init_gravity(cfg, golden):
    from fbus_ekf_b200 import capi
    imu = golden["land_imu"][:500]
    o, r = _pair(cfg, 1)
    stream = capi.make_imu_stream(np.ascontiguousarray(imu[:, 0]), np.ascontiguousarray(imu[:, 1:7, None]), 1)
    o.init_gravity_gyrobias(stream, 0, 500)
    r.init_gravity_gyrobias(stream, 0, 500)
    so, sr = o.get_state(), r.get_state()
    assert _rel(so["bg"], sr["bg"]) <= 1e-15 and _rel(so["g"], sr["g"]) <= 1e-15


def _replay(cls, cfg, imu, img, n_init=500):
    from fbus_ekf_b200 import capi, replay
    f = cls(cfg, 1)
    t_imu = np.ascontiguousarray(imu[:, 0])
    stream = capi.make_imu_stream(t_imu, np.ascontiguousarray(imu[:, 1:7, None]), 1)
    f.init_gravity_gyrobias(stream, 0, n_init)
    t_frames, groups = replay.group_frames(img)
    ids, pose = replay.frames_to_soa(t_frames, groups, 1)
    det = capi.make_det_frames(t_frames, ids, pose, 1, ids.shape[1])
    off = replay.window_offsets(t_imu, t_frames, n_init)
    trace = np.zeros((len(t_frames), 17, 1))
    f.step_windows(stream, det, off, 0, len(t_frames), trace)
    return trace[:, :, 0], f.get_state(), f


@pytest.mark.parametrize("name", ["land", "water"])
def test_full_log_replay(cfg, golden, name):
    """both bundled logs, every frame (1257 / 1062 frames incl. the frames where the reference resets), through the
    reference's own FilterThreadFunction body vs the oracle: whole fusion.txt trace and final covariance"""
    orc = _need_ref()
    imu, img = golden[f"{name}_imu"], golden[f"{name}_image"]
    rows_o, st_o, _ = _replay(orc.Oracle, cfg, imu, img)
    rows_r, st_r, fr = _replay(orc.Ref, cfg, imu, img)
    assert np.array_equal(rows_o[:, 0], rows_r[:, 0])
    assert np.isfinite(rows_r).all()
    assert (st_o["status"] & 0x4).all(), "the replay is expected to contain reset frames"
    assert np.abs(rows_o[:, 1:4] - rows_r[:, 1:4]).max() <= 1e-11, "position trace"
    assert _rel(rows_o, rows_r) <= 1e-10
    _assert_same(st_o, st_r, tol=1e-10, what=name)
    cam, vis = fr.poses(0)  # what the viewer reads (filter.cpp:71-82,128-139)
    assert np.abs(cam[:3, 3] - st_o["p"][:, 0]).max() <= 1e-11 and np.abs(cam[:3, :3].ravel() - st_o["R"][:, 0]).max() <= 1e-11
    assert np.abs(vis[:3, 3] - st_o["pv"][:, 0]).max() <= 1e-11


def test_replay_prefix_tight_and_assertions(cfg, golden):
    """the first 60 frames before rounding differences have been amplified by the filter dynamics: 1e-13; run through the
    build of oracle/_ref that keeps the stand-in headers' bounds / shape assertions"""
    orc = _need_ref()
    img = golden["land_image"][:60]
    imu = golden["land_imu"]
    imu = imu[imu[:, 0] <= img[-1, 0] + 0.01]
    rows_o, st_o, _ = _replay(orc.Oracle, cfg, imu, img)
    rows_r, st_r, _ = _replay(orc.RefDebug, cfg, imu, img)
    assert _rel(rows_o, rows_r) <= RTOL
    _assert_same(st_o, st_r, tol=1e-12, what="prefix")
    # the three builds of oracle/_ref (-O2, -O2 with assertions, -O3 -mavx2: bench.py's CPU baseline) are the same arithmetic
    rows_p, st_p, _ = _replay(orc.Ref, cfg, imu, img)
    assert np.array_equal(rows_p, rows_r) and np.array_equal(st_p["P"], st_r["P"])
    if orc.host_has_avx2() and os.path.exists(orc.REF_AVX2_LIB_PATH):
        rows_f, st_f, _ = _replay(orc.RefFast, cfg, imu, img)
        assert np.array_equal(rows_f, rows_r) and np.array_equal(st_f["P"], st_r["P"])


def test_init_frame_keeps_later_samples(cfg):
    """InitializePose erases only the samples not later than the frame (filter.cpp:299-305,390): with arrival-based windows
    that overshoot the frame time, the later samples stay buffered and are propagated by the next frame"""
    from fbus_ekf_b200 import capi
    B = 2
    rng = np.random.default_rng(2)
    o, r = _pair(cfg, B)
    N = 40
    t = 1.0 + 0.005 * np.arange(N)
    data = np.zeros((N, 6, B))
    data[:, 0:3] = rng.normal(size=(N, 3, B)) * 0.3 + np.array([0.0, 9.8, 0.0])[None, :, None]
    data[:, 3:6] = rng.normal(size=(N, 3, B)) * 0.02
    imu = capi.make_imu_stream(t, data, B)
    tdet = np.array([t[9] + 0.001, t[19] + 0.001, t[29] + 0.001])
    ids = np.zeros((3, 1, B), dtype=np.int32)
    pose = np.zeros((3, 1, 7, B))
    pose[:, 0, 2] = 0.8
    pose[:, 0, 3] = 1.0
    det = capi.make_det_frames(tdet, ids, pose, B, 1)
    off = np.array([0, 14, 24, 34], dtype=np.uint32)  # every window holds 4 samples later than its frame
    tr_o, tr_r = np.zeros((3, 17, B)), np.zeros((3, 17, B))
    o.step_windows(imu, det, off, 0, 3, tr_o)
    r.step_windows(imu, det, off, 0, 3, tr_r)
    assert _rel(tr_o, tr_r) <= RTOL
    _assert_same(o.get_state(), r.get_state(), what="overshoot")
    assert tr_r[1, 0, 0] == t[19]  # the frame after the init consumed samples 10..19, incl. the four the init frame left


def test_set_imu_data_iir_and_cap(cfg, golden):
    """FILTER::SetImuData (filter.cpp:24-55) itself: the 1-pole IIR and the 2000-sample cap (oldest 500 erased), against
    the closed forms the product's host side uses (replay.buffer_cap_keep) and a literal float64 recurrence"""
    orc = _need_ref()
    from fbus_ekf_b200 import capi, replay
    imu = golden["land_imu"][:2600]
    stream = capi.make_imu_stream(np.ascontiguousarray(imu[:, 0]), np.ascontiguousarray(imu[:, 1:7, None]), 1)
    r = orc.Ref(cfg, 1)
    for count in (1, 700, 2000, 2001, 2600):
        t, d = r.set_imu_data(stream, 0, count)
        f = np.empty((count, 6))
        f[0] = imu[0, 1:7]
        for i in range(1, count):
            f[i] = f[i - 1] * (1 - 0.1) + imu[i, 1:7] * 0.1
        keep = replay.buffer_cap_keep(imu[:count, 0], np.array([]), 0)
        assert len(t) == int(keep.sum()), count
        assert np.array_equal(t, imu[:count, 0][keep])
        assert np.array_equal(d, f[keep]), count
