"""Host logic of the replay driver and the synthetic workload generators (no device work)."""
import numpy as np


def test_group_frames_and_windows():
    from fbus_ekf_b200 import replay
    img = np.array([[1.00, 0, 0, 0, 1, 1, 0, 0, 0], [1.00, 3, 0, 0, 2, 1, 0, 0, 0], [1.04, 0, 0, 0, 1, 1, 0, 0, 0], [1.30, 5, 0, 0, 1, 1, 0, 0, 0]])
    t, groups = replay.group_frames(img)
    assert list(t) == [1.00, 1.04, 1.30] and [len(g) for g in groups] == [2, 1, 1]
    ids, pose = replay.frames_to_soa(t, groups, batch=3)
    assert ids.shape == (3, 2, 3) and pose.shape == (3, 2, 7, 3)
    assert ids[0, 1, 0] == 3 and ids[1, 1, 2] == -1 and pose[0, 1, 2, 1] == 2.0
    t_imu = np.array([0.98, 0.99, 1.00, 1.01, 1.02, 1.04, 1.05, 1.2])
    off = replay.window_offsets(t_imu, t, start=1)
    # window w = samples not later than frame w (filter.cpp:501-503); never before the initialisation block
    assert list(off) == [1, 3, 6, 8]


def test_oracle_iir_is_the_reference_recurrence():
    """SetImuData: first sample raw, then 0.9*previous filtered + 0.1*raw (filter.cpp:36-48); the GPU version
    (fbus_iir_prefilter) is checked against this one in tests/test_gpu_sensor_f32.py"""
    import fbus_oracle_np
    rng = np.random.default_rng(0)
    imu = np.concatenate([np.arange(20)[:, None] * 0.001, rng.normal(size=(20, 6))], axis=1)
    out = fbus_oracle_np.iir_prefilter(imu, restart_at=(10,))
    assert np.array_equal(out[0], imu[0]) and np.array_equal(out[10], imu[10])
    assert np.allclose(out[3, 1:], 0.9 * out[2, 1:] + 0.1 * imu[3, 1:], atol=0, rtol=0)
    assert np.array_equal(out[:, 0], imu[:, 0])


def test_forward_projection_round_trip(cfg):
    """flat-port forward projection is the inverse of the oracle's refractive triangulation; on noise-free synthetic
    corners the closed-form solve returns the true pose (up to float32 corners and the reference's truncated pi)"""
    import orc
    from fbus_ekf_b200 import synth
    rng = np.random.default_rng(5)
    Rm, p = synth.random_marker_poses(300, rng)
    corners = synth.marker_corners_from_pose(cfg, Rm, p, noise=0.0)
    pose, c3, valid = orc.refract_solve(cfg, corners)
    assert valid.all()
    assert np.abs(pose[:3].T - p).max() < 2e-6  # float32 corner rounding ~6e-8 amplified by the stereo geometry
    far = synth.random_marker_corners(cfg, 50, rng, far_fraction=1.0)
    _, _, v2 = orc.refract_solve(cfg, far)
    assert not v2.any()  # beyond makrer_dect_dist_thres -> rejected


def test_periodic_truth_is_periodic(cfg):
    from fbus_ekf_b200 import synth
    a = synth.truth_trajectory(cfg, 1.0, periodic=True)
    b = synth.truth_trajectory(cfg, 2.0, periodic=False)
    assert a["base_imu"].shape == (200, 6) and a["base_pose"].shape == (25, 7) and list(a["win_off"][:3]) == [0, 8, 16]
    assert b["base_imu"].shape == (400, 6)
    # periodic: the sample half a step before t=0 equals the one before t=period -> first and last+1 coincide
    a2 = synth.truth_trajectory(cfg, 1.0, periodic=True)
    assert np.array_equal(a["base_imu"], a2["base_imu"])
    assert np.abs(a["base_imu"][:, 0:3]).max() < 14 and np.abs(a["base_imu"][:, 3:6]).max() < 0.5


def test_recorded_frames_match_the_reference(cfg, golden):
    """replay.recorded_frames (host logic of the replay CLI: which frames get a data/fusion.txt row) against the reference's
    own FilterThreadFunction body run frame by frame (oracle/_ref): a row is recorded iff the pose was initialised BEFORE the
    frame (filter.cpp:207-248).  Includes frames whose initialisation fails (marker out of range / unknown id)."""
    import orc
    from fbus_ekf_b200 import capi, replay
    if not orc.ref_available():
        import pytest
        pytest.skip("oracle/_ref not built")
    imu = golden["land_imu"][:4000]
    img = golden["land_image"].copy()
    img = img[img[:, 0] <= imu[-1, 0]][:40]
    img[0, 1] = 99        # unknown id on the first frame: initialisation fails
    img[1, 4] = 7.0       # marker farther than marker_max_dist on the second
    t_imu = np.ascontiguousarray(imu[:, 0])
    t_frames, groups = replay.group_frames(img)
    rec = replay.recorded_frames(t_imu, t_frames, groups, cfg, 500)
    assert not rec[:3].any() and rec[3:].all()
    stream = capi.make_imu_stream(t_imu, np.ascontiguousarray(imu[:, 1:7, None]), 1)
    ids, pose = replay.frames_to_soa(t_frames, groups, 1)
    det = capi.make_det_frames(t_frames, ids, pose, 1, ids.shape[1])
    off = replay.window_offsets(t_imu, t_frames, 500)
    r = orc.Ref(cfg, 1)
    r.init_gravity_gyrobias(stream, 0, 500)
    got = []
    for w in range(len(t_frames)):
        got.append(bool(r.get_state(with_cov=False)["initialised"][0]))
        r.step_windows(stream, det, off, w, w + 1)
    assert np.array_equal(rec, np.array(got))
