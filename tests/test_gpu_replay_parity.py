"""GPU parity on the reference's bundled logs through the fused window kernel (fbus_step_windows).
BASELINE configs[0] (land replay) and configs[1] (water: refractive solve feeding the update).
Tolerances (north_star): per-frame state / covariance 1e-9 relative; whole-trajectory position divergence <= 1e-6 m."""
import numpy as np
import pytest

from helpers import cov_close

pytestmark = pytest.mark.gpu


def _oracle_replay(cfg, imu, image_rows, n_init=500, use_iir=False, cls=None):
    import orc
    from fbus_ekf_b200 import capi, replay
    if use_iir:
        import fbus_oracle_np
        imu = fbus_oracle_np.iir_prefilter(imu, restart_at=(n_init,))
    o = (cls or orc.Oracle)(cfg, 1)
    t_imu = np.ascontiguousarray(imu[:, 0])
    data = np.ascontiguousarray(imu[:, 1:7, None])
    stream = capi.make_imu_stream(t_imu, data, 1)
    o.init_gravity_gyrobias(stream, 0, n_init)
    t_frames, groups = replay.group_frames(image_rows)
    ids, pose = replay.frames_to_soa(t_frames, groups, 1)
    det = capi.make_det_frames(t_frames, ids, pose, 1, ids.shape[1])
    off = replay.window_offsets(t_imu, t_frames, n_init)
    trace = np.zeros((len(t_frames), 17, 1))
    o.step_windows(stream, det, off, 0, len(t_frames), trace)
    return trace[:, :, 0], o.get_state()


@pytest.mark.parametrize("use_iir", [False, True])
@pytest.mark.parametrize("name", ["land", "water"])
def test_log_replay(cfg, golden, name, use_iir):
    from fbus_ekf_b200 import replay
    imu, img = golden[f"{name}_imu"], golden[f"{name}_image"]
    ref_rows, ref_state = _oracle_replay(cfg, imu, img, use_iir=use_iir)
    out = replay.replay_log(imu, img, cfg, use_iir=use_iir, chunk=400)
    rows, st = out["rows"], out["state"]
    assert np.isfinite(rows).all()
    assert np.array_equal(rows[:, 0], ref_rows[:, 0])                      # timestamps
    div = np.abs(rows[:, 1:4] - ref_rows[:, 1:4]).max()
    assert div <= 1e-6, f"trajectory position divergence {div} m"
    assert np.abs(rows[:, 4:8] - ref_rows[:, 4:8]).max() <= 1e-9           # quaternion per frame
    assert np.abs(rows[:, 8:17] - ref_rows[:, 8:17]).max() <= 1e-9         # v, b_a, b_g per frame
    ok, w = cov_close(st["P"], ref_state["P"], 1e-9)
    assert ok, f"final covariance off by {w}"
    assert int(st["status"][0]) == int(ref_state["status"][0])
    assert int(st["status"][0]) & 0x4                                      # both logs contain vision gaps -> resets


@pytest.mark.parametrize("name", ["land", "water"])
def test_log_replay_against_the_reference_itself(cfg, golden, name):
    """the same replay against oracle/_ref -- the reference's own filter.cpp, compiled unmodified and executed (prebuilt
    library travelling with the snapshot) -- at north_star's tolerances, without the oracle in between"""
    import orc
    from fbus_ekf_b200 import replay
    if not orc.ref_available():
        pytest.skip("oracle/_ref did not travel")
    imu, img = golden[f"{name}_imu"], golden[f"{name}_image"]
    ref_rows, ref_state = _oracle_replay(cfg, imu, img, cls=orc.Ref)
    out = replay.replay_log(imu, img, cfg, chunk=400)
    rows, st = out["rows"], out["state"]
    assert np.array_equal(rows[:, 0], ref_rows[:, 0])
    assert np.abs(rows[:, 1:4] - ref_rows[:, 1:4]).max() <= 1e-6
    assert np.abs(rows[:, 4:17] - ref_rows[:, 4:17]).max() <= 1e-9
    ok, w = cov_close(st["P"], ref_state["P"], 1e-9)
    assert ok, f"final covariance off by {w}"
    for k in ("t", "q", "R", "p", "v", "ba", "bg", "g", "pv", "qv"):
        assert np.abs(st[k] - ref_state[k]).max() <= 1e-9, k


def test_water_refraction_feeds_update(cfg, golden):
    """configs[1]: corners.txt -> refractive solve on the GPU -> detections -> EKF on the GPU, against the same chain
    through the oracle; and the solved poses against the logged image.txt (6 significant digits)."""
    import orc
    from fbus_ekf_b200 import BatchFilter, replay
    wc, wi, imu = golden["water_corners"], golden["water_image"], golden["water_imu"]
    corners = np.ascontiguousarray(wc[:, 2:18].T.astype(np.float32))
    f = BatchFilter(cfg, batch=1)
    pose, c3, valid = f.RefractSolve(corners)
    pose_o, c3_o, valid_o = orc.refract_solve(cfg, corners)
    assert np.array_equal(valid, valid_o) and valid.all()
    assert np.abs(pose[:3] - pose_o[:3]).max() <= 1e-8 and np.abs(pose[3:] - pose_o[3:]).max() <= 1e-8
    assert np.abs(pose[:3].T - wi[:, 2:5]).max() <= 2e-5 and np.abs(pose[3:].T - wi[:, 5:9]).max() <= 5e-5
    img_gpu = np.concatenate([wc[:, 0:2], pose.T], axis=1)
    img_orc = np.concatenate([wc[:, 0:2], pose_o.T], axis=1)
    ref_rows, _ = _oracle_replay(cfg, imu, img_orc)
    out = replay.replay_log(imu, img_gpu, cfg)
    assert np.abs(out["rows"][:, 1:4] - ref_rows[:, 1:4]).max() <= 1e-6
