"""BASELINE configs[3] as a parity case: per frame 8 board markers x 4 corners x stereo -> refractive solve (closed form or
+ Gauss-Newton) -> detection frames -> nearest-marker EKF update, all on the GPU through the C ABI
(fbus_solve_to_detections + fbus_step_windows), against the same chain through the oracle."""
import numpy as np
import pytest

from helpers import cov_close, state_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("gn_iters", [0, 4])
def test_board_pipeline(cfg, gn_iters):
    import torch
    import orc
    import fbus_oracle_np as onp
    from fbus_ekf_b200 import BatchFilter, capi, synth
    B, m = 48, 8
    bcfg = synth.board_config(cfg)
    traj = synth.truth_trajectory(bcfg, 0.6, standoff=1.0)
    W, N = len(traj["t_frames"]), len(traj["t_imu"])
    base, ids, p_all = synth.board_base_corners(bcfg, traj)
    assert np.linalg.norm(p_all, axis=1).max() < 2.0
    rng = np.random.default_rng(8)
    # per-filter corner noise; item index (frame*m + slot)*B + filter
    corners = (np.repeat(base, B, axis=1) + rng.normal(size=(16, W * m * B)) * 2e-4).astype(np.float32)
    mids = np.ascontiguousarray(np.repeat(ids[:, :, None], B, axis=2))
    mids[3, 5, ::4] = -1                                  # a few undetected markers
    corners = np.ascontiguousarray(corners)
    f = BatchFilter(bcfg, batch=B)
    det_id, det_pose = f.SolveToDetections(corners, mids, W, m, underwater=True, gn_iters=gn_iters)
    assert det_id.shape == (W, m, B) and (det_id[3, 5, ::4] == -1).all() and (det_id >= -1).all()
    # the solve against the oracle (closed form) / the NumPy GN restatement (spot check)
    po, _, vo = orc.refract_solve(bcfg, corners)
    po = po.reshape(7, W, m, B).transpose(1, 2, 0, 3)
    vo = vo.reshape(W, m, B).astype(bool) & (mids >= 0)
    assert np.array_equal(det_id >= 0, vo)
    if gn_iters == 0:
        sel = np.broadcast_to(vo[:, :, None, :], det_pose.shape)
        assert np.abs(det_pose[sel] - po[sel]).max() <= 1e-8
    else:
        k = onp.Consts(onp.Config(tsc_left=np.array(bcfg.tsc_left).reshape(4, 4), tsc_right=np.array(bcfg.tsc_right).reshape(4, 4)))
        for (w, s, b) in [(0, 0, 0), (5, 7, 11), (9, 3, 47)]:
            i = (w * m + s) * B + b
            pg, qg, _ = onp.refract_solve_gn(k, corners[:, i].astype(np.float64), gn_iters)
            assert np.abs(pg - det_pose[w, s, :3, b]).max() <= 1e-8 and np.abs(qg - det_pose[w, s, 3:, b]).max() <= 1e-8
        # GN moves the closed-form pose by less than the corner noise allows (millimetres), never wildly
        assert np.abs(det_pose[:, :, :3][np.broadcast_to(vo[:, :, None, :], (W, m, 3, B))] -
                      po[:, :, :3][np.broadcast_to(vo[:, :, None, :], (W, m, 3, B))]).max() < 0.02
    # EKF on the GPU detections vs the oracle EKF on the very same detections
    imu = np.ascontiguousarray(traj["base_imu"][:, :, None] + rng.normal(size=(N, 6, B)) * np.array([0.015] * 3 + [1e-3] * 3)[None, :, None])
    s = capi.make_imu_stream(traj["t_imu"], imu, B)
    d = capi.make_det_frames(traj["t_frames"], det_id, det_pose, B, m)
    f.StepWindows(s, d, traj["win_off"], 0, W)
    o = orc.Oracle(bcfg, B)
    o.step_windows(s, d, traj["win_off"], 0, W, None, 4)
    sg, so = f.GetState(), o.get_state()
    assert state_close(sg, so, 1e-9)[0] and cov_close(sg["P"], so["P"], 1e-9)[0]
    assert np.array_equal(sg["prev_marker_id"], so["prev_marker_id"]) and np.array_equal(sg["status"], so["status"])
    assert len(np.unique(sg["prev_marker_id"])) >= 1
    err = np.linalg.norm(sg["p"] - traj["truth_p"][-1][:, None], axis=0)
    assert err.max() < 0.03                               # the refraction front end + filter track the truth
