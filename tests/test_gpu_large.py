"""Full-size runs (BASELINE configs[3]/[4] sizes) checked through size-independent properties, plus an oracle spot
check on a random sample of filters."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_million_filters_properties(cfg):
    import torch
    import orc
    from fbus_ekf_b200 import BatchFilter, capi, synth
    B = 1 << 20
    traj = synth.truth_trajectory(cfg, 0.4)
    N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
    f = BatchFilter(cfg, batch=B)
    imu_d = torch.empty((N, 6, B), dtype=torch.float64, device="cuda")
    id_d = torch.empty((W, 1, B), dtype=torch.int32, device="cuda")
    pose_d = torch.empty((W, 1, 7, B), dtype=torch.float64, device="cuda")
    f.SynthStreams(synth.make_synth_spec(traj, seed=77), imu_d.data_ptr(), id_d.data_ptr(), pose_d.data_ptr())
    imu = capi.make_imu_stream(traj["t_imu"], imu_d.data_ptr(), B, capi.FBUS_MEM_DEVICE)
    det = capi.make_det_frames(traj["t_frames"], id_d.data_ptr(), pose_d.data_ptr(), B, 1, capi.FBUS_MEM_DEVICE)
    f.StepWindows(imu, det, traj["win_off"], 0, W)
    tp = torch.tensor(traj["truth_p"][-1], device="cuda").reshape(3, 1).expand(3, B).contiguous()
    tq = torch.tensor(traj["truth_q"][-1], device="cuda").reshape(4, 1).expand(4, B).contiguous()
    st = f.Stats(tp.data_ptr(), tq.data_ptr(), capi.FBUS_MEM_DEVICE)
    assert st[3] == B and st[4] == 0                      # every filter finite, pose covariance block positive definite
    assert np.sqrt(st[0] / B) < 5e-3 and st[5] < 0.03     # the ensemble tracks the truth
    sg = f.GetState(with_cov=False)
    assert sg["initialised"].all() and not (sg["status"] & capi.ST_NONFINITE).any()
    assert np.abs(np.linalg.norm(sg["q"], axis=0) - 1).max() < 1e-12
    # oracle spot check on a strided sample of the same device-generated streams
    sel = np.arange(0, B, B // 64)[:64]
    idx = torch.from_numpy(sel).cuda()
    s = capi.make_imu_stream(traj["t_imu"], np.ascontiguousarray(imu_d[:, :, idx].cpu().numpy()), 64)
    d = capi.make_det_frames(traj["t_frames"], np.ascontiguousarray(id_d[:, :, idx].cpu().numpy()),
                             np.ascontiguousarray(pose_d[:, :, :, idx].cpu().numpy()), 64, 1)
    o = orc.Oracle(cfg, 64)
    o.step_windows(s, d, traj["win_off"], 0, W, None, 8)
    so = o.get_state(with_cov=False)
    for k in ("q", "p", "v", "ba", "bg", "g"):
        assert np.abs(sg[k][:, sel] - so[k]).max() <= 1e-9, k


def test_strong_scaling_shard_with_covariance(cfg):
    """131,072 filters = the shard of BASELINE configs[4] on 8 GPUs (1,024 CTAs of the tensor-memory kernel, 6.9 waves):
    state AND covariance of filters sampled across the grid -- first / last lanes of CTAs, the partial last wave -- against the
    oracle and, where oracle/_ref travelled, against the reference's own filter.cpp"""
    import torch
    import orc
    from fbus_ekf_b200 import BatchFilter, capi, synth
    from helpers import cov_close
    B = 131072
    traj = synth.truth_trajectory(cfg, 1.0)
    N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
    f = BatchFilter(cfg, batch=B)
    imu_d = torch.empty((N, 6, B), dtype=torch.float64, device="cuda")
    id_d = torch.empty((W, 1, B), dtype=torch.int32, device="cuda")
    pose_d = torch.empty((W, 1, 7, B), dtype=torch.float64, device="cuda")
    f.SynthStreams(synth.make_synth_spec(traj, seed=78), imu_d.data_ptr(), id_d.data_ptr(), pose_d.data_ptr())
    imu = capi.make_imu_stream(traj["t_imu"], imu_d.data_ptr(), B, capi.FBUS_MEM_DEVICE)
    det = capi.make_det_frames(traj["t_frames"], id_d.data_ptr(), pose_d.data_ptr(), B, 1, capi.FBUS_MEM_DEVICE)
    f.StepWindows(imu, det, traj["win_off"], 0, W)
    sg = f.GetState(with_cov=True)
    rng = np.random.default_rng(5)
    sel = np.unique(np.concatenate([[0, 31, 32, 127, 128, B - 129, B - 128, B - 1], 128 * rng.integers(0, B // 128, 40) + rng.integers(0, 128, 40),
                                    np.arange(148 * 128 * 6, 148 * 128 * 6 + 128 * 8, 61)]))
    n = len(sel)
    idx = torch.from_numpy(sel).cuda()
    s = capi.make_imu_stream(traj["t_imu"], np.ascontiguousarray(imu_d[:, :, idx].cpu().numpy()), n)
    d = capi.make_det_frames(traj["t_frames"], np.ascontiguousarray(id_d[:, :, idx].cpu().numpy()),
                             np.ascontiguousarray(pose_d[:, :, :, idx].cpu().numpy()), n, 1)
    checkers = [orc.Oracle] + ([orc.Ref] if orc.ref_available() else [])
    for cls in checkers:
        o = cls(cfg, n)
        o.step_windows(s, d, traj["win_off"], 0, W, None, 8)
        so = o.get_state(with_cov=True)
        for k in ("t", "q", "R", "p", "v", "ba", "bg", "g"):
            assert np.abs(sg[k][..., sel] - so[k]).max() <= 1e-9, (cls.__name__, k)
        ok, worst = cov_close(sg["P"][:, sel], so["P"], 1e-9)
        assert ok, (cls.__name__, worst)


def test_half_million_refractive_solves(cfg):
    """configs[3]: 65,536 filters x 8 markers per frame"""
    import orc
    from fbus_ekf_b200 import BatchFilter, synth
    rng = np.random.default_rng(4)
    base = synth.random_marker_corners(cfg, 65536, rng, far_fraction=0.02)
    corners = np.ascontiguousarray(np.tile(base, (1, 8)))
    f = BatchFilter(cfg, batch=1)
    pose, c3, valid = f.RefractSolve(corners)
    n = corners.shape[1]
    assert valid.shape == (n,) and 0.9 < valid.mean() < 1.0
    ok = valid == 1
    assert np.isfinite(pose[:, ok]).all()
    assert np.abs(np.linalg.norm(pose[3:, ok], axis=0) - 1).max() < 1e-6      # R -> q of an orthonormal frame
    assert (np.linalg.norm(c3[0:3, ok], axis=0) <= 2.0 + 1e-9).all()           # range gate
    # tiles are identical inputs -> identical outputs (determinism across CTAs)
    assert np.array_equal(pose[:, :65536], pose[:, 65536:2 * 65536])
    sel = np.arange(0, 65536, 97)
    po, co, vo = orc.refract_solve(cfg, np.ascontiguousarray(corners[:, sel]))
    assert np.array_equal(valid[sel], vo)
    good = vo == 1
    assert np.abs(pose[:, sel][:, good] - po[:, good]).max() <= 1e-8
    # host arrays go through the chunked, pipelined path (column chunks, two streams); device arrays through one launch:
    # same kernel, same inputs -> identical bits (a ragged size so that the last chunk is partial)
    import torch
    m = n - 1000
    ch = np.ascontiguousarray(corners[:, :m])
    ph, c3h, vh = f.RefractSolve(ch)
    cd = torch.from_numpy(ch).cuda()
    pd = torch.empty((7, m), dtype=torch.float64, device="cuda")
    c3d = torch.empty((12, m), dtype=torch.float64, device="cuda")
    vd = torch.empty(m, dtype=torch.int32, device="cuda")
    f.RefractSolveDevice(cd.data_ptr(), m, pd.data_ptr(), c3d.data_ptr(), vd.data_ptr())
    f.Synchronize()
    assert np.array_equal(ph, pd.cpu().numpy(), equal_nan=True) and np.array_equal(c3h, c3d.cpu().numpy(), equal_nan=True)
    assert np.array_equal(vh, vd.cpu().numpy())
