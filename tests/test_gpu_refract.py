"""GPU parity of the refractive flat-port marker-pose solve (fbus_refract_solve / fbus_marker_pose).
Tolerance (north_star): 1e-8 rad / 1e-8 m against the oracle; 2e-5 / 5e-5 against the reference's logs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_water_log(cfg, golden):
    import orc
    from fbus_ekf_b200 import BatchFilter
    wc, wi = golden["water_corners"], golden["water_image"]
    corners = np.ascontiguousarray(wc[:, 2:18].T.astype(np.float32))
    f = BatchFilter(cfg, batch=1)
    pose, c3, valid = f.RefractSolve(corners)
    po, co, vo = orc.refract_solve(cfg, corners)
    assert np.array_equal(valid, vo)
    assert np.abs(c3 - co).max() <= 1e-8
    assert np.abs(pose - po).max() <= 1e-8
    assert np.abs(pose[:3].T - wi[:, 2:5]).max() <= 2e-5
    assert np.abs(pose[3:].T - wi[:, 5:9]).max() <= 5e-5


def test_land_log_marker_pose(cfg, golden):
    import orc
    from fbus_ekf_b200 import BatchFilter
    lc, li = golden["land_corners"], golden["land_image"]
    c3 = np.ascontiguousarray(lc[:, 2:14].T)
    f = BatchFilter(cfg, batch=1)
    pose = f.MarkerPose(c3)
    po = orc.marker_pose(cfg, c3)
    assert np.abs(pose - po).max() <= 1e-8
    assert np.abs(pose[:3].T - li[:, 2:5]).max() <= 2e-5
    assert np.abs(pose[3:].T - li[:, 5:9]).max() <= 5e-5


def test_synthetic_and_edges(cfg):
    """noisy synthetic corners incl. far markers (rejected: valid = 0, later corners zeroed), n not a multiple of the CTA"""
    import orc
    from fbus_ekf_b200 import BatchFilter
    from fbus_ekf_b200 import synth
    rng = np.random.default_rng(7)
    n = 1000 + 37
    corners = synth.random_marker_corners(cfg, n, rng, far_fraction=0.1)
    f = BatchFilter(cfg, batch=1)
    pose, c3, valid = f.RefractSolve(corners)
    po, co, vo = orc.refract_solve(cfg, corners, n_threads=4)
    assert np.array_equal(valid, vo)
    assert 0 < valid.sum() < n
    assert np.abs(c3 - co).max() <= 1e-8
    ok = valid == 1
    assert np.abs(pose[:, ok] - po[:, ok]).max() <= 1e-8
    assert np.array_equal(pose[:, ~ok], po[:, ~ok])


def test_empty(cfg):
    from fbus_ekf_b200 import BatchFilter
    f = BatchFilter(cfg, batch=1)
    pose, c3, valid = f.RefractSolve(np.zeros((16, 0), dtype=np.float32))
    assert pose.shape == (7, 0)
