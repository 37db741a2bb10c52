"""Arrival-based IMU windows (every window holds samples later than its frame's timestamp, the live case): the init frame
erases only the samples not later than the frame (filter.cpp:299-305,390), the propagate path leaves later samples buffered
(filter.cpp:501-503,520).  GPU (both covariance stores, both CTA sizes) vs the oracle and -- where oracle/_ref travelled --
vs the reference's own filter.cpp."""
import numpy as np
import pytest

from helpers import cov_close

pytestmark = pytest.mark.gpu


def _case(B, rng):
    N = 64
    t = 1.0 + 0.005 * np.arange(N)
    data = np.zeros((N, 6, B))
    data[:, 0:3] = rng.normal(size=(N, 3, B)) * 0.3 + np.array([0.0, 9.8, 0.0])[None, :, None]
    data[:, 3:6] = rng.normal(size=(N, 3, B)) * 0.02
    W = 5
    tdet = np.array([t[9 + 10 * w] + 0.001 for w in range(W)])
    ids = np.zeros((W, 1, B), dtype=np.int32)
    ids[0, 0, 1::4] = -1   # these filters see their first detection (and initialise) one frame later, with a longer buffer
    ids[1, 0, 2::4] = 99   # unknown marker on the second frame: for the filters initialised in frame 0, update skipped
    pose = np.zeros((W, 1, 7, B))
    pose[:, 0, 0:3] = rng.normal(size=(W, 3, B)) * 0.05 + np.array([0.0, 0.0, 0.8])[None, :, None]
    q = rng.normal(size=(W, 4, B)) * 0.02 + np.array([1.0, 0, 0, 0])[None, :, None]
    pose[:, 0, 3:7] = q / np.linalg.norm(q, axis=1, keepdims=True)
    off = np.array([0] + [14 + 10 * w for w in range(W)], dtype=np.uint32)  # 4 samples later than each frame
    return t, data, tdet, ids, pose, off


@pytest.mark.parametrize("B", [48, 160])
def test_init_frame_overshoot(cfg, B, cov_store):
    import orc
    from fbus_ekf_b200 import BatchFilter, capi
    rng = np.random.default_rng(9)
    t, data, tdet, ids, pose, off = _case(B, rng)
    W = len(tdet)
    imu = capi.make_imu_stream(t, data, B)
    det = capi.make_det_frames(tdet, ids, pose, B, 1)
    f = BatchFilter(cfg, batch=B)
    bg = capi.make_imu_stream(t[:8], np.ascontiguousarray(data[:8]), B)
    f.InitGravityAndGyrobias(bg, 0, 8)
    tr_g = f.StepWindows(imu, det, off, 0, W, trace=True)
    sg = f.GetState()
    o = orc.Oracle(cfg, B)
    o.init_gravity_gyrobias(bg, 0, 8)
    tr_o = np.zeros((W, 17, B))
    o.step_windows(imu, det, off, 0, W, tr_o)
    so = o.get_state()
    assert np.array_equal(sg["status"], so["status"])
    assert np.array_equal(tr_g[:, 0], tr_o[:, 0])
    # the frame after the init consumed the four samples the init frame left buffered
    assert tr_g[1, 0, 0] == t[19]
    assert np.abs(tr_g - tr_o).max() <= 1e-9
    assert cov_close(sg["P"], so["P"], 1e-9)[0]
    if orc.ref_available():
        r = orc.Ref(cfg, B)
        r.init_gravity_gyrobias(bg, 0, 8)
        tr_r = np.zeros((W, 17, B))
        r.step_windows(imu, det, off, 0, W, tr_r)
        assert np.abs(tr_g - tr_r).max() <= 1e-9
        assert cov_close(sg["P"], r.get_state()["P"], 1e-9)[0]
