"""N4: fisheye undistortion of the detected corner pixels, cv::fisheye::undistortPoints(K, D) as
VISION::DetectArucoTag applies it (vision.cpp:203,253,318,369).  The oracle restates OpenCV 3.4.3's algorithm and is
pinned against the OpenCV that is importable here (cv2); the GPU kernel is checked against the oracle bit for bit
(both round to float32 like cv::Point2f)."""
import numpy as np
import pytest


def _pixels(n, seed):
    rng = np.random.default_rng(seed)
    px = np.zeros((16, n), dtype=np.float32)
    px[0::2] = rng.uniform(20, 620, size=(8, n))
    px[1::2] = rng.uniform(20, 380, size=(8, n))
    return np.ascontiguousarray(px)


def test_oracle_matches_opencv(cfg, built):
    cv2 = pytest.importorskip("cv2")
    import orc
    px = _pixels(400, 0)
    out = orc.undistort_fisheye(cfg, px)
    for cam in range(2):
        K = np.array([[cfg.cam_k[cam][0], 0, cfg.cam_k[cam][2]], [0, cfg.cam_k[cam][1], cfg.cam_k[cam][3]], [0, 0, 1.0]])
        D = np.array(list(cfg.cam_d[cam]), dtype=np.float64).reshape(4, 1)
        pts = np.stack([px[cam * 8:cam * 8 + 8:2].ravel(), px[cam * 8 + 1:cam * 8 + 8:2].ravel()], axis=1).reshape(-1, 1, 2)
        ref = cv2.fisheye.undistortPoints(pts.astype(np.float32), K, D).reshape(-1, 2)
        mine = np.stack([out[cam * 8:cam * 8 + 8:2].ravel(), out[cam * 8 + 1:cam * 8 + 8:2].ravel()], axis=1)
        ok = np.isfinite(ref).all(axis=1) & (np.abs(ref).max(axis=1) < 50)   # newer OpenCV flags non-converged points
        assert ok.mean() > 0.9
        assert np.abs(mine[ok] - ref[ok]).max() <= 2e-6 * max(1.0, np.abs(ref[ok]).max())
    # the principal point maps to the origin
    pp = np.zeros((16, 1), dtype=np.float32)
    for e in range(8):
        pp[2 * e, 0], pp[2 * e + 1, 0] = cfg.cam_k[e >> 2][2], cfg.cam_k[e >> 2][3]
    assert np.abs(orc.undistort_fisheye(cfg, pp)).max() < 1e-6


@pytest.mark.gpu
def test_gpu_undistort_and_chain(cfg):
    """pixels -> undistort (GPU) == oracle bit for bit; and the result feeds the refractive solve unchanged"""
    import orc
    from fbus_ekf_b200 import BatchFilter
    px = _pixels(5000 + 3, 1)
    f = BatchFilter(cfg, batch=1)
    out = f.UndistortFisheye(px)
    ref = orc.undistort_fisheye(cfg, px)
    assert np.abs(out - ref).max() <= 2e-7 * max(1.0, float(np.abs(ref).max()))
    assert (out == ref).mean() > 0.99          # identical but for rare 1-ulp float32 rounding of libm vs CUDA tan
    pose, c3, valid = f.RefractSolve(out)
    assert pose.shape == (7, px.shape[1])
