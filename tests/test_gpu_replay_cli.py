"""The replay driver as a caller of the path: dataset directory in the reference's text formats ->
`python -m fbus_ekf_b200.replay` -> data/fusion.txt format; and the bounded IMU buffer against the C++ FILTER shim,
which pushes and erases sample by sample like FILTER::SetImuData."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _round6(a):
    return np.array([[float("%g" % v) for v in r] for r in a])


def test_cli_land_and_water(cfg, golden, tmp_path):
    from fbus_ekf_b200 import logio, replay
    for name, extra in (("land", []), ("water", ["--water"])):
        d = tmp_path / name
        d.mkdir()
        imu = golden[f"{name}_imu"][:9000]
        img = golden[f"{name}_image"]
        img = img[img[:, 0] <= imu[-1, 0]]
        logio.write_log(str(d / "imu.txt"), imu, "imu", newline="\r\n")
        logio.write_log(str(d / "image.txt"), img, "image", newline="\r\n")
        cor = golden[f"{name}_corners"]
        cor = cor[cor[:, 0] <= imu[-1, 0]]
        logio.write_log(str(d / "corners.txt"), cor, "corners", newline="\r\n")
        out = d / "fusion_out.txt"
        assert replay.main([str(d), "--out", str(out)] + extra) == 0
        got = logio.read_fusion_log(str(out))
        if extra:  # poses solved from corners.txt on the GPU instead of the logged image.txt
            img = replay.solve_image_rows(cor, cfg)
            assert len(img) == len(cor)
            assert np.abs(img[:, 2:5] - golden["water_image"][:len(img), 2:5]).max() <= 2e-5
        res = replay.replay_log(imu, img, cfg)
        # the reference records no row for the frame that initialises the pose (filter.cpp:207-226): the log starts one frame later
        assert not res["recorded"][0] and res["recorded"][1:].all()
        want = res["rows"][res["recorded"]]
        assert got.shape == want.shape and np.array_equal(got, _round6(want)), name
        assert np.isfinite(got).all()


def test_cli_errors(tmp_path):
    from fbus_ekf_b200 import replay
    with pytest.raises(SystemExit):
        replay.main([str(tmp_path)])  # no imu.txt


def test_buffer_cap_equals_the_shim(cfg, golden, tmp_path):
    """a 3 s stretch without detections: the shim's live buffer (push, erase 500 beyond 2000) and the replay driver's
    closed form keep the same samples -> identical fusion rows"""
    from fbus_ekf_b200 import replay
    exe = tmp_path / "shim_demo"
    libdir = os.path.join(ROOT, "fbus_ekf_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "shim_demo.cpp"),
                           "-L" + libdir, "-lfbus_ekf", "-Wl,-rpath," + libdir, "-pthread"])
    imu = golden["land_imu"][:12000]
    img = golden["land_image"]
    img = img[img[:, 0] <= imu[-1, 0]]
    t0 = img[40, 0]
    # choose the end of the gap so that the buffer is far from an erase when the next frame arrives (the shim has three
    # samples later than the frame in its buffer, the deterministic replay none)
    for gap in np.arange(2.6, 3.4, 0.04):
        rest = img[img[:, 0] > t0 + gap]
        cnt = np.searchsorted(imu[:, 0], rest[0, 0], side="right") - np.searchsorted(imu[:, 0], t0, side="right")
        if cnt > 2100 and 50 <= (cnt - 2001) % 500 <= 440:
            break
    else:
        pytest.fail("no suitable gap")
    img = np.concatenate([img[:41], rest])
    keep = replay.buffer_cap_keep(imu[:, 0], replay.group_frames(img)[0], 500)
    assert 0 < (~keep).sum() < cnt
    np.savetxt(tmp_path / "imu.txt", imu, fmt="%.17g")
    np.savetxt(tmp_path / "image.txt", img, fmt="%.17g")
    out = subprocess.check_output([str(exe), str(tmp_path / "imu.txt"), str(tmp_path / "image.txt"), "500", "0"], text=True)
    rows = np.array([[float(x) for x in line.split()] for line in out.strip().splitlines()])
    ref = replay.replay_log(imu, img, cfg, n_init=500, buffer_cap=replay.IMU_BUFFER_MAX_SIZE)["rows"]
    assert rows.shape == ref.shape and np.array_equal(rows, ref), np.abs(rows - ref).max()
