"""FBUS_IMU_F32_SENSOR: IMU samples as the IMSEE SDK delivers them (float32, accel in g, gyro in deg/s;
driver/IMSEE-SDK/include/types.h:122-127) converted on the device exactly as the reference's IMU callback does
(main.cpp:254).  The filter must see the same doubles as with the converted stream: results are compared BITWISE with the
FBUS_IMU_F64_SI path on `capi.sensor_to_si(raw)`, and to 1e-9 with the oracle."""
import numpy as np
import pytest

from helpers import cov_close

pytestmark = pytest.mark.gpu
KEYS = ("t", "q", "R", "p", "v", "ba", "bg", "g", "P", "status", "initialised", "prev_marker_id")


def _streams(cfg, B, duration, seed):
    import torch
    from fbus_ekf_b200 import BatchFilter, capi, synth
    traj = synth.truth_trajectory(cfg, duration)
    N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
    f = BatchFilter(cfg, batch=B)
    imu_d = torch.empty((N, 6, B), dtype=torch.float64, device="cuda")
    id_d = torch.empty((W, 1, B), dtype=torch.int32, device="cuda")
    pose_d = torch.empty((W, 1, 7, B), dtype=torch.float64, device="cuda")
    f.SynthStreams(synth.make_synth_spec(traj, seed=seed), imu_d.data_ptr(), id_d.data_ptr(), pose_d.data_ptr())
    f.close()
    raw = capi.si_to_sensor(imu_d.cpu().numpy(), cfg.imu_g)          # what the sensor would deliver
    si = np.ascontiguousarray(capi.sensor_to_si(raw, cfg.imu_g))      # what main.cpp:254 makes of it
    return traj, np.ascontiguousarray(raw), si, np.ascontiguousarray(id_d.cpu().numpy()), np.ascontiguousarray(pose_d.cpu().numpy())


def test_conversion_constants(cfg):
    from fbus_ekf_b200 import capi
    assert cfg.imu_g == 9.802                                         # camerainfo1.yml "g"
    raw = np.array([[[1.0], [-0.5], [0.25], [180.0], [90.0], [-45.0]]], dtype=np.float32)
    si = capi.sensor_to_si(raw, cfg.imu_g)
    assert si[0, 0, 0] == 9.802 and si[0, 3, 0] == 3.1415926         # the reference's own M_PI (common.hpp:14)
    # float division first, as `imu.gyro[0]/180*M_PI` evaluates: differs from the double division for most inputs
    x = np.float32(0.7)
    assert float(np.float32(x / np.float32(180.0))) * 3.1415926 == capi.sensor_to_si(np.array([[[0], [0], [0], [x], [0], [0]]], dtype=np.float32), 1.0)[0, 3, 0]


@pytest.mark.parametrize("where", ["host", "device"])
def test_fused_windows_bitwise_and_oracle(cfg, where):
    import orc
    import torch
    from fbus_ekf_b200 import BatchFilter, capi
    B = 320
    traj, raw, si, ids, pose = _streams(cfg, B, 0.6, seed=41)
    W = len(traj["t_frames"])
    det = capi.make_det_frames(traj["t_frames"], ids, pose, B, 1)
    fa = BatchFilter(cfg, batch=B)
    fa.StepWindows(capi.make_imu_stream(traj["t_imu"], si, B), det, traj["win_off"], 0, W)
    a = fa.GetState()
    fb = BatchFilter(cfg, batch=B)
    if where == "host":
        imu32 = capi.make_imu_stream(traj["t_imu"], raw, B)
        assert imu32.format == capi.FBUS_IMU_F32_SENSOR
        tr = fb.StepWindows(imu32, det, traj["win_off"], 0, W, trace=True)
    else:
        raw_d = torch.from_numpy(raw).cuda()
        id_d, pose_d = torch.from_numpy(ids).cuda(), torch.from_numpy(pose).cuda()
        det_d = capi.make_det_frames(traj["t_frames"], id_d.data_ptr(), pose_d.data_ptr(), B, 1, capi.FBUS_MEM_DEVICE)
        imu32 = capi.make_imu_stream(traj["t_imu"], raw_d.data_ptr(), B, capi.FBUS_MEM_DEVICE, fmt=capi.FBUS_IMU_F32_SENSOR)
        # two calls: the second continues the device-resident stream (per-filter cursor)
        fb.StepWindows(imu32, det_d, traj["win_off"], 0, W // 2)
        tr = fb.StepWindows(imu32, det_d, traj["win_off"], W // 2, W, trace=True)
    b = fb.GetState()
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(tr[-1, 1:4, :], b["p"])
    o = orc.Oracle(cfg, B)
    o.step_windows(capi.make_imu_stream(traj["t_imu"], si, B), det, traj["win_off"], 0, W, None, 8)
    so = o.get_state()
    for k in ("t", "q", "p", "v", "ba", "bg", "g"):
        assert np.abs(b[k] - so[k]).max() <= 1e-9, k
    assert cov_close(b["P"], so["P"], 1e-9)[0]
    assert np.array_equal(b["status"], so["status"])


def test_pipelined_host_stream_bitwise(cfg):
    """large host-resident float32 stream through the chunked copy / compute pipeline (half the bytes per sample)"""
    from fbus_ekf_b200 import BatchFilter, capi
    B = 16384
    traj, raw, si, ids, pose = _streams(cfg, B, 1.0, seed=43)
    W = len(traj["t_frames"])
    assert raw.nbytes >= (64 << 20) and W >= 10                        # above the pipeline thresholds
    det = capi.make_det_frames(traj["t_frames"], ids, pose, B, 1)
    fa = BatchFilter(cfg, batch=B)
    fa.StepWindows(capi.make_imu_stream(traj["t_imu"], si, B), det, traj["win_off"], 0, W)
    a = fa.GetState()
    fb = BatchFilter(cfg, batch=B)
    fb.StepWindows(capi.make_imu_stream(traj["t_imu"], raw, B), det, traj["win_off"], 0, W)
    b = fb.GetState()
    for k in KEYS:
        assert np.array_equal(a[k], b[k]), k


def test_unfused_calls_and_bad_format(cfg):
    """InitGravityAndGyrobias / ImuUpdate take the sensor format too; an unknown format is FBUS_E_BADARG"""
    from fbus_ekf_b200 import BatchFilter, capi
    B = 96
    traj, raw, si, ids, pose = _streams(cfg, B, 0.3, seed=47)
    det = capi.make_det_frames(traj["t_frames"], ids, pose, B, 1)
    out = []
    for data in (si, raw):
        f = BatchFilter(cfg, batch=B)
        imu = capi.make_imu_stream(traj["t_imu"], data, B)
        f.InitGravityAndGyrobias(imu, 0, 20)
        f.InitPositionAndQuaternion(det, 0, 1)
        f.ImuUpdate(imu, 20, 30)
        f.MeasureUpdate(det, 1)
        out.append(f.GetState())
    for k in KEYS:
        assert np.array_equal(out[0][k], out[1][k]), k
    assert np.abs(out[0]["bg"]).max() > 0
    f = BatchFilter(cfg, batch=B)
    bad = capi.make_imu_stream(traj["t_imu"], si, B)
    bad.format = 7
    import ctypes as C
    assert capi.lib().fbus_propagate(f._h, C.byref(bad), 0, 5, float("inf")) == capi.FBUS_E_BADARG
    assert capi.lib().fbus_step_windows(f._h, C.byref(bad), C.byref(det), capi.dptr(np.ascontiguousarray(traj["win_off"], dtype=np.uint32), capi.c_uint32_p), 0, 1, None, 0) == capi.FBUS_E_BADARG


def test_iir_prefilter_on_the_device(cfg, golden):
    """fbus_iir_prefilter == the oracle's restatement of SetImuData's recurrence, bit for bit (two rounded products and a
    rounded sum per sample, no FMA); both element formats, host and device outputs, in place, restarts"""
    import fbus_oracle_np
    import torch
    from fbus_ekf_b200 import BatchFilter, capi, replay
    imu = golden["land_imu"][:6000]
    want = fbus_oracle_np.iir_prefilter(imu, restart_at=(500, 4100))
    got = replay.iir_prefilter(imu, restart_at=(500, 4100))
    assert np.array_equal(got, want)
    # batch of filters, float32 sensor samples in, device-resident out
    B = 130
    traj, raw, si, ids, pose = _streams(cfg, B, 0.5, seed=51)
    N = raw.shape[0]
    f = BatchFilter(cfg, batch=B)
    ref = np.empty_like(si)
    for b in range(0, B, 43):
        rows = np.concatenate([traj["t_imu"][:, None], si[:, :, b]], axis=1)
        ref[:, :, b] = fbus_oracle_np.iir_prefilter(rows)[:, 1:7]
    out_h = f.IirPrefilter(capi.make_imu_stream(traj["t_imu"], raw, B), 0, N)
    assert np.array_equal(out_h[:, :, ::43], ref[:, :, ::43])
    si_d = torch.from_numpy(si).cuda()
    s_d = capi.make_imu_stream(traj["t_imu"], si_d.data_ptr(), B, capi.FBUS_MEM_DEVICE)
    f.IirPrefilter(s_d, 0, N, out=si_d.data_ptr(), mem=capi.FBUS_MEM_DEVICE)   # in place
    f.Synchronize()
    assert np.array_equal(si_d.cpu().numpy(), out_h)
    part = f.IirPrefilter(capi.make_imu_stream(traj["t_imu"], si, B), 10, 25)
    assert np.array_equal(part[0], si[10]) and part.shape == (25, 6, B)
