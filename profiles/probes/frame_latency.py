"""Where the time of one live frame goes (one filter, 8 IMU samples + 1 marker pose per fbus_step_windows call, host arrays):
time until the call returns (packing, one H2D copy, the launch) and time until the synchronisation returns, through the ctypes
binding.  python profiles/probes/frame_latency.py"""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbus_ekf_b200 import BatchFilter, capi, synth  # noqa: E402

cfg = capi.config_default()
traj = synth.truth_trajectory(cfg, 1.0, 200.0, 25.0, periodic=True)
N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
lib = capi.lib()
for env in ({}, {"FBUS_LANE": "0", "FBUS_SMALL_BATCH": "1"}):
    for k in ("FBUS_LANE", "FBUS_SMALL_BATCH"):
        os.environ.pop(k, None)
    os.environ.update(env)
    f1 = BatchFilter(cfg, batch=1, device=0)
    imu1 = np.ascontiguousarray(traj["base_imu"][:, :, None])
    id1 = np.zeros((W, 1, 1), dtype=np.int32)
    pose1 = np.ascontiguousarray(traj["base_pose"][:, None, :, None])
    off = np.ascontiguousarray(traj["win_off"], dtype=np.uint32)
    offp = capi.dptr(off, capi.c_uint32_p)
    t_call, t_sync = [], []
    for kk in range(6):
        ti, tf = traj["t_imu"] + kk, traj["t_frames"] + kk
        imu_v = capi.make_imu_stream(ti, imu1, 1)
        det_v = capi.make_det_frames(tf, id1, pose1, 1, 1)
        pi, pd = C.byref(imu_v), C.byref(det_v)
        for w in range(W):
            t0 = time.perf_counter()
            lib.fbus_step_windows(f1._h, pi, pd, offp, w, w + 1, None, 0)
            t1 = time.perf_counter()
            lib.fbus_synchronize(f1._h)
            t2 = time.perf_counter()
            if kk > 0:
                t_call.append(t1 - t0)
                t_sync.append(t2 - t1)
    a, b = np.array(t_call) * 1e6, np.array(t_sync) * 1e6
    print(f"{env or 'default (lanes-per-filter kernel)'}: call returns after {np.median(a):.1f} us (p95 {np.percentile(a, 95):.1f}), "
          f"synchronise adds {np.median(b):.1f} us (p95 {np.percentile(b, 95):.1f}), total median {np.median(a + b):.1f} us")
    f1.close()
