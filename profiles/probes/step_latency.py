"""Latency of ONE filter group through the window kernels, split into its parts: a pure propagate call (2000 IMU samples, no
detections: time per IMU sample) and the fused call (200 samples + 25 updates).  Run on a B200: python profiles/probes/step_latency.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbus_ekf_b200 import BatchFilter, capi, synth  # noqa: E402

cfg = capi.config_default()
traj = synth.truth_trajectory(cfg, 1.0, 200.0, 25.0, periodic=True)
N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
dev = torch.device("cuda:0")
FORMS = {"lane9": {"FBUS_LANE": "1"}, "lane9-gen1": {"FBUS_LANE": "1", "FBUS_LANE_GEN": "1"}, "smem32": {"FBUS_LANE": "0", "FBUS_SMALL_BATCH": "1"}, "tmem128": {"FBUS_LANE": "0", "FBUS_SMALL_BATCH": "0"}}
print("| kernel | filters | us per IMU sample (propagate only) | us per frame (8 samples + update) | => us per update |")
print("|---|---|---|---|---|")
for B in (1, 32, 128):
    for name, env in FORMS.items():
        for k in ("FBUS_LANE", "FBUS_LANE_GEN", "FBUS_SMALL_BATCH"):
            os.environ.pop(k, None)
        os.environ.update(env)
        f = BatchFilter(cfg, batch=B, device=0)
        imu_d = torch.empty((N, 6, B), dtype=torch.float64, device=dev)
        id_d = torch.empty((W, 1, B), dtype=torch.int32, device=dev)
        pose_d = torch.empty((W, 1, 7, B), dtype=torch.float64, device=dev)
        f.SynthStreams(synth.make_synth_spec(traj, seed=3), imu_d.data_ptr(), id_d.data_ptr(), pose_d.data_ptr())
        stream = torch.cuda.ExternalStream(f.stream, device=dev)
        imu = capi.make_imu_stream(traj["t_imu"], imu_d.data_ptr(), B, capi.FBUS_MEM_DEVICE)
        det = capi.make_det_frames(traj["t_frames"], id_d.data_ptr(), pose_d.data_ptr(), B, 1, capi.FBUS_MEM_DEVICE)
        f.StepWindows(imu, det, traj["win_off"], 0, W)  # initialise
        f.Synchronize()

        def timed(fn, reps):
            fn(0)
            f.Synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for r in range(reps):
                fn(1 + r)
            e1.record(stream)
            f.Synchronize()
            return e0.elapsed_time(e1) * 1e-3 / reps

        def prop(k):  # 10 periods of the stream in one un-fused propagate call each: 200 samples per launch
            ti = traj["t_imu"] + (k + 1) * 1.0
            f.ImuUpdate(capi.make_imu_stream(ti, imu_d.data_ptr(), B, capi.FBUS_MEM_DEVICE), 0, N)

        def fused(k):
            ti, tf = traj["t_imu"] + (k + 40) * 1.0, traj["t_frames"] + (k + 40) * 1.0
            f.StepWindows(capi.make_imu_stream(ti, imu_d.data_ptr(), B, capi.FBUS_MEM_DEVICE),
                          capi.make_det_frames(tf, id_d.data_ptr(), pose_d.data_ptr(), B, 1, capi.FBUS_MEM_DEVICE), traj["win_off"], 0, W)
        tp = timed(prop, 10)
        tf_ = timed(fused, 10)
        us_step = tp / N * 1e6
        us_frame = tf_ / W * 1e6
        print(f"| {name} | {B} | {us_step:.3f} | {us_frame:.2f} | {us_frame - 8 * us_step:.2f} |", flush=True)
        f.close()
