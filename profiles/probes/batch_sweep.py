"""Throughput of the window kernel vs batch size for its three forms (nine lanes per filter, FBUS_LANE=1; 32-filter
shared-memory CTAs, FBUS_SMALL_BATCH=1; 128-filter tensor-memory CTAs, FBUS_SMALL_BATCH=0): where should fbus_create
switch?  Run on a B200: python profiles/probes/batch_sweep.py"""
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
print("| filters | 9 lanes per filter (registers) | 32-filter CTAs (smem) | 128-filter CTAs (TMEM) |")
print("|---|---|---|---|")
SIZES = (1, 32, 256, 1024, 2048, 4096, 4736, 6144, 8192, 9472, 10240, 12288, 14336, 16384, 18944, 32768, 65536)
if len(sys.argv) > 1:
    SIZES = tuple(int(x) for x in sys.argv[1:])
for B in SIZES:
    row = []
    for env in ({"FBUS_LANE": "1"}, {"FBUS_LANE": "0", "FBUS_SMALL_BATCH": "1"}, {"FBUS_LANE": "0", "FBUS_SMALL_BATCH": "0"}):
        os.environ.update(env)
        f = BatchFilter(cfg, batch=B, device=0)
        imu_d = torch.empty((N, 6, B), dtype=torch.float64, device=dev)
        id_d = torch.empty((W, 1, B), dtype=torch.int32, device=dev)
        pose_d = torch.empty((W, 1, 7, B), dtype=torch.float64, device=dev)
        f.SynthStreams(synth.make_synth_spec(traj, seed=3), imu_d.data_ptr(), id_d.data_ptr(), pose_d.data_ptr())
        stream = torch.cuda.ExternalStream(f.stream, device=dev)

        def step(k):
            ti = traj["t_imu"] + k * 1.0
            tf = traj["t_frames"] + k * 1.0
            f.StepWindows(capi.make_imu_stream(ti, imu_d.data_ptr(), B, capi.FBUS_MEM_DEVICE),
                          capi.make_det_frames(tf, id_d.data_ptr(), pose_d.data_ptr(), B, 1, capi.FBUS_MEM_DEVICE), traj["win_off"], 0, W)
        for k in range(3):
            step(k)
        f.Synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(stream)
        for k in range(reps):
            step(3 + k)
        e1.record(stream)
        f.Synchronize()
        sec = e0.elapsed_time(e1) * 1e-3 / reps
        row.append(B * (N + W) / sec)
        f.close()
    print(f"| {B} | {row[0]:.3g} | {row[1]:.3g} | {row[2]:.3g} |", flush=True)
