"""one fused launch of a small batch (for ncu): python profiles/probes/lane_prof_run.py [B]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbus_ekf_b200 import BatchFilter, capi, synth  # noqa: E402
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
cfg = capi.config_default()
traj = synth.truth_trajectory(cfg, 1.0, 200.0, 25.0, periodic=True)
N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
dev = torch.device("cuda:0")
f = BatchFilter(cfg, batch=B, device=0)
imu_d = torch.empty((N, 6, B), dtype=torch.float64, device=dev)
id_d = torch.empty((W, 1, B), dtype=torch.int32, device=dev)
pose_d = torch.empty((W, 1, 7, B), dtype=torch.float64, device=dev)
f.SynthStreams(synth.make_synth_spec(traj, seed=3), imu_d.data_ptr(), id_d.data_ptr(), pose_d.data_ptr())
for k in range(4):
    ti, tf = traj["t_imu"] + k * 1.0, traj["t_frames"] + k * 1.0
    f.StepWindows(capi.make_imu_stream(ti, imu_d.data_ptr(), B, capi.FBUS_MEM_DEVICE),
                  capi.make_det_frames(tf, id_d.data_ptr(), pose_d.data_ptr(), B, 1, capi.FBUS_MEM_DEVICE), traj["win_off"], 0, W)
f.Synchronize()
