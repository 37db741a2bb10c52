"""Small batches: filter-steps/s of the two generations of the lanes-per-filter kernel and the 32-filter shared-memory CTAs.
python profiles/probes/lane_gen_sweep.py [B ...]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbus_ekf_b200 import BatchFilter, capi, synth  # noqa: E402

cfg = capi.config_default()
traj = synth.truth_trajectory(cfg, 1.0, 200.0, 25.0, periodic=True)
N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
dev = torch.device("cuda:0")
FORMS = (("lane gen 2", {"FBUS_LANE": "1", "FBUS_LANE_GEN": "2"}), ("lane gen 1", {"FBUS_LANE": "1", "FBUS_LANE_GEN": "1"}),
         ("smem32", {"FBUS_LANE": "0", "FBUS_SMALL_BATCH": "1"}))
print("| filters | " + " | ".join(n for n, _ in FORMS) + " | us per frame (gen 2 / gen 1 / smem32) |")
print("|---|---|---|---|---|")
SIZES = tuple(int(x) for x in sys.argv[1:]) or (1, 2, 4, 8, 16, 32, 64, 256, 1024, 4096)
for B in SIZES:
    row = []
    for _, env in FORMS:
        for k in ("FBUS_LANE", "FBUS_LANE_GEN", "FBUS_SMALL_BATCH"):
            os.environ.pop(k, None)
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
        row.append((B * (N + W) / sec, sec / W * 1e6))
        f.close()
    print(f"| {B} | " + " | ".join(f"{r[0]:.3g}" for r in row) + " | " + " / ".join(f"{r[1]:.2f}" for r in row) + " |", flush=True)
