"""Phase timing of the second-generation lanes-per-filter kernel from clock64 stamps (library built with -DFBUS_L2_TRACE:
python -c "from fbus_ekf_b200 import build; build.build_variant(128, 'l2trace', ['-DFBUS_L2_TRACE'])", then
FBUS_EKF_LIB=fbus_ekf_b200/libfbus_ekf_l2trace.so python profiles/probes/lane2_trace.py [B])."""
import collections
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fbus_ekf_b200 import BatchFilter, capi, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
os.environ["FBUS_LANE"] = "1"
cfg = capi.config_default()
traj = synth.truth_trajectory(cfg, 1.0, 200.0, 25.0, periodic=True)
N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
dev = torch.device("cuda:0")
f = BatchFilter(cfg, batch=B, device=0)
imu_d = torch.empty((N, 6, B), dtype=torch.float64, device=dev)
id_d = torch.empty((W, 1, B), dtype=torch.int32, device=dev)
pose_d = torch.empty((W, 1, 7, B), dtype=torch.float64, device=dev)
f.SynthStreams(synth.make_synth_spec(traj, seed=3), imu_d.data_ptr(), id_d.data_ptr(), pose_d.data_ptr())
L = capi.lib()
buf = (C.c_longlong * (2 * 8192))()


def run(k):
    ti, tf = traj["t_imu"] + k * 1.0, traj["t_frames"] + k * 1.0
    f.StepWindows(capi.make_imu_stream(ti, imu_d.data_ptr(), B, capi.FBUS_MEM_DEVICE),
                  capi.make_det_frames(tf, id_d.data_ptr(), pose_d.data_ptr(), B, 1, capi.FBUS_MEM_DEVICE), traj["win_off"], 0, W)
    f.Synchronize()


for k in range(3):
    run(k)
L.fbus_debug_l2_trace(buf, 0, 8192)
L.fbus_debug_l2_trace(buf, 1, 8192)
run(3)
for role, name in ((0, "nominal lane 0"), (1, "covariance warp 0")):
    n = L.fbus_debug_l2_trace(buf, role, 8192)
    a = np.array(buf[: 2 * n], dtype=np.int64).reshape(n, 2)
    # split into frames at tag 1, drop the first two frames (initialisation), average the time of each tag relative to the frame start
    starts = np.nonzero(a[:, 0] == 1)[0]
    acc = collections.OrderedDict()
    cnt = 0
    for s0, s1 in zip(starts[2:-1], starts[3:]):
        fr = a[s0:s1 + 1]
        t0 = fr[0, 1]
        for j, (tag, t) in enumerate(fr):
            acc.setdefault((j, int(tag)), []).append(int(t - t0))
        cnt += 1
    print(f"{name}: {cnt} frames; clk since the frame start (mean) and step from the previous stamp")
    prev = 0
    for (j, tag), v in acc.items():
        m = float(np.mean(v))
        print(f"  tag {tag:3d}: {m:9.0f}  (+{m - prev:7.0f})")
        prev = m
f.close()
