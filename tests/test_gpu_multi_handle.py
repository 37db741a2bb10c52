"""BASELINE configs[4] from a C++ host through the C ABI alone (tests/cpp/multi_gpu_demo.cpp): one process, one handle per
GPU, contiguous shards, and fbus_stats_allreduce (NCCL, loaded at run time) as the only collective.  The combined statistics
equal those of ONE handle holding the whole batch (shards = whole: the Philox streams are keyed by the global filter index).
With one visible GPU the collective degenerates (n = 1, no NCCL); with two or more it is a real ncclAllReduce."""
import os
import struct
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path):
    exe = tmp_path / "multi_gpu_demo"
    libdir = os.path.join(ROOT, "fbus_ekf_b200")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "cpp", "multi_gpu_demo.cpp"),
                           "-I/usr/local/cuda/include", "-L/usr/local/cuda/lib64", "-L" + libdir, "-lfbus_ekf", "-lcudart",
                           "-Wl,-rpath," + libdir, "-Wl,-rpath,/usr/local/cuda/lib64", "-pthread"])
    return exe


def _traj_file(cfg, path):
    from fbus_ekf_b200 import synth
    traj = synth.truth_trajectory(cfg, 1.0, 200.0, 25.0, periodic=True)
    N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
    with open(path, "wb") as f:
        f.write(struct.pack("qq", N, W))
        for a in (traj["t_imu"], traj["t_frames"], traj["base_imu"], traj["base_pose"], traj["truth_p"][-1], traj["truth_q"][-1]):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(traj["win_off"], dtype=np.uint32).tobytes())
    return traj


def _whole(cfg, traj, total, seconds):
    """the same workload on ONE handle through the Python binding"""
    import torch
    from fbus_ekf_b200 import BatchFilter, capi, synth
    N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
    dev = torch.device("cuda:0")
    f = BatchFilter(cfg, batch=total, device=0)
    imu_d = torch.empty((N, 6, total), dtype=torch.float64, device=dev)
    id_d = torch.empty((W, 1, total), dtype=torch.int32, device=dev)
    pose_d = torch.empty((W, 1, 7, total), dtype=torch.float64, device=dev)
    f.SynthStreams(synth.make_synth_spec(traj, seed=20260117 + 5), imu_d.data_ptr(), id_d.data_ptr(), pose_d.data_ptr())
    for k in range(seconds):
        f.StepWindows(capi.make_imu_stream(traj["t_imu"] + k, imu_d.data_ptr(), total, capi.FBUS_MEM_DEVICE),
                      capi.make_det_frames(traj["t_frames"] + k, id_d.data_ptr(), pose_d.data_ptr(), total, 1, capi.FBUS_MEM_DEVICE),
                      traj["win_off"], 0, W)
    tp = np.ascontiguousarray(np.repeat(traj["truth_p"][-1][:, None], total, axis=1))
    tq = np.ascontiguousarray(np.repeat(traj["truth_q"][-1][:, None], total, axis=1))
    return f.Stats(tp, tq)


@pytest.mark.parametrize("n_gpu", [1, 2, 8])
def test_cpp_host_shards_and_allreduce(cfg, tmp_path, n_gpu):
    import torch
    if torch.cuda.device_count() < n_gpu:
        pytest.skip(f"needs {n_gpu} GPUs")
    exe = _build(tmp_path)
    traj = _traj_file(cfg, tmp_path / "traj.bin")
    total, seconds = 10007, 2  # odd total: shard sizes differ by one
    out = subprocess.check_output([str(exe), str(tmp_path / "traj.bin"), str(total), str(n_gpu), str(seconds)], text=True)
    lines = dict((ln.split()[0], ln.split()[1:]) for ln in out.strip().splitlines())
    stats = np.array([float(x) for x in lines["stats"]])
    shards = [int(x) for x in lines["shards"]]
    assert sum(shards) == total and max(shards) - min(shards) <= 1 and len(shards) == n_gpu
    ref = _whole(cfg, traj, total, seconds)
    assert stats[3] == ref[3] == total and stats[4] == ref[4] == 0
    assert np.allclose(stats[:3], ref[:3], rtol=1e-12, atol=0)  # sums in a different order
    assert stats[5] == ref[5]                                   # the maximum is exact
    assert stats[6] == 0 and stats[7] == 0


def test_allreduce_argument_errors(cfg):
    import ctypes as C
    from fbus_ekf_b200 import BatchFilter, capi
    lib = capi.lib()
    f1, f2 = BatchFilter(cfg, batch=4), BatchFilter(cfg, batch=4)
    import torch
    v = torch.zeros(16, dtype=torch.float64, device="cuda:0")
    hs = (C.c_void_p * 2)(f1._h, f2._h)
    vs = (C.c_void_p * 2)(v.data_ptr(), v.data_ptr() + 64)
    out = np.zeros(8)
    # two handles on the same device: refused (those are combined with fbus_stats_combine)
    assert lib.fbus_stats_allreduce(hs, 2, vs, capi.dptr(out)) == capi.FBUS_E_BADARG
    assert lib.fbus_stats_allreduce(hs, 0, vs, capi.dptr(out)) == capi.FBUS_E_BADARG
    # n = 1: the vector is returned as it is (entries 6..7 zeroed)
    v[:8] = torch.arange(1, 9, dtype=torch.float64)
    assert lib.fbus_stats_allreduce(hs, 1, vs, capi.dptr(out)) == 0
    assert np.array_equal(out, [1, 2, 3, 4, 5, 6, 0, 0])
