"""N>1 path on CPU: two ranks over gloo each own a contiguous shard of the filters (no data-path collective), run their
shard, and all-reduce the statistics vector.  The result must equal the single-rank run of the whole batch -- the same
host logic bench.py uses with NCCL on the GPUs (the per-shard filter arithmetic here is the oracle's)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _make_inputs(B):
    from fbus_ekf_b200 import capi, synth
    cfg = capi.config_default()
    traj = synth.truth_trajectory(cfg, 0.4)
    N, W = traj["base_imu"].shape[0], traj["base_pose"].shape[0]
    rng = np.random.default_rng(2026)
    imu = traj["base_imu"][:, :, None] + rng.normal(size=(N, 6, B)) * 0.01
    pose = np.repeat(traj["base_pose"][:, None, :, None], B, axis=3) + rng.normal(size=(W, 1, 7, B)) * 3e-4
    ids = np.zeros((W, 1, B), dtype=np.int32)
    return cfg, traj, imu, ids, pose


def _run_shard(cfg, traj, imu, ids, pose, lo, hi):
    import orc
    from fbus_ekf_b200 import capi
    b = hi - lo
    s = capi.make_imu_stream(traj["t_imu"], np.ascontiguousarray(imu[:, :, lo:hi]), b)
    d = capi.make_det_frames(traj["t_frames"], np.ascontiguousarray(ids[:, :, lo:hi]), np.ascontiguousarray(pose[:, :, :, lo:hi]), b, 1)
    o = orc.Oracle(cfg, b)
    o.step_windows(s, d, traj["win_off"], 0, len(traj["t_frames"]))
    tp = np.ascontiguousarray(np.repeat(traj["truth_p"][-1][:, None], b, axis=1))
    tq = np.ascontiguousarray(np.repeat(traj["truth_q"][-1][:, None], b, axis=1))
    return o.stats(tp, tq), o.get_state(with_cov=False)


def _worker(rank, world, port, B, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch
    import torch.distributed as dist
    from fbus_ekf_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, traj, imu, ids, pose = _make_inputs(B)
    lo, hi = shard.shard_range(B, rank, world)
    st, state = _run_shard(cfg, traj, imu, ids, pose, lo, hi)
    vec = torch.from_numpy(st.copy())
    shard.combine_stats(vec, dist)
    np.save(os.path.join(out_dir, f"stats_{rank}.npy"), vec.numpy())
    np.save(os.path.join(out_dir, f"p_{rank}.npy"), state["p"])
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range():
    from fbus_ekf_b200 import shard
    for total, world in ((10, 3), (1 << 20, 8), (5, 8)):
        r = [shard.shard_range(total, k, world) for k in range(world)]
        assert r[0][0] == 0 and r[-1][1] == total
        assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_two_ranks_equal_one(built, tmp_path):
    import torch.multiprocessing as mp
    from fbus_ekf_b200 import shard
    B, world = 24, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    cfg, traj, imu, ids, pose = _make_inputs(B)
    full, state = _run_shard(cfg, traj, imu, ids, pose, 0, B)
    s0, s1 = np.load(tmp_path / "stats_0.npy"), np.load(tmp_path / "stats_1.npy")
    assert np.array_equal(s0, s1)                      # every rank holds the reduced vector
    assert np.allclose(s0[:5], full[:5], rtol=1e-13, atol=0) and s0[5] == full[5]
    assert s0[3] == B and s0[4] == 0
    # shards never communicate: the concatenated shard states are bitwise the single-rank states
    p = np.concatenate([np.load(tmp_path / "p_0.npy"), np.load(tmp_path / "p_1.npy")], axis=1)
    assert np.array_equal(p, state["p"])
    summ = shard.summarize_stats(s0)
    assert summ["filters_finite"] == B and summ["rmse_pos_m"] < 0.05


def test_stats_combine_c_abi(built):
    """the host-side combine rule exported by the C ABI equals shard.combine_stats' rule (sums for [0..4], max for [5])"""
    import ctypes as C
    from fbus_ekf_b200 import capi
    rng = np.random.default_rng(5)
    parts = np.ascontiguousarray(rng.uniform(0.0, 10.0, (3, capi.FBUS_NSTATS)))
    out = np.full(capi.FBUS_NSTATS, -1.0)
    rc = capi.lib().fbus_stats_combine(capi.dptr(parts), 3, capi.dptr(out))
    assert rc == 0
    assert np.allclose(out[:5], parts[:, :5].sum(axis=0), rtol=0, atol=1e-12)
    assert out[5] == parts[:, 5].max() and out[6] == 0.0 and out[7] == 0.0
    assert capi.lib().fbus_stats_combine(None, 0, capi.dptr(out)) != 0
