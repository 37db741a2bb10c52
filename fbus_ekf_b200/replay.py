"""Deterministic replay driver over the C ABI: the caller of the hot path for recorded logs.

Models the loop of matlab/FBUS_EKF.m:116-197 with the shipped C++ semantics: gravity / gyro-bias
initialisation from the first `n_init` IMU rows, pose initialisation at the first detection frame, then per frame
ResetSystemState -> BatchImuProcessing -> ObservationUpdate.  Host code only prepares the SoA streams and the
per-frame IMU windows; every arithmetic step runs on the GPU through BatchFilter.
"""
from __future__ import annotations

import numpy as np

from . import capi
from .filter import BatchFilter


def iir_prefilter(imu: np.ndarray, restart_at=()) -> np.ndarray:
    """FILTER::SetImuData's 1-pole IIR (filter.cpp:36-48): f[i] = 0.9 f[i-1] + 0.1 raw[i]; restarts on an empty buffer."""
    out = imu.copy()
    restarts = set(restart_at) | {0}
    for i in range(len(imu)):
        if i in restarts:
            continue
        out[i, 1:7] = out[i - 1, 1:7] * (1 - 0.1) + imu[i, 1:7] * 0.1
    return out


def group_frames(image_rows: np.ndarray):
    """rows `t id p(3) q(4)` sharing a timestamp form one detection frame (FBUS_EKF.m:155-164)."""
    t, groups, i = [], [], 0
    while i < len(image_rows):
        j = i + 1
        while j < len(image_rows) and image_rows[j, 0] == image_rows[i, 0]:
            j += 1
        t.append(image_rows[i, 0])
        groups.append(image_rows[i:j, 1:9])
        i = j
    return np.array(t), groups


def frames_to_soa(t: np.ndarray, groups, batch: int = 1):
    """-> (ids int32 [W,m,B], pose float64 [W,m,7,B]) replicated over the batch"""
    m = max(len(g) for g in groups)
    W = len(groups)
    ids = -np.ones((W, m, batch), dtype=np.int32)
    pose = np.zeros((W, m, 7, batch))
    for w, g in enumerate(groups):
        for s, row in enumerate(g):
            ids[w, s, :] = int(row[0])
            pose[w, s, :, :] = row[1:8, None]
    return ids, pose


def window_offsets(t_imu: np.ndarray, t_frames: np.ndarray, start: int) -> np.ndarray:
    """win_off[w+1] = first IMU index with t > t_frames[w] (samples the reference would erase, filter.cpp:493-520)."""
    off = np.searchsorted(t_imu, t_frames, side="right")
    off = np.maximum(off, start)
    return np.concatenate([[start], off]).astype(np.uint32)


def replay_log(imu: np.ndarray, image_rows: np.ndarray, cfg=None, n_init: int = 500, use_iir: bool = False, batch: int = 1,
               device: int = 0, chunk: int | None = None):
    """Replays one recorded log (imu rows `t a(3) g(3)`, image rows `t id p q`) on the GPU.
    Returns dict(rows [W,17] of filter 0 in the data/fusion.txt layout, state, filter)."""
    if use_iir:
        imu = iir_prefilter(imu, restart_at=(n_init,))
    f = BatchFilter(cfg, batch=batch, device=device)
    t_imu = np.ascontiguousarray(imu[:, 0])
    data = np.ascontiguousarray(np.repeat(imu[:, 1:7, None], batch, axis=2))
    stream = capi.make_imu_stream(t_imu, data, batch)
    f.InitGravityAndGyrobias(stream, 0, n_init)
    t_frames, groups = group_frames(image_rows)
    ids, pose = frames_to_soa(t_frames, groups, batch)
    det = capi.make_det_frames(t_frames, ids, pose, batch, ids.shape[1])
    off = window_offsets(t_imu, t_frames, n_init)
    W = len(t_frames)
    chunk = chunk or W
    traces = []
    for w0 in range(0, W, chunk):
        w1 = min(W, w0 + chunk)
        traces.append(f.StepWindows(stream, det, off, w0, w1, trace=True))
    trace = np.concatenate(traces, axis=0)
    return {"rows": trace[:, :, 0].copy(), "trace": trace, "state": f.GetState(), "filter": f, "win_off": off}
