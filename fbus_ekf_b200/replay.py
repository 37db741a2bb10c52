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


def iir_prefilter(imu: np.ndarray, restart_at=(), device: int = 0) -> np.ndarray:
    """FILTER::SetImuData's 1-pole IIR (filter.cpp:36-48) on log rows `t a(3) g(3)`: f[i] = 0.9 f[i-1] + 0.1 raw[i], restarted
    at row 0 and at every row of `restart_at` (where the live buffer was empty).  Runs on the GPU (fbus_iir_prefilter)."""
    n = len(imu)
    out = np.array(imu, dtype=np.float64, copy=True)
    if n == 0:
        return out
    f = BatchFilter(None, batch=1, device=device)
    try:
        data = np.ascontiguousarray(out[:, 1:7, None])
        stream = capi.make_imu_stream(np.ascontiguousarray(out[:, 0]), data, 1)
        cuts = sorted({0, n} | {int(r) for r in restart_at if 0 < int(r) < n})
        for a, b in zip(cuts[:-1], cuts[1:]):
            out[a:b, 1:7] = f.IirPrefilter(stream, a, b - a)[:, :, 0]
    finally:
        f.close()
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


def solve_image_rows(corner_rows: np.ndarray, cfg=None, device: int = 0, gn_iters: int = 0) -> np.ndarray:
    """water-mode corners.txt rows (`t id` + 16 undistorted normalised stereo coordinates, stored as float32 like the
    reference's cv::Point2f) -> image.txt rows `t id p(3) q(4)` through the GPU refractive solve (R1 + R2, optionally the
    Gauss-Newton refinement); markers the solve rejects (vision.cpp:600-609: a corner farther than makrer_dect_dist_thres)
    produce no row, as in the reference."""
    corners = np.ascontiguousarray(corner_rows[:, 2:18].T.astype(np.float32))
    f = BatchFilter(cfg, batch=1, device=device)
    try:
        pose, _, valid = f.RefractSolveGN(corners, gn_iters) if gn_iters > 0 else f.RefractSolve(corners)
    finally:
        f.close()
    rows = np.concatenate([corner_rows[:, 0:2], pose.T], axis=1)
    return rows[np.asarray(valid).astype(bool)]


def recorded_frames(t_imu: np.ndarray, t_frames: np.ndarray, groups, cfg, start: int) -> np.ndarray:
    """Which detection frames the reference writes a data/fusion.txt row for: FILTER::FilterThreadFunction records only after
    ResetSystemState / BatchImuProcessing / ObservationUpdate ran (filter.cpp:229-248); the frame that initialises the pose and
    every frame before it `continue` past the recording block (filter.cpp:207-226).  Decided from the inputs alone, with the
    conditions of InitializePose (filter.cpp:295-358): a buffered IMU sample not later than the frame, the nearest marker
    within marker_max_dist, its id in the marker map."""
    known = {int(cfg.marker_id[i]) for i in range(cfg.n_markers)}
    rec = np.zeros(len(t_frames), dtype=bool)
    initialised = False
    for w, g in enumerate(groups):
        if len(g) == 0:
            continue  # the filter thread is not woken (vision.cpp:136-140)
        if initialised:
            rec[w] = True
            continue
        n_before = int(np.searchsorted(t_imu, t_frames[w], side="right")) - start
        dist = np.sqrt((g[:, 1:4] ** 2).sum(axis=1))
        near, md = 0, 10.0
        for c, d in enumerate(dist):  # first strict minimum below 10 (filter.cpp:331-341)
            if d < md:
                md, near = d, c
        if n_before > 0 and not md > cfg.marker_max_dist and int(g[near, 0]) in known:
            initialised = True
    return rec


IMU_BUFFER_MAX_SIZE = 2000  # filter.hpp:25
IMU_BUFFER_DROP = 500        # filter.cpp:52


def buffer_cap_keep(t_imu: np.ndarray, t_frames: np.ndarray, start: int, cap: int = IMU_BUFFER_MAX_SIZE,
                    drop: int = IMU_BUFFER_DROP) -> np.ndarray:
    """FILTER::SetImuData's bounded buffer (filter.cpp:50-54) in the deterministic replay: samples are pushed one by one,
    a push that makes the buffer longer than `cap` erases its oldest `drop` entries, and every detection frame erases what
    it consumed (all samples with t <= t_frame, filter.cpp:493-520).  Returns the boolean mask of the IMU rows (from
    `start` on; earlier rows belong to the gravity initialisation and are kept) that are still buffered when a frame
    consumes them.  With the bundled logs (25 Hz frames, 1 kHz IMU) nothing is ever dropped; a detection gap longer than
    `cap` samples loses the oldest part of the gap exactly as the live system does."""
    n = len(t_imu)
    keep = np.ones(n, dtype=bool)
    bounds = np.searchsorted(t_imu, t_frames, side="right")
    lo = start
    for hi in list(bounds) + [n]:
        hi = max(int(hi), lo)
        cnt = hi - lo  # pushes between two consumptions
        if cnt > cap:
            # size after k pushes: k while k <= cap, then cap+1-drop + (k-cap-1) % drop ; the survivors are the newest ones
            left = cap + 1 - drop + (cnt - cap - 1) % drop
            keep[lo:hi - left] = False
        lo = hi
    return keep


def replay_log(imu: np.ndarray, image_rows: np.ndarray, cfg=None, n_init: int = 500, use_iir: bool = False, batch: int = 1,
               device: int = 0, chunk: int | None = None, buffer_cap: int | None = None):
    """Replays one recorded log (imu rows `t a(3) g(3)`, image rows `t id p q`) on the GPU.
    Returns dict(rows [W,17] of filter 0 in the data/fusion.txt layout, state, filter).
    `buffer_cap` (e.g. IMU_BUFFER_MAX_SIZE) applies the live system's bounded IMU buffer to detection gaps."""
    if use_iir:
        imu = iir_prefilter(imu, restart_at=(n_init,), device=device)
    if buffer_cap:
        imu = imu[buffer_cap_keep(imu[:, 0], group_frames(image_rows)[0], n_init, buffer_cap)]
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
    return {"rows": trace[:, :, 0].copy(), "trace": trace, "state": f.GetState(), "filter": f, "win_off": off,
            "recorded": recorded_frames(t_imu, t_frames, groups, f.cfg, n_init)}


def main(argv=None) -> int:
    """python -m fbus_ekf_b200.replay DATASET_DIR [--water] [--iir] [--out fusion.txt]

    Replays a recorded dataset directory (imu.txt + image.txt, or with --water imu.txt + corners.txt through the GPU
    refractive solve) and writes the filter output in the reference's data/fusion.txt format (filter.cpp:241-246)."""
    import argparse

    from . import logio
    ap = argparse.ArgumentParser(prog="python -m fbus_ekf_b200.replay", description="Replays a recorded dataset directory on the GPU and writes the filter output in the data/fusion.txt format.")
    ap.add_argument("dataset", help="directory with imu.txt and image.txt / corners.txt (matlab/dataset/*/dataset-NN layout)")
    ap.add_argument("--water", action="store_true", help="solve marker poses from corners.txt (refractive flat-port model) instead of reading image.txt")
    ap.add_argument("--gn-iters", type=int, default=0, help="Gauss-Newton refinement iterations of the refractive solve (--water)")
    ap.add_argument("--iir", action="store_true", help="apply SetImuData's 1-pole IIR to the raw IMU log (live behaviour)")
    ap.add_argument("--n-init", type=int, default=500, help="IMU rows of the gravity / gyro-bias initialisation (FBUS_EKF.m:118)")
    ap.add_argument("--buffer-cap", type=int, default=IMU_BUFFER_MAX_SIZE, help="bounded IMU buffer of the live system, 0 = off")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--out", default="fusion.txt")
    ap.add_argument("--crlf", action="store_true", help="CRLF line ends like the bundled logs")
    a = ap.parse_args(argv)
    logs = logio.read_dataset(a.dataset)
    if "imu" not in logs:
        ap.error(f"{a.dataset}: no imu.txt")
    if a.water:
        if "corners" not in logs or logs["corners"].shape[1] != 18:
            ap.error(f"{a.dataset}: --water needs a corners.txt with 16 stereo coordinates per row")
        image = solve_image_rows(logs["corners"], device=a.device, gn_iters=a.gn_iters)
    else:
        if "image" not in logs:
            ap.error(f"{a.dataset}: no image.txt")
        image = logs["image"]
    res = replay_log(logs["imu"], image, None, n_init=a.n_init, use_iir=a.iir, device=a.device, buffer_cap=a.buffer_cap or None)
    # one row per frame the reference records: not the frame that initialises the pose, nor the ones before it (filter.cpp:207-248)
    rows = res["rows"][res["recorded"]]
    st = res["filter"].GetState(with_cov=False)
    logio.write_fusion_log(a.out, rows, newline="\r\n" if a.crlf else "\n")
    print(f"{len(res['rows'])} frames, {len(rows)} recorded -> {a.out}; final t = {rows[-1, 0]:.6g} s, p = ({rows[-1, 1]:.6g}, {rows[-1, 2]:.6g}, {rows[-1, 3]:.6g}) m, "
          f"status bits seen: 0x{int(np.bitwise_or.reduce(st['status'])) if 'status' in st else 0:x}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
