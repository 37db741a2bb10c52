"""NumPy ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT) -- second, independent restatement.

Restates the reference hot path (C++/src/filter.cpp, C++/src/vision.cpp of
CASIA-RoboticFish/FBUS-EKF) with NumPy float64 for single filters.  It is written
independently of oracle/fbus_oracle.cpp (matrix form, LAPACK solve / eig instead of
hand-written LDLT / Jacobi) so the two can be cross-checked to ~1e-13; see
tests/test_oracle_cross.py.  Only tests/ may import this module.

Pinning: R1+R2 are pinned by the reference's bundled water/land logs (tests/golden);
F1-F6 are "parity unpinned" beyond cm level (see oracle/fbus_oracle.h).
All citations are relative to /root/reference.
"""
from __future__ import annotations

import dataclasses
import numpy as np

REF_M_PI = 3.1415926  # common.hpp:14


# ----------------------------------------------------------------------------- helpers
def skew(v):  # matrix_math.hpp:26-36
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def quat_left(q):  # matrix_math.hpp:38-62
    w, x, y, z = q
    return np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]], dtype=np.float64)


def quat_right(q):  # matrix_math.hpp:64-88
    w, x, y, z = q
    return np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]], dtype=np.float64)


def qmul(a, b):  # Hamilton product == left-product matrix applied to b
    return quat_left(a) @ np.asarray(b, dtype=np.float64)


def qconj(a):
    return np.array([a[0], -a[1], -a[2], -a[3]])


def q2R(q):  # Eigen toRotationMatrix
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def R2q(m):  # Eigen Quaterniond(Matrix3d), w,x,y,z, not normalised
    m = np.asarray(m, dtype=np.float64)
    q = np.zeros(4)
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[0] = 0.5 * t
        t = 0.5 / t
        q[1] = (m[2, 1] - m[1, 2]) * t
        q[2] = (m[0, 2] - m[2, 0]) * t
        q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3
        k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t
        q[1 + j] = (m[j, i] + m[i, j]) * t
        q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def rodrigues(angle, u):  # Eigen AngleAxisd::matrix
    u = np.asarray(u, dtype=np.float64)
    K = skew(u)
    return np.cos(angle) * np.eye(3) + np.sin(angle) * K + (1 - np.cos(angle)) * np.outer(u, u)


# ----------------------------------------------------------------------------- configuration
@dataclasses.dataclass
class Config:
    tsc_left: np.ndarray
    tsc_right: np.ndarray
    accel_n_cov: float = 1e-3
    gyro_n_cov: float = 1e-4
    accel_b_cov: float = 1e-3
    gyro_b_cov: float = 1e-4
    pos_n_cov: float = 1e-3
    quat_n_cov: float = 1e-3
    marker_max_dist: float = 2.0
    marker_switch_thres: float = 0.5
    p0_diag: tuple = (1e-4, 1e-2, 1e-4, 1e-2, 1e-2, 100.0)
    reset_gap: float = 0.1
    n_air: float = 1.00
    n_glass: float = 1.49
    n_water: float = 1.32
    d_air: float = 0.002
    d_glass: float = 0.02
    normal: tuple = (0.0, 0.0, 1.0)
    marker_dect_dist_thres: float = 2.0
    markers: dict = dataclasses.field(default_factory=dict)  # id -> (pos(3), rot(3x3))


TSC_LEFT_1 = np.array([[-0.999862, 0.015685, -0.00548, 0.059967],
                       [-0.015639, -0.999843, -0.00827, 0.000127837],
                       [-0.005609, -0.008183, 0.999951, -0.002],
                       [0, 0, 0, 1.0]])
TSC_RIGHT_1 = np.array([[-0.999826, 0.00929485, -0.0161445, -0.0601272],
                        [-0.00937869, -0.999942, 0.00514829, 0.000124714],
                        [-0.0160959, 0.00529897, 0.999857, -0.002],
                        [0, 0, 0, 1.0]])


def default_markers():
    """C++/config/markersetup.yml"""
    I = np.eye(3)
    Rx90 = np.array([[1, 0, 0], [0, 0, -1], [0, 1, 0.0]])
    Rx180 = np.array([[1, 0, 0], [0, -1, 0], [0, 0, -1.0]])
    m = {0: ((0, 0, 0), I), 1: ((0, 0.61, 0.285), Rx90), 2: ((0, 0.61, 1.185), Rx90),
         3: ((0, 0.61, 2.085), Rx90), 4: ((0, 0.61, 2.985), Rx90), 5: ((0, 0.265, 4.12), Rx180),
         6: ((0, -0.635, 4.12), Rx180), 7: ((0, -1.535, 4.12), Rx180), 8: ((0, -2.435, 4.12), Rx180),
         16: ((0, -2.7, 0), I), 17: ((0, -1.8, 0), I), 18: ((0, -0.9, 0), I)}
    return {k: (np.array(p, dtype=np.float64), np.array(r, dtype=np.float64)) for k, (p, r) in m.items()}


def default_config():
    """camerainfo1.yml + paramconfig.yml + markersetup.yml (the bundled logs' configuration)."""
    return Config(tsc_left=TSC_LEFT_1.copy(), tsc_right=TSC_RIGHT_1.copy(), markers=default_markers())


class Consts:
    def __init__(self, cfg: Config):
        T = np.diag([-1.0, -1.0, 1.0, 1.0]) @ cfg.tsc_left  # filter.hpp:67-69
        self.R_IL = T[:3, :3].copy()
        self.Q_IL = R2q(self.R_IL)  # not normalised, filter.cpp:370
        self.P_IL = -self.R_IL.T @ T[:3, 3]
        self.markers = {k: (p.copy(), R2q(r)) for k, (p, r) in cfg.markers.items()}
        self.Qbar = np.zeros(18)
        self.Qbar[3:6] = cfg.accel_n_cov
        self.Qbar[6:9] = cfg.gyro_n_cov
        self.Qbar[9:12] = cfg.accel_b_cov
        self.Qbar[12:15] = cfg.gyro_b_cov
        self.Rn = np.diag([cfg.pos_n_cov] * 3 + [cfg.quat_n_cov] * 4)
        self.cfg = cfg
        # vision.cpp:476-481 (raw T_SC)
        R_IL_raw, R_IR_raw = cfg.tsc_left[:3, :3], cfg.tsc_right[:3, :3]
        self.R_RL = R_IL_raw @ R_IR_raw.T
        self.P_LR = cfg.tsc_left[:3, 3] - self.R_RL @ cfg.tsc_right[:3, 3]


class Filter:
    def __init__(self, k: Consts):
        self.k = k
        self.t = 0.0
        self.q = np.array([1.0, 0, 0, 0])
        self.R = np.zeros((3, 3))
        self.p = np.zeros(3)
        self.v = np.zeros(3)
        self.ba = np.zeros(3)
        self.bg = np.zeros(3)
        self.g = np.zeros(3)
        self.pv = np.zeros(3)
        self.qv = np.array([1.0, 0, 0, 0])
        self.P = np.diag(np.repeat(np.array(k.cfg.p0_diag, dtype=np.float64), 3))
        self.prev_marker_id = 0
        self.initialised = False
        self.n_resets = 0

    # filter.cpp:256-285
    def init_gravity_gyrobias(self, imu):
        self.bg = imu[:, 4:7].sum(axis=0) / len(imu)
        self.g = np.array([0, 0, -np.linalg.norm(imu[:, 1:4].sum(axis=0) / len(imu))])

    def _nearest(self, dets):
        md, idx = 10.0, 0
        for c, d in enumerate(dets):
            dist = np.linalg.norm(d[1:4])
            if dist < md:
                md, idx = dist, c
        return idx, md

    # filter.cpp:291-399; dets = list of (id, px,py,pz, qw,qx,qy,qz)
    def init_pose(self, dets, t_det, n_imu_before=1):
        if n_imu_before <= 0 or len(dets) == 0:
            return False
        idx, md = self._nearest(dets)
        if md > self.k.cfg.marker_max_dist:
            return False
        d = dets[idx]
        if int(d[0]) not in self.k.markers:
            return False
        P_M, Q_M = self.k.markers[int(d[0])]
        self.t = t_det
        self.q = qmul(qmul(Q_M, qconj(d[4:8])), self.k.Q_IL)
        self.R = q2R(self.q)
        self.p = P_M - self.R @ self.k.P_IL - self.R @ self.k.R_IL.T @ d[1:4]
        self.g = np.array([9.8, 0, 0])
        self.initialised = True
        return True

    # filter.cpp:405-477
    def reset_state(self, dets, t_det):
        if len(dets) == 0:
            return
        idx, md = self._nearest(dets)
        if md > self.k.cfg.marker_max_dist:
            return
        d = dets[idx]
        if int(d[0]) not in self.k.markers:
            return
        P_M, Q_M = self.k.markers[int(d[0])]
        self.qv = qmul(qmul(Q_M, qconj(d[4:8])), self.k.Q_IL)
        R_IG = q2R(self.qv)
        self.pv = -R_IG @ self.k.R_IL.T @ d[1:4] + P_M - R_IG @ self.k.P_IL
        if t_det - self.t > self.k.cfg.reset_gap and self.initialised:
            self.t = t_det
            self.q = self.qv.copy()
            self.p = self.pv.copy()
            self.v = np.zeros(3)
            self.ba = np.zeros(3)
            self.bg = np.zeros(3)
            self.n_resets += 1

    # filter.cpp:588-616
    def update_covariance(self, dt, accel, gyro):
        w = gyro - self.bg
        a = accel - self.ba
        F = np.eye(18)
        F[0:3, 3:6] = np.eye(3) * dt
        F[3:6, 6:9] = -self.R @ skew(a) * dt
        F[3:6, 9:12] = -self.R * dt
        F[3:6, 15:18] = np.eye(3) * dt
        F[6:9, 6:9] = np.eye(3) - skew(w) * dt
        F[6:9, 12:15] = -np.eye(3) * dt
        P = F @ self.P @ F.T + np.diag(self.k.Qbar)
        self.P = (P + P.T) / 2.0

    # filter.cpp:533-582
    def update_nominal(self, dt, accel, gyro):
        w = gyro - self.bg
        wn = np.linalg.norm(w)
        R0 = q2R(self.q)
        if wn > 10e-5:
            ax = w / wn
            ah = wn * dt / 2
            qh = qmul(self.q, np.concatenate([[np.cos(ah / 2)], np.sin(ah / 2) * ax]))
            af = wn * dt
            qn = qmul(self.q, np.concatenate([[np.cos(af / 2)], np.sin(af / 2) * ax]))
        else:
            qh = qmul(self.q, np.concatenate([[1.0], 0.5 * dt * w / 2]))
            qn = qmul(self.q, np.concatenate([[1.0], 0.5 * dt * w]))
        qh = qh / np.linalg.norm(qh)
        self.q = qn / np.linalg.norm(qn)
        Rh = q2R(qh)
        self.R = q2R(self.q)
        a = accel - self.ba
        k1 = R0 @ a + self.g
        k2 = Rh @ a + self.g
        k3 = k2
        k4 = self.R @ a + self.g
        v0 = self.v.copy()
        self.v = v0 + dt / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        self.p = self.p + dt / 6 * (v0 + 2 * (v0 + k1 * dt / 2) + 2 * (v0 + k2 * dt / 2) + (v0 + k3 * dt / 2))

    # filter.cpp:483-531
    def batch_imu(self, imu_rows, t_end):
        """-> imuCnt, the number of leading rows the reference erases afterwards (filter.cpp:492-503,520)"""
        start = self.t
        cnt = 0
        for row in imu_rows:
            if row[0] < start:
                cnt += 1
                continue
            if row[0] > t_end:
                break
            cnt += 1
            dt = row[0] - self.t
            self.update_covariance(dt, row[1:4], row[4:7])
            self.update_nominal(dt, row[1:4], row[4:7])
            self.t = row[0]
        return cnt

    def measurement_model(self, d):
        """returns (hP, hQ, H) for detection d before sign disambiguation (filter.cpp:677-694)"""
        k = self.k
        P_M, Q_M = k.markers[int(d[0])]
        hP = k.R_IL @ self.R.T @ (P_M - self.p - self.R @ k.P_IL)
        hQ = qmul(qmul(k.Q_IL, qconj(self.q)), Q_M)
        H = np.zeros((7, 18))
        H[0:3, 0:3] = -k.R_IL @ self.R.T
        H[0:3, 6:9] = k.R_IL @ skew(self.R.T @ (P_M - self.p))
        L1 = np.zeros((4, 3))
        L1[1:4, :] = 0.5 * np.eye(3)
        L2 = np.diag([1.0, -1, -1, -1])
        H[3:7, 6:9] = quat_right(Q_M) @ quat_left(k.Q_IL) @ L2 @ quat_left(self.q) @ L1
        return hP, hQ, H

    # filter.cpp:622-739
    def observation_update(self, dets, joseph=False):
        if len(dets) == 0:
            return
        k = self.k
        md, idx, prev_dist, prev_idx = 10.0, 0, 0.0, 0
        for c, d in enumerate(dets):
            dist = np.linalg.norm(d[1:4])
            if dist < md:
                md, idx = dist, c
            if int(d[0]) == self.prev_marker_id:
                prev_dist, prev_idx = dist, c
        if abs(prev_dist - md) < k.cfg.marker_switch_thres and prev_dist != 0:
            idx = prev_idx
        d = dets[idx]
        if int(d[0]) not in k.markers:
            return
        self.prev_marker_id = int(d[0])
        hP, hQ, H = self.measurement_model(d)
        yP, yQ = d[1:4], d[4:8]
        if np.sum((yQ - hQ) ** 2) > np.sum((yQ + hQ) ** 2):
            hQ = -hQ
            H[3:7, 6:9] = -H[3:7, 6:9]
        S = H @ self.P @ H.T + k.Rn
        K = np.linalg.solve(S, H @ self.P).T
        r = np.concatenate([yP - hP, yQ - hQ])
        dx = K @ r
        self.p = self.p + dx[0:3]
        self.v = self.v + dx[3:6]
        th = dx[6:9]
        vn = np.linalg.norm(th)
        with np.errstate(invalid="ignore", divide="ignore"):
            dq = np.concatenate([[np.cos(vn / 2)], th / vn * np.sin(vn / 2)])
        q = qmul(self.q, dq)
        self.q = q / np.linalg.norm(q)  # R not refreshed
        self.ba = self.ba + dx[9:12]
        self.bg = self.bg + dx[12:15]
        self.g = self.g + dx[15:18]
        if joseph:
            IKH = np.eye(18) - K @ H
            P = IKH @ self.P @ IKH.T + K @ k.Rn @ K.T
        else:
            P = (np.eye(18) - K @ H) @ self.P
        self.P = (P + P.T) / 2.0


# ----------------------------------------------------------------------------- refraction (R1, R2)
def refraction_triangulation(k: Consts, c16, as_float32=True):
    """vision.cpp:488-608.  c16: 16 numbers (L xy x4, R xy x4), rounded to float32 like cv::Point2f unless
    as_float32=False. -> (4x3 corners, ok)"""
    c16 = np.asarray(c16, dtype=np.float32).astype(np.float64) if as_float32 else np.asarray(c16, dtype=np.float64)
    cfg = k.cfg
    nv = np.array(cfg.normal, dtype=np.float64)
    out = np.zeros((4, 3))
    ok = True

    def refract(r, n_from, n_to, glass_rule):
        alpha = n_from / n_to
        v = r @ nv
        root = np.sqrt(1 - alpha * alpha * (1 - v * v))
        beta = (root - alpha * v) if glass_rule else (alpha * v - root)
        return alpha * r + beta * nv, v

    for i in range(4):
        lp = np.array([c16[2 * i], c16[2 * i + 1], 1.0])
        rp = np.array([c16[8 + 2 * i], c16[8 + 2 * i + 1], 1.0])
        rays = []
        for pt in (lp, rp):
            r0 = pt / np.linalg.norm(pt)
            r1, v0 = refract(r0, cfg.n_air, cfg.n_glass, cfg.n_air < cfg.n_glass)
            r2, v1 = refract(r1, cfg.n_glass, cfg.n_water, cfg.n_glass > cfg.n_water)
            P0 = (cfg.d_air / v0) * r0
            P1 = P0 + (cfg.d_glass / v1) * r1
            rays.append((r2, P1))
        r2L, P1L = rays[0]
        r2R = k.R_RL @ rays[1][0]
        P1R = k.P_LR + k.R_RL @ rays[1][1]
        c = np.cross(r2L, r2R)
        dd = P1R - P1L
        X1 = np.column_stack([c, dd, r2R])
        X2 = np.column_stack([c, r2L, dd])
        X3 = np.column_stack([c, r2L, r2R])
        t1 = np.linalg.det(X1) / np.linalg.det(X3)
        t2 = -np.linalg.det(X2) / np.linalg.det(X3)
        P = 0.5 * (P1L + t1 * r2L + P1R + t2 * r2R)
        out[i] = np.array([-P[0], -P[1], P[2]])
        if np.linalg.norm(P) > cfg.marker_dect_dist_thres:
            ok = False
            break
    return out, ok


def normal_triangulation(cfg: Config, c16, dect_thres=2.0):
    """vision.cpp:395-466 (land mode): homogeneous DLT, smallest right singular vector by LAPACK SVD"""
    c16 = np.asarray(c16, dtype=np.float32).astype(np.float64)
    T0 = np.hstack([np.eye(3), np.zeros((3, 1))])
    R = cfg.tsc_right[:3, :3] @ cfg.tsc_left[:3, :3].T
    TLR = np.hstack([R, (cfg.tsc_left[:3, 3] - R @ cfg.tsc_right[:3, 3])[:, None]])
    out, ok = np.zeros((4, 3)), True
    for i in range(4):
        lp = np.array([c16[2 * i], c16[2 * i + 1], 1.0])
        rp = np.array([c16[8 + 2 * i], c16[8 + 2 * i + 1], 1.0])
        A = np.vstack([skew(lp) @ T0, skew(rp) @ TLR])
        P = np.linalg.svd(A)[2][3]
        if P[3] == 0:
            continue
        Pn = np.diag([-1.0, -1.0, 1.0]) @ (P[:3] / P[3])
        out[i] = (-1.0 if Pn[2] < 0 else 1.0) * Pn
        if np.linalg.norm(Pn) > dect_thres:
            ok = False
            break
    return out, ok


def compute_marker_pose(C):
    """vision.cpp:635-759.  C: 4x3 corners in the flipped left-camera frame -> (p, q, R)"""
    C = np.asarray(C, dtype=np.float64)
    vs = [C[1] - C[0], C[2] - C[0], C[3] - C[0], C[2] - C[1], C[3] - C[1], C[3] - C[2]]
    M = sum(np.outer(v, v) for v in vs)
    w, V = np.linalg.eig(M)  # general real eigen-solver, as EigenSolver<Matrix3d>
    w, V = w.real, V.real
    if w[0] < w[1]:
        col = 0 if w[0] < w[2] else 2
    else:
        col = 1 if w[1] < w[2] else 2
    Z = V[:, col] / np.linalg.norm(V[:, col])
    sg = lambda x: -1.0 if x < 0 else 1.0
    if Z[2] > 0.1:
        Z = -Z
    elif Z[2] < -0.1:
        pass
    else:
        Z = -sg(C[0][0]) * sg(Z[0]) * Z
    D = 0.25 * Z @ (C[0] + C[1] + C[2] + C[3])
    Pp = [c - (Z @ c - D) * Z for c in C]
    V12, V14 = Pp[1] - Pp[0], Pp[3] - Pp[0]
    m = V12 / np.linalg.norm(V12) + V14 / np.linalg.norm(V14)
    X = rodrigues(-REF_M_PI / 4, Z) @ m / np.linalg.norm(m)
    Y = np.cross(Z, X)
    R = np.column_stack([X, Y, Z])
    return Pp[0], R2q(R), R


# ----------------------------------------------------------------------------- replay driver
def iir_prefilter(imu, restart_at=()):
    """FILTER::SetImuData 1-pole IIR (filter.cpp:36-48); restarts where the live buffer was empty."""
    out = imu.copy()
    restarts = set(restart_at) | {0}
    for i in range(len(imu)):
        if i in restarts:
            continue
        out[i, 1:7] = out[i - 1, 1:7] * (1 - 0.1) + imu[i, 1:7] * 0.1
    return out


def group_frames(image_rows):
    """rows 't id p q' sharing a timestamp form one frame (FBUS_EKF.m:155-164)"""
    frames, i = [], 0
    while i < len(image_rows):
        j = i + 1
        while j < len(image_rows) and image_rows[j, 0] == image_rows[i, 0]:
            j += 1
        frames.append((image_rows[i, 0], [image_rows[r, 1:9] for r in range(i, j)]))
        i = j
    return frames


def window_offsets(t_imu, t_frames, start):
    """CSR offsets: window w consumes samples [off[w], off[w+1]) = buffered samples with t <= t_frames[w]"""
    off = np.searchsorted(t_imu, np.asarray(t_frames), side="right")
    off = np.maximum(off, start)
    return np.concatenate([[start], off]).astype(np.uint32)


def replay(cfg: Config, imu, image_rows, n_init=500, use_iir=False, joseph=False, trace_cov=False):
    """Deterministic replay of FILTER::FilterThreadFunction, modelled on FBUS_EKF.m:116-197.
    Returns dict with per-frame rows [t p q v ba bg] and optionally P."""
    k = Consts(cfg)
    f = Filter(k)
    if use_iir:
        imu = iir_prefilter(imu, restart_at=(n_init,))
    f.init_gravity_gyrobias(imu[:n_init])
    frames = group_frames(image_rows)
    off = window_offsets(imu[:, 0], [fr[0] for fr in frames], n_init)
    rows, covs = [], []
    cursor = int(off[0])
    for w, (t_det, dets) in enumerate(frames):
        hi = int(off[w + 1])
        if not f.initialised:
            later = np.nonzero(imu[cursor:hi, 0] > t_det)[0]
            cnt = int(later[0]) if len(later) else hi - cursor  # imuCnt: leading rows not later than the frame (filter.cpp:299-305)
            if f.init_pose(dets, t_det, cnt):
                cursor += cnt  # only those are erased (filter.cpp:390)
        else:
            f.reset_state(dets, t_det)
            cursor += f.batch_imu(imu[cursor:hi], t_det)
            f.observation_update(dets, joseph=joseph)
        rows.append(np.concatenate([[f.t], f.p, f.q, f.v, f.ba, f.bg]))
        if trace_cov:
            covs.append(f.P.copy())
    return {"rows": np.array(rows), "P": np.array(covs) if trace_cov else None, "filter": f, "win_off": off,
            "frames": frames, "imu": imu}


# ----------------------------------------------------------------------------- R3: Gauss-Newton refinement
# NOT in the reference ("parity unpinned").  Independent restatement of the formulation
# the CUDA kernel implements: flat-port forward projection by a bracketing root finder (the kernel uses Newton), Jacobian
# by the implicit-function theorem, 6x6 normal equations solved by LAPACK (the kernel uses Cholesky).
def project_refr(cfg: Config, X):
    """X: point in a camera's own frame -> (uv[2], J[2x3])"""
    from scipy.optimize import brentq
    d0, d1 = cfg.d_air, cfg.d_glass
    k1, k2 = cfg.n_air / cfg.n_glass, cfg.n_air / cfg.n_water
    rho = np.hypot(X[0], X[1])
    Zw = X[2] - d0 - d1
    t = lambda s: s / np.sqrt(1 - s * s)
    dt = lambda s: (1 - s * s) ** -1.5
    if rho < 1e-12:
        den = 1.0 / (d0 + k1 * d1 + k2 * Zw)
        J = np.array([[den, 0, -X[0] * den * den * k2], [0, den, -X[1] * den * den * k2]])
        return np.array([X[0] * den, X[1] * den]), J
    f = lambda s: d0 * t(s) + d1 * t(k1 * s) + Zw * t(k2 * s) - rho
    s0 = brentq(f, 0.0, 1 - 1e-14, xtol=1e-17, rtol=8.9e-16, maxiter=200)
    xh = np.array([X[0], X[1]]) / rho
    tau = t(s0)
    gs = d0 * dt(s0) + d1 * k1 * dt(k1 * s0) + Zw * k2 * dt(k2 * s0)
    Jxy = (tau / rho) * (np.eye(2) - np.outer(xh, xh)) + (dt(s0) / gs) * np.outer(xh, xh)
    Jz = -dt(s0) * t(k2 * s0) / gs * xh
    return tau * xh, np.column_stack([Jxy, Jz])


def gn_residuals(k: Consts, c16, Rm, p, size=0.28):
    """-> (r[16] ordered Lx0,Ly0..Lx3,Ly3,Rx0..Ry3, J[16x6] w.r.t. (dp, dphi) with R <- R Exp(dphi))"""
    cfg = k.cfg
    F = np.diag([-1.0, -1.0, 1.0])
    Rinv = np.linalg.inv(k.R_RL)
    cm = np.array([[0, 0, 0], [size, 0, 0], [size, size, 0], [0, size, 0]], dtype=np.float64)
    r = np.zeros(16)
    J = np.zeros((16, 6))
    for i in range(4):
        XL = F @ (p + Rm @ cm[i])
        D = np.column_stack([F, -F @ Rm @ skew(cm[i])])
        for cam in range(2):
            if cam == 0:
                X, DX = XL, D
            else:
                X, DX = Rinv @ (XL - k.P_LR), Rinv @ D
            uv, Jp = project_refr(cfg, X)
            r[cam * 8 + 2 * i:cam * 8 + 2 * i + 2] = uv - c16[cam * 8 + 2 * i:cam * 8 + 2 * i + 2]
            J[cam * 8 + 2 * i:cam * 8 + 2 * i + 2, :] = Jp @ DX
    return r, J


def so3_exp(phi):
    th = np.linalg.norm(phi)
    K = skew(phi)
    if th < 1e-8:
        return np.eye(3) + K + 0.5 * K @ K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def gn_refine(k: Consts, c16, Rm, p, iters=5, size=0.28):
    c16 = np.asarray(c16, dtype=np.float64)
    Rm, p = np.array(Rm, dtype=np.float64), np.array(p, dtype=np.float64)
    for _ in range(iters):
        r, J = gn_residuals(k, c16, Rm, p, size)
        d = np.linalg.solve(J.T @ J, -J.T @ r)
        p = p + d[:3]
        Rm = Rm @ so3_exp(d[3:])
    r, _ = gn_residuals(k, c16, Rm, p, size)
    return Rm, p, float(r @ r)


def refract_solve_gn(k: Consts, c16, iters=5, size=0.28):
    """closed form (R1+R2) then GN; c16 are the coordinates as given (float32- or float64-valued)"""
    C, ok = refraction_triangulation(k, c16, as_float32=False)
    if not ok:
        return None
    p, q, _ = compute_marker_pose(C)
    qn = q / np.linalg.norm(q)
    Rm, p, cost = gn_refine(k, c16, q2R(qn), p, iters, size)
    return p, R2q(Rm), cost
