"""NumPy ORACLE of the reference's MATLAB implementation (TEST INFRASTRUCTURE, NOT PRODUCT).

Restates matlab/*.m of CASIA-RoboticFish/FBUS-EKF function by function: the same filter as C++/src/filter.cpp with
different numerics.  It is the checker of the product's FBUS_FLAG_MATLAB mode (tests/test_gpu_matlab_mode.py).
"parity unpinned": there is no MATLAB / Octave in this image and the script's output is not among the bundled logs, so this
restatement is a reading of the .m files, cross-checked only against the C++-semantics oracle where the two coincide.

One choice MATLAB leaves open is made explicit: rotmat_to_quaternion.m takes the eigenvector V(:,4) of a symmetric 4x4 matrix,
whose SIGN is whatever LAPACK returns.  Every filter output except the sign of State.quaternion is invariant under that sign
(the update's gain and covariance contain H only as H' inv(S) H, and the quaternion rows of the residual are zeroed,
MeasureUpdate.m:88); here and in the product the eigenvector is taken with a non-negative scalar part.

All citations are relative to /root/reference/matlab.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import expm


# ----------------------------------------------------------------------------- helpers (one .m file each)
def vector_to_crossmat(v):  # vector_to_crossmat.m
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]], dtype=np.float64)


def quaternion_add(p, q):  # quaternion_add.m:24-27 (Hamilton product)
    return np.array([p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3],
                     p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2],
                     p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1],
                     p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0]])


def quaternion_conjugate(q):
    return np.array([q[0], -q[1], -q[2], -q[3]])


def quaternion_normalize(q):
    return q / np.linalg.norm(q)


def axisangle_to_quaternion(axis, angle):  # axisangle_to_quaternion.m:22-28 (0/0 = NaN for a zero axis, as MATLAB)
    with np.errstate(invalid="ignore", divide="ignore"):
        axis = np.asarray(axis, dtype=np.float64) / np.linalg.norm(axis)
    return np.array([np.cos(angle / 2), axis[0] * np.sin(angle / 2), axis[1] * np.sin(angle / 2), axis[2] * np.sin(angle / 2)])


def quaternion_to_rotmat(q):  # quaternion_to_rotmat.m:24-32 (the w^2+x^2-y^2-z^2 form; NOT Eigen's for a non-unit q)
    w, x, y, z = q
    return np.array([[w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z]])


def rotmat_to_quaternion(R):  # rotmat_to_quaternion.m:23-43
    R = np.asarray(R, dtype=np.float64).T
    K = np.zeros((4, 4))
    K[0, 0] = R[0, 0] - R[1, 1] - R[2, 2]
    K[0, 1] = K[1, 0] = R[1, 0] + R[0, 1]
    K[0, 2] = K[2, 0] = R[2, 0] + R[0, 2]
    K[0, 3] = K[3, 0] = R[1, 2] - R[2, 1]
    K[1, 1] = R[1, 1] - R[0, 0] - R[2, 2]
    K[1, 2] = K[2, 1] = R[2, 1] + R[1, 2]
    K[1, 3] = K[3, 1] = R[2, 0] - R[0, 2]
    K[2, 2] = R[2, 2] - R[0, 0] - R[1, 1]
    K[2, 3] = K[3, 2] = R[0, 1] - R[1, 0]
    K[3, 3] = R[0, 0] + R[1, 1] + R[2, 2]
    K /= 3.0
    _, V = np.linalg.eigh(K)  # ascending eigenvalues, as MATLAB's eig of a symmetric matrix
    v = V[:, 3]
    q = np.array([v[3], v[0], v[1], v[2]])
    # sign convention (see the module docstring): scalar part >= 0, first non-zero component positive otherwise
    for c in q:
        if abs(c) > 1e-12:
            if c < 0:
                q = -q
            break
    return q


class Config:
    """FBUS_EKF.m:28-41 / 83-112 + config/camerainfo.yml (same values as C++/config/camerainfo1.yml) + GetMarkerMap.m"""

    def __init__(self, tsc_left, markers):
        self.T_IL = np.diag([-1.0, -1.0, 1.0, 1.0]) @ np.asarray(tsc_left, dtype=np.float64).reshape(4, 4)  # FBUS_EKF.m:64
        self.markers = markers  # id -> (position(3), rotation(3x3))
        self.P0 = np.diag(np.repeat([1e-4, 0.1, 1e-4, 1e-3, 1e-3, 100.0], 3))          # FBUS_EKF.m:86-98
        self.systemNoise = np.diag(np.repeat([1e-3, 1e-4, 1e-3, 1e-4], 3))              # FBUS_EKF.m:101-105
        self.measureNoise = np.diag([0.01] * 3 + [0.01] * 4)                            # FBUS_EKF.m:108-110
        self.reset_gap = 0.1                                                             # FBUS_EKF.m:168


class State:
    def __init__(self, cfg: Config):
        self.position = np.zeros(3)
        self.quaternion = np.zeros(4)
        self.velocity = np.zeros(3)
        self.accelBias = np.zeros(3)
        self.gyroBias = np.zeros(3)
        self.gravity = np.zeros(3)
        self.rotateMat = np.zeros((3, 3))
        self.covariance = cfg.P0.copy()
        self.systemNoise = cfg.systemNoise
        self.measureNoise = cfg.measureNoise


def _nearest(visionMeas):
    """the loop of every .m file that takes a measurement (e.g. MeasureUpdate.m:51-60); visionMeas: [n][8] rows id p(3) q(4)"""
    min_id, min_distance = -1, 10.0
    for i, row in enumerate(visionMeas):
        d = np.linalg.norm(row[1:4])
        if d < min_distance:
            min_distance, min_id = d, i
    return min_id


def _extrinsics(cfg):
    R_IL = cfg.T_IL[:3, :3]
    return R_IL, -R_IL.T @ cfg.T_IL[:3, 3], rotmat_to_quaternion(R_IL)


def InitGravityAndGyrobias(imuData):  # InitGravityAndGyrobias.m
    mean = imuData.sum(axis=0) / len(imuData)
    return -np.array([0, 0, np.linalg.norm(mean[1:4])]), mean[4:7].copy()


def _vision_pose(cfg, visionMeas):
    i = _nearest(visionMeas)
    posMeas, quatMeas = visionMeas[i, 1:4], visionMeas[i, 4:8]
    markerPos, markerRot = cfg.markers[int(visionMeas[i, 0])]
    markerQuat = rotmat_to_quaternion(markerRot)
    R_IL, P_IL, Q_IL = _extrinsics(cfg)
    Q_IG = quaternion_add(quaternion_add(markerQuat, quaternion_conjugate(quatMeas)), Q_IL)
    return Q_IG, posMeas, markerPos, R_IL, P_IL


def InitPositionAndQuaternion(S: State, visionMeas, cfg):  # InitPositionAndQuaternion.m
    Q_IG, P_ML, P_MG, R_IL, P_IL = _vision_pose(cfg, visionMeas)
    R_IG = quaternion_to_rotmat(Q_IG)
    S.quaternion = Q_IG
    S.position = -R_IG @ R_IL.T @ P_ML + P_MG - R_IG @ P_IL
    S.rotateMat = quaternion_to_rotmat(S.quaternion)
    S.gravity = np.array([9.8, 0.0, 0.0])


def ResetState(S: State, visionMeas, cfg):  # ResetState.m: zeroes velocity and accelBias only
    Q_IG, P_ML, P_MG, R_IL, P_IL = _vision_pose(cfg, visionMeas)
    R_IG = quaternion_to_rotmat(Q_IG)
    S.position = -R_IG @ R_IL.T @ P_ML + P_MG - R_IG @ P_IL
    S.quaternion = Q_IG
    S.rotateMat = R_IG
    S.velocity = np.zeros(3)
    S.accelBias = np.zeros(3)


def ComputeVisionOnlyResults(visionMeas, cfg):  # ComputeVisionOnlyResults.m (normalises Q_IG)
    Q_IG, P_ML, P_MG, R_IL, P_IL = _vision_pose(cfg, visionMeas)
    Q_IG = quaternion_normalize(Q_IG)
    R_IG = quaternion_to_rotmat(Q_IG)
    return -R_IG @ R_IL.T @ P_ML + P_MG - R_IG @ P_IL, Q_IG


def ImuUpdate(S: State, accel, gyro, dt):  # ImuUpdate.m:37-81
    a = accel - S.accelBias
    w = gyro - S.gyroBias
    dtheta = np.linalg.norm(w * dt)
    qT = quaternion_add(S.quaternion, axisangle_to_quaternion(w, dtheta))
    qHalfT = quaternion_add(S.quaternion, axisangle_to_quaternion(w, dtheta / 2))
    R0 = S.rotateMat
    RHalfT = quaternion_to_rotmat(qHalfT)
    RT = quaternion_to_rotmat(qT)
    kv1 = R0 @ a + S.gravity
    kv2 = RHalfT @ a + S.gravity
    kv3 = kv2
    kv4 = RT @ a + S.gravity
    v = S.velocity + dt / 6 * (kv1 + 2 * kv2 + 2 * kv3 + kv4)
    kp1 = S.velocity
    kp2 = S.velocity + kv1 * dt / 2
    kp3 = S.velocity + kv2 * dt / 2
    kp4 = S.velocity + kv3 * dt / 2
    p = S.position + dt / 6 * (kp1 + 2 * kp2 + 2 * kp3 + kp4)
    Fx = np.eye(18)
    Fx[0:3, 3:6] = np.eye(3) * dt
    Fx[3:6, 6:9] = -S.rotateMat @ vector_to_crossmat(a) * dt
    Fx[3:6, 9:12] = -S.rotateMat * dt
    Fx[3:6, 15:18] = np.eye(3) * dt
    Fx[6:9, 6:9] = expm(-vector_to_crossmat(w) * dt)
    Fx[6:9, 12:15] = -np.eye(3) * dt
    Fi = np.vstack([np.zeros((3, 12)), np.eye(12), np.zeros((3, 12))])
    P = Fx @ S.covariance @ Fx.T + Fi @ S.systemNoise @ Fi.T
    S.quaternion = quaternion_normalize(qT)
    S.rotateMat = RT
    S.velocity = v
    S.position = p
    S.covariance = (P + P.T) / 2


def quaternion_left_product_matrix(q):
    w, x, y, z = q
    return np.array([[w, -x, -y, -z], [x, w, -z, y], [y, z, w, -x], [z, -y, x, w]], dtype=np.float64)


def quaternion_right_product_matrix(q):
    w, x, y, z = q
    return np.array([[w, -x, -y, -z], [x, w, z, -y], [y, -z, w, x], [z, y, -x, w]], dtype=np.float64)


def MeasureUpdate(S: State, visionMeas, cfg):  # MeasureUpdate.m:37-102
    L1 = np.zeros((4, 3))
    L1[1, 0] = L1[2, 1] = L1[3, 2] = 0.5
    L2 = np.diag([1.0, -1.0, -1.0, -1.0])
    R_IL, P_IL, Q_IL = _extrinsics(cfg)
    i = _nearest(visionMeas)
    posMeas, quatMeas = visionMeas[i, 1:4], visionMeas[i, 4:8]
    markerPos, markerRot = cfg.markers[int(visionMeas[i, 0])]
    markerQuat = rotmat_to_quaternion(markerRot)
    posEst = (S.rotateMat @ R_IL.T).T @ (markerPos - S.position - S.rotateMat @ P_IL)
    quatEst = quaternion_add(quaternion_add(Q_IL, quaternion_conjugate(S.quaternion)), markerQuat)
    H = np.zeros((7, 18))
    H[0:3, 0:3] = -(S.rotateMat @ R_IL.T).T
    H[0:3, 6:9] = R_IL @ vector_to_crossmat(S.rotateMat.T @ (markerPos - S.position))
    Hq = quaternion_right_product_matrix(markerQuat) @ quaternion_left_product_matrix(Q_IL) @ L2 @ \
        quaternion_left_product_matrix(S.quaternion) @ L1
    H[3:7, 6:9] = Hq
    if np.linalg.norm(quatMeas - quatEst) > np.linalg.norm(quatMeas + quatEst):
        quatEst = -quatEst
        H[3:7, 6:9] = -Hq
    K = S.covariance @ H.T @ np.linalg.inv(H @ S.covariance @ H.T + S.measureNoise)
    err = np.concatenate([posMeas - posEst, np.zeros(4)])  # MeasureUpdate.m:88: the quaternion rows are zeroed
    dX = K @ err
    S.position = S.position + dX[0:3]
    S.velocity = S.velocity + dX[3:6]
    S.quaternion = quaternion_normalize(quaternion_add(S.quaternion, axisangle_to_quaternion(dX[6:9], np.linalg.norm(dX[6:9]))))
    S.accelBias = S.accelBias + dX[9:12]
    S.gyroBias = S.gyroBias + dX[12:15]
    S.gravity = S.gravity + dX[15:18]
    P = (np.eye(18) - K @ H) @ S.covariance
    S.covariance = (P + P.T) / 2


def run_script(cfg: Config, imudata, imgdata, n_init=500, max_frames=None, trace_cov=False):
    """FBUS_EKF.m:114-210: gravity / gyro bias from the first 500 rows, pose from the first image row, then the main loop --
    which starts AGAIN at the first image row (nImgData = 1), so the initialising frame also gets a measurement update.
    Returns rows [t p(3) q(4) v(3) ba(3) bg(3)] per loop iteration (the product's trace layout), the vision-only poses and,
    optionally, the covariances."""
    S = State(cfg)
    S.gravity, S.gyroBias = InitGravityAndGyrobias(imudata[:n_init])
    firstImgDataTime = imgdata[0, 0]
    i = 0
    while i < len(imudata) - 1 and not imudata[i, 0] > firstImgDataTime:  # FBUS_EKF.m:124-129
        i += 1
    imuDataIdx = i
    InitPositionAndQuaternion(S, imgdata[0:1, 1:9], cfg)
    rows, vis, covs, kinds = [], [], [], []
    preImgTime, n = 0.0, 0
    nImg = len(imgdata)
    while n < nImg - 1:  # FBUS_EKF.m:151 (the last row never starts a frame)
        m = n + 1
        while m < nImg and imgdata[m, 0] == imgdata[n, 0]:
            m += 1
        curImgTime = imgdata[n, 0]
        visionMeas = imgdata[n:m, 1:9]
        n = m
        if curImgTime - preImgTime > cfg.reset_gap and preImgTime != 0:
            ResetState(S, visionMeas, cfg)
            preImgTime = curImgTime
            kinds.append("reset")
        else:
            preImuTime = imudata[imuDataIdx - 1, 0]
            j = imuDataIdx
            while j < len(imudata):
                if imudata[j, 0] > curImgTime:
                    break
                if imudata[j, 0] < preImgTime:
                    preImuTime = imudata[j, 0]
                    j += 1
                    continue
                dt = imudata[j, 0] - preImuTime
                preImuTime = imudata[j, 0]
                ImuUpdate(S, imudata[j, 1:4], imudata[j, 4:7], dt)
                j += 1
            imuDataIdx = j
            preImgTime = curImgTime
            MeasureUpdate(S, visionMeas, cfg)
            kinds.append("update")
        rows.append(np.concatenate([[curImgTime], S.position, S.quaternion, S.velocity, S.accelBias, S.gyroBias]))
        vis.append(np.concatenate(ComputeVisionOnlyResults(visionMeas, cfg)))
        if trace_cov:
            covs.append(S.covariance.copy())
        if max_frames and len(rows) >= max_frames:
            break
    return {"rows": np.array(rows), "vision": np.array(vis), "P": np.array(covs) if trace_cov else None, "state": S, "kinds": kinds}


def default_config():
    import fbus_oracle_np as onp
    d = onp.default_config()
    return Config(d.tsc_left, d.markers)
