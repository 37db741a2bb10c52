"""ctypes binding of the CPU ORACLE (oracle/fbus_oracle.cpp).  TEST INFRASTRUCTURE, NOT PRODUCT:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
It reuses the product's ctypes struct definitions (the oracle takes the same C structs)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "_build", "libfbus_oracle.so")


if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from fbus_ekf_b200 import capi  # noqa: E402  (struct definitions only; the oracle never calls the product library)
_lib = None


def build():
    subprocess.check_call(["make", "-C", HERE, "-s"])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    H = C.c_void_p
    P = C.POINTER
    L.orc_create.restype = H
    L.orc_create.argtypes = [P(capi.FbusConfig), C.c_size_t]
    L.orc_destroy.restype = None
    L.orc_destroy.argtypes = [H]
    L.orc_init_gravity_gyrobias.argtypes = [H, P(capi.ImuStream), C.c_size_t, C.c_size_t]
    L.orc_init_position_quaternion.argtypes = [H, P(capi.DetFrames), C.c_size_t, C.c_size_t]
    L.orc_propagate.argtypes = [H, P(capi.ImuStream), C.c_size_t, C.c_size_t, C.c_double]
    L.orc_reset_state.argtypes = [H, P(capi.DetFrames), C.c_size_t]
    L.orc_update.argtypes = [H, P(capi.DetFrames), C.c_size_t]
    L.orc_step_windows.argtypes = [H, P(capi.ImuStream), P(capi.DetFrames), capi.c_uint32_p, C.c_size_t, C.c_size_t,
                                   C.c_void_p, C.c_int]
    L.orc_refract_solve.argtypes = [P(capi.FbusConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    L.orc_marker_pose.argtypes = [P(capi.FbusConfig), C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_inair_solve.argtypes = [P(capi.FbusConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
    L.orc_inair_solve.restype = C.c_int
    L.orc_get_state.argtypes = [H, P(capi.StateSoa)]
    L.orc_set_state.argtypes = [H, P(capi.StateSoa)]
    L.orc_stats.argtypes = [H, C.c_void_p, C.c_void_p, capi.c_double_p]
    for n in ("orc_init_gravity_gyrobias", "orc_init_position_quaternion", "orc_propagate", "orc_reset_state", "orc_update",
              "orc_step_windows", "orc_refract_solve", "orc_marker_pose", "orc_get_state", "orc_set_state", "orc_stats"):
        getattr(L, n).restype = C.c_int
    _lib = L
    return L


class Oracle:
    """Batch of independent scalar reference filters on the CPU."""

    _pfx = "orc_"

    @staticmethod
    def _lib():
        return lib()

    def _fn(self, name):
        return getattr(self._lib(), self._pfx + name)

    def __init__(self, cfg, batch: int):
        self.cfg = cfg
        self.batch = batch
        self.h = self._fn("create")(C.byref(cfg), batch)
        if not self.h:
            raise RuntimeError(f"{self._pfx}create refused this configuration")

    def __del__(self):
        if getattr(self, "h", None):
            self._fn("destroy")(self.h)
            self.h = None

    @staticmethod
    def _ck(rc):
        if rc != 0:
            raise RuntimeError(f"oracle call failed rc={rc}")

    def init_gravity_gyrobias(self, imu, first, count):
        self._ck(self._fn("init_gravity_gyrobias")(self.h, C.byref(imu), first, count))

    def init_position_quaternion(self, det, frame, n_imu_before=1):
        self._ck(self._fn("init_position_quaternion")(self.h, C.byref(det), frame, n_imu_before))

    def propagate(self, imu, first, count, t_end):
        self._ck(self._fn("propagate")(self.h, C.byref(imu), first, count, t_end))

    def reset_state(self, det, frame):
        self._ck(self._fn("reset_state")(self.h, C.byref(det), frame))

    def update(self, det, frame):
        self._ck(self._fn("update")(self.h, C.byref(det), frame))

    def step_windows(self, imu, det, win_off, w0, w1, trace=None, n_threads=1):
        win_off = np.ascontiguousarray(win_off, dtype=np.uint32)
        tp = trace.ctypes.data if trace is not None else None
        self._ck(self._fn("step_windows")(self.h, C.byref(imu), C.byref(det), win_off.ctypes.data_as(capi.c_uint32_p), w0, w1,
                                        tp, n_threads))

    def get_state(self, with_cov=True):
        arrs = capi.alloc_state(self.batch, with_cov)
        sv = capi.state_view(arrs, self.batch)
        self._ck(self._fn("get_state")(self.h, C.byref(sv)))
        return arrs

    def set_state(self, arrs):
        sv = capi.state_view(arrs, self.batch)
        self._ck(self._fn("set_state")(self.h, C.byref(sv)))

    def stats(self, truth_p, truth_q):
        out = np.zeros(capi.FBUS_NSTATS)
        self._ck(lib().orc_stats(self.h, truth_p.ctypes.data, truth_q.ctypes.data, capi.dptr(out)))
        return out


def refract_solve(cfg, corners: np.ndarray, n_threads=1):
    """corners float32 [16][n] -> pose [7][n], corners3d [12][n], valid [n]"""
    n = corners.shape[1]
    assert corners.dtype == np.float32 and corners.flags["C_CONTIGUOUS"]
    pose = np.zeros((7, n))
    c3 = np.zeros((12, n))
    valid = np.zeros(n, dtype=np.int32)
    rc = lib().orc_refract_solve(C.byref(cfg), corners.ctypes.data, n, pose.ctypes.data, c3.ctypes.data, valid.ctypes.data, n_threads)
    assert rc == 0
    return pose, c3, valid


def undistort_fisheye(cfg, pixels: np.ndarray):
    n = pixels.shape[1]
    assert pixels.dtype == np.float32 and pixels.flags["C_CONTIGUOUS"] and pixels.shape[0] == 16
    out = np.zeros_like(pixels)
    L = lib()
    L.orc_undistort_fisheye.argtypes = [C.POINTER(capi.FbusConfig), C.c_void_p, C.c_size_t, C.c_void_p]
    L.orc_undistort_fisheye.restype = C.c_int
    assert L.orc_undistort_fisheye(C.byref(cfg), pixels.ctypes.data, n, out.ctypes.data) == 0
    return out


def inair_solve(cfg, corners: np.ndarray):
    """corners float32 [16][n] -> pose [7][n], corners3d [12][n], valid [n]  (NormalTriangulation + ComputeMarkerPose)"""
    n = corners.shape[1]
    assert corners.dtype == np.float32 and corners.flags["C_CONTIGUOUS"]
    pose, c3, valid = np.zeros((7, n)), np.zeros((12, n)), np.zeros(n, dtype=np.int32)
    rc = lib().orc_inair_solve(C.byref(cfg), corners.ctypes.data, n, pose.ctypes.data, c3.ctypes.data, valid.ctypes.data)
    assert rc == 0
    return pose, c3, valid


def marker_pose(cfg, corners3d: np.ndarray):
    n = corners3d.shape[1]
    assert corners3d.dtype == np.float64 and corners3d.flags["C_CONTIGUOUS"]
    pose = np.zeros((7, n))
    rc = lib().orc_marker_pose(C.byref(cfg), corners3d.ctypes.data, n, pose.ctypes.data)
    assert rc == 0
    return pose


# ---------------------------------------------------------------------------------------------------------------------
# oracle/_ref: the REFERENCE'S OWN filter.cpp, compiled unmodified against stand-in headers (oracle/ref_build/)
# ---------------------------------------------------------------------------------------------------------------------
REF_DIR = os.path.join(HERE, "_ref")
REF_LIB_PATH = os.path.join(REF_DIR, "libfbus_ref.so")
REF_DBG_LIB_PATH = os.path.join(REF_DIR, "libfbus_ref_dbg.so")
REF_AVX2_LIB_PATH = os.path.join(REF_DIR, "libfbus_ref_avx2.so")
_ref_libs = {}


def build_ref():
    """compiles /root/reference/C++/src/filter.cpp where it lies (no-op when the reference tree is absent, e.g. on the GPU box)"""
    subprocess.check_call(["make", "-C", os.path.join(HERE, "ref_build"), "-s"])


def ref_available() -> bool:
    return os.path.exists(REF_LIB_PATH)


def ref_lib(variant: str = ""):
    path = {"": REF_LIB_PATH, "dbg": REF_DBG_LIB_PATH, "avx2": REF_AVX2_LIB_PATH}[variant]
    if path in _ref_libs:
        return _ref_libs[path]
    if not os.path.exists(path):
        build_ref()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and /root/reference is not here to build it from")
    L = C.CDLL(path)
    H = C.c_void_p
    P = C.POINTER
    L.ref_create.restype = H
    L.ref_create.argtypes = [P(capi.FbusConfig), C.c_size_t]
    L.ref_destroy.restype = None
    L.ref_destroy.argtypes = [H]
    L.ref_init_gravity_gyrobias.argtypes = [H, P(capi.ImuStream), C.c_size_t, C.c_size_t]
    L.ref_init_position_quaternion.argtypes = [H, P(capi.DetFrames), C.c_size_t, C.c_size_t]
    L.ref_propagate.argtypes = [H, P(capi.ImuStream), C.c_size_t, C.c_size_t, C.c_double]
    L.ref_reset_state.argtypes = [H, P(capi.DetFrames), C.c_size_t]
    L.ref_update.argtypes = [H, P(capi.DetFrames), C.c_size_t]
    L.ref_step_windows.argtypes = [H, P(capi.ImuStream), P(capi.DetFrames), capi.c_uint32_p, C.c_size_t, C.c_size_t,
                                   C.c_void_p, C.c_int]
    L.ref_set_imu_data.argtypes = [H, P(capi.ImuStream), C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                   P(C.c_size_t)]
    L.ref_get_poses.argtypes = [H, C.c_size_t, C.c_void_p, C.c_void_p]
    L.ref_get_state.argtypes = [H, P(capi.StateSoa)]
    L.ref_set_state.argtypes = [H, P(capi.StateSoa)]
    for n in ("ref_init_gravity_gyrobias", "ref_init_position_quaternion", "ref_propagate", "ref_reset_state", "ref_update",
              "ref_step_windows", "ref_set_imu_data", "ref_get_poses", "ref_get_state", "ref_set_state", "ref_abi_version"):
        getattr(L, n).restype = C.c_int
    assert L.ref_abi_version() == capi.FBUS_ABI_VERSION
    _ref_libs[path] = L
    return L


class Ref(Oracle):
    """Batch of FBUSEKF::FILTER objects of the reference itself (oracle/_ref); same calls as Oracle."""

    _pfx = "ref_"
    _variant = ""

    @classmethod
    def _lib(cls):
        return ref_lib(cls._variant)

    def stats(self, truth_p, truth_q):
        raise NotImplementedError("the reference has no statistics")

    def set_imu_data(self, imu, first, count, cap=4096):
        """FILTER::SetImuData fed sample by sample into an empty buffer -> (t [n], data [n,6]) left in the buffer"""
        t = np.zeros(cap)
        d = np.zeros((cap, 6))
        n = C.c_size_t(0)
        self._ck(self._fn("set_imu_data")(self.h, C.byref(imu), first, count, t.ctypes.data, d.ctypes.data, cap, C.byref(n)))
        return t[:n.value].copy(), d[:n.value].copy()

    def poses(self, b=0):
        cam, vis = np.zeros(16), np.zeros(16)
        self._ck(self._fn("get_poses")(self.h, b, cam.ctypes.data, vis.ctypes.data))
        return cam.reshape(4, 4), vis.reshape(4, 4)


def host_has_avx2() -> bool:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    return " avx2" in line
    except OSError:
        pass
    return False


class RefFast(Ref):
    """oracle/_ref built -O3 -mavx2 (bench.py's CPU baseline on hosts with AVX2); same arithmetic, one rounding per operation"""
    _variant = "avx2"


class RefDebug(Ref):
    """the same library with the stand-in headers' bounds / shape assertions compiled in"""
    _variant = "dbg"


def config_default():
    """fbus_config of the bundled logs WITHOUT the product library: filled from the NumPy oracle's own restatement of
    camerainfo1.yml / paramconfig.yml / markersetup.yml (tests assert it is equal to the product's fbus_config_default)."""
    import fbus_oracle_np as onp
    d = onp.default_config()
    c = capi.FbusConfig()
    for i in range(16):
        c.tsc_left[i] = float(d.tsc_left.ravel()[i])
        c.tsc_right[i] = float(d.tsc_right.ravel()[i])
    for k in ("accel_n_cov", "gyro_n_cov", "accel_b_cov", "gyro_b_cov", "pos_n_cov", "quat_n_cov", "marker_max_dist",
              "marker_switch_thres", "reset_gap", "n_air", "n_glass", "n_water", "d_air", "d_glass", "marker_dect_dist_thres"):
        setattr(c, k, float(getattr(d, k)))
    for i in range(6):
        c.p0_diag[i] = float(d.p0_diag[i])
    for i in range(3):
        c.normal[i] = float(d.normal[i])
    c.marker_size = 0.28
    c.imu_g = 9.802
    # camerainfo1.yml K / D of the left and right camera (fx, fy, cx, cy ; Kannala-Brandt k1..k4): fisheye undistortion only
    for cam, (kk, dd) in enumerate((((246.134, 246.265, 325.504, 178.694), (0.584804, 0.158016, -0.5657, 0.272636)),
                                    ((245.124, 244.704, 341.197, 179.214), (0.590953, 0.140311, -0.490475, 0.206821)))):
        for i in range(4):
            c.cam_k[cam][i] = kk[i]
            c.cam_d[cam][i] = dd[i]
    c.n_markers = len(d.markers)
    for m, (mid, (p, R)) in enumerate(d.markers.items()):
        c.marker_id[m] = int(mid)
        for i in range(3):
            c.marker_pos[3 * m + i] = float(p[i])
        for i in range(9):
            c.marker_rot[9 * m + i] = float(R.ravel()[i])
    return c
