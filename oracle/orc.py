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

    def __init__(self, cfg, batch: int):
        self.cfg = cfg
        self.batch = batch
        self.h = lib().orc_create(C.byref(cfg), batch)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    @staticmethod
    def _ck(rc):
        if rc != 0:
            raise RuntimeError(f"oracle call failed rc={rc}")

    def init_gravity_gyrobias(self, imu, first, count):
        self._ck(lib().orc_init_gravity_gyrobias(self.h, C.byref(imu), first, count))

    def init_position_quaternion(self, det, frame, n_imu_before=1):
        self._ck(lib().orc_init_position_quaternion(self.h, C.byref(det), frame, n_imu_before))

    def propagate(self, imu, first, count, t_end):
        self._ck(lib().orc_propagate(self.h, C.byref(imu), first, count, t_end))

    def reset_state(self, det, frame):
        self._ck(lib().orc_reset_state(self.h, C.byref(det), frame))

    def update(self, det, frame):
        self._ck(lib().orc_update(self.h, C.byref(det), frame))

    def step_windows(self, imu, det, win_off, w0, w1, trace=None, n_threads=1):
        win_off = np.ascontiguousarray(win_off, dtype=np.uint32)
        tp = trace.ctypes.data if trace is not None else None
        self._ck(lib().orc_step_windows(self.h, C.byref(imu), C.byref(det), win_off.ctypes.data_as(capi.c_uint32_p), w0, w1,
                                        tp, n_threads))

    def get_state(self, with_cov=True):
        arrs = capi.alloc_state(self.batch, with_cov)
        sv = capi.state_view(arrs, self.batch)
        self._ck(lib().orc_get_state(self.h, C.byref(sv)))
        return arrs

    def set_state(self, arrs):
        sv = capi.state_view(arrs, self.batch)
        self._ck(lib().orc_set_state(self.h, C.byref(sv)))

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
