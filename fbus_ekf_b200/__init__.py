"""fbus_ekf_b200 -- B200-native (sm_100a) batched implementation of FBUS-EKF's filter-and-refraction hot path.

The product is the CUDA library `libfbus_ekf.so` behind the C ABI of include/fbus_ekf.h; this package is the
thin host-side mirror of the reference's filter interface (FBUSEKF::FILTER / the MATLAB function API).
Importing the package does not load the library; constructing a BatchFilter does, and fails loudly if the
library or a CUDA device is missing (there is no CPU fallback).
"""
from . import capi  # noqa: F401
from .filter import BatchFilter, FbusError  # noqa: F401

__all__ = ["capi", "BatchFilter", "FbusError"]
