"""interpn_b200 — B200-native (sm_100a) implementation of InterpN's interpolation hot path.

Mirrors the reference's Python surface for that path (names, argument order, error behaviour):

* ``interpn_b200.raw`` — the 16 raw bindings of ``interpn.raw`` (src/python.rs);
* ``interpn_b200.interpn(...)`` — the convenience function of ``interpn/__init__.py:48-194``;
* ``MultilinearRegular`` … ``NearestRectilinear`` — the six pydantic interpolator classes;
* ``interpn_b200.one_dim`` — the Rust ``one_dim`` module;
* ``Interpolator`` — grid-resident evaluation on host arrays, raw device pointers or torch tensors.

All arithmetic runs in hand-written CUDA kernels behind the C ABI of include/interpn_b200.h.
There is no CPU fallback: importing fails without the built library and compute calls raise
``InterpnDeviceError`` without an sm_100 GPU.
"""

from __future__ import annotations

from . import _lib, one_dim, raw
from ._lib import (
    InterpnDeviceError,
    copy_threads,
    device_count,
    host_devices,
    launch_count,
    set_device,
    set_host_devices,
    swept_launch_count,
)
from .interpolator import Interpolator
from .api import (
    MulticubicRectilinear,
    MulticubicRegular,
    MultilinearRectilinear,
    MultilinearRegular,
    NearestRectilinear,
    NearestRegular,
    interpn,
)

__version__ = "0.1.0"

__all__ = [
    "__version__",
    "raw",
    "one_dim",
    "interpn",
    "Interpolator",
    "InterpnDeviceError",
    "MultilinearRegular",
    "MultilinearRectilinear",
    "MulticubicRegular",
    "MulticubicRectilinear",
    "NearestRegular",
    "NearestRectilinear",
    "copy_threads",
    "device_count",
    "host_devices",
    "launch_count",
    "set_device",
    "set_host_devices",
    "swept_launch_count",
]
