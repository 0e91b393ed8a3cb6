"""Loader for the nvcc-built C-ABI library ``libinterpn_b200.so`` (include/interpn_b200.h).

There is deliberately no fallback: if the shared library is missing the import fails, and if no
sm_100 device is usable every compute call raises ``InterpnDeviceError``. Nothing under
``oracle/`` is ever imported from here.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# Arithmetic flavour (include/interpn_b200.h interpn_b200_arithmetic): "strict" reproduces the reference crate built
# with default features, "fma" the crate's `fma` feature — the build of the reference's Python wheel
# (pyproject.toml:72). Chosen once per process by INTERPN_B200_ARITHMETIC; each flavour is its own shared library.
ARITHMETIC = os.environ.get("INTERPN_B200_ARITHMETIC", "strict").lower()
if ARITHMETIC not in ("strict", "fma"):
    raise ImportError(f"INTERPN_B200_ARITHMETIC must be 'strict' or 'fma', not {ARITHMETIC!r}")
# INTERPN_B200_LIBRARY selects another build of the same library (kernel-tuning experiments only).
LIB_PATH = os.environ.get("INTERPN_B200_LIBRARY") or os.path.join(
    _HERE, "libinterpn_b200_fma.so" if ARITHMETIC == "fma" else "libinterpn_b200.so"
)

# Status codes of include/interpn_b200.h
OK = 0
ERR_UNREPRESENTABLE = 7
ERR_UNREPRESENTABLE_NUM = 11
ERR_CUDA = 100
ERR_NO_DEVICE = 101
ERR_INVALID_ARG = 102
ERR_TOO_LARGE = 103

LINEAR, CUBIC, NEAREST = 0, 1, 2
VALS_HOST, VALS_DEVICE, VALS_UNINIT = 0, 1, 2
KINDS_1D = {"linear": 0, "linear_hold_last": 1, "left": 2, "right": 3, "nearest": 4}
NO_BAD = (1 << 64) - 1


class InterpnDeviceError(RuntimeError):
    """CUDA-side failure (no device, wrong architecture, out of memory, runtime error)."""


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C interpn_b200/csrc`. interpn_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    lib.interpn_b200_strerror.restype = C.c_char_p
    lib.interpn_b200_strerror.argtypes = [C.c_int]
    lib.interpn_b200_last_error_detail.restype = C.c_char_p
    lib.interpn_b200_launch_count.restype = C.c_uint64
    lib.interpn_b200_swept_launch_count.restype = C.c_uint64
    lib.interpn_b200_interp_vals_ptr.restype = C.c_void_p
    lib.interpn_b200_interp_vals_ptr.argtypes = [C.c_void_p]
    for name in ("interpn_b200_interp_vals_len", "interpn_b200_interp_elem_size", "interpn_b200_interp_ndims"):
        getattr(lib, name).restype = C.c_size_t
        getattr(lib, name).argtypes = [C.c_void_p]
    lib.interpn_b200_interp_vals_updated.argtypes = [C.c_void_p, C.c_void_p]
    lib.interpn_b200_interp_free.restype = None
    lib.interpn_b200_interp_free.argtypes = [C.c_void_p]
    lib.interpn_b200_interp_status.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_size_t)]
    return lib


lib = _load()


def check(status: int) -> None:
    """Map a C status to the exception the reference's Python bindings raise.

    The reference maps every ``Err(msg)`` to ``AssertionError(msg)`` (src/python.rs:78);
    library-level CUDA failures have no reference analogue and raise InterpnDeviceError.
    """
    if status == OK:
        return
    msg = lib.interpn_b200_strerror(status).decode()
    if status >= ERR_CUDA and status != ERR_INVALID_ARG:
        detail = lib.interpn_b200_last_error_detail().decode()
        raise InterpnDeviceError(f"{msg}: {detail}" if detail else msg)
    if status == ERR_INVALID_ARG:
        raise ValueError(msg)
    raise AssertionError(msg)


def launch_count() -> int:
    return int(lib.interpn_b200_launch_count())


def swept_launch_count() -> int:
    return int(lib.interpn_b200_swept_launch_count())


def device_count() -> int:
    return int(lib.interpn_b200_device_count())


def set_device(device: int) -> None:
    check(lib.interpn_b200_set_device(C.c_int(device)))


def set_host_devices(n: int) -> None:
    """How many GPUs one host-buffer call (``raw.*``, ``Interpolator.eval``, the model classes) may use: 0 = every
    visible device (the default), 1 = only the interpolator's own (one process per GPU, e.g. under torchrun)."""
    check(lib.interpn_b200_set_host_devices(C.c_int(n)))


def host_devices() -> int:
    return int(lib.interpn_b200_host_devices())


def copy_threads() -> int:
    """Host threads that stage pageable caller memory (INTERPN_B200_COPY_THREADS)."""
    return int(lib.interpn_b200_copy_threads())
