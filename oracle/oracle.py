"""ctypes binding of the CPU ORACLE (``oracle/liboracle.so``).

TEST INFRASTRUCTURE ONLY. Importers: ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``. The product package
``interpn_b200`` never imports this module.

The oracle restates interpn 0.8.2's arithmetic operation-for-operation (see
``interpn_oracle.hpp``); argument order of the public helpers below follows the reference's Rust
``interpn(...)`` functions (e.g. ``/root/reference/src/multilinear/regular.rs:51-58``).
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

METHODS = {"linear": 0, "cubic": 1, "nearest": 2}
KINDS_1D = {"linear": 0, "linear_hold_last": 1, "left": 2, "right": 3, "nearest": 4}
ORDERS = {"reference": 0, "flattened": 1, "recursive": 2}

NO_BAD = np.iinfo(np.uint64).max
# Arithmetic the checks run in when a test does not say: the flavour of the library under test
# (interpn_b200/_lib.py reads the same variable), so the whole GPU suite can be re-run against the `fma` build.
DEFAULT_FMA = os.environ.get("INTERPN_B200_ARITHMETIC", "strict").lower() == "fma"


def build(force: bool = False) -> str:
    """Compile ``liboracle.so`` with the committed Makefile (g++ only)."""
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "interpn_oracle.hpp", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs
    )
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B", "liboracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_strerror.restype = C.c_char_p
        _lib.oracle_strerror.argtypes = [C.c_int]
    return _lib


class OracleError(AssertionError):
    """Mirrors the reference's ``Err(&'static str)`` (PyAssertionError at python.rs:78)."""

    def __init__(self, status: int, first_bad: int | None = None):
        self.status = status
        self.first_bad = first_bad
        super().__init__(lib().oracle_strerror(status).decode())


def _suffix(dtype) -> tuple[str, type]:
    dtype = np.dtype(dtype)
    if dtype == np.float64:
        return "f64", C.c_double
    if dtype == np.float32:
        return "f32", C.c_float
    raise TypeError(f"oracle supports float32/float64 only, got {dtype}")


def _ptr_array(arrs: Sequence[np.ndarray], ctype):
    n = len(arrs)
    ptrs = (C.POINTER(ctype) * max(n, 1))()
    lens = (C.c_size_t * max(n, 1))()
    for i, a in enumerate(arrs):
        ptrs[i] = a.ctypes.data_as(C.POINTER(ctype))
        lens[i] = a.size
    return ptrs, lens


def _prep(a, dtype) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=dtype)
    return a.reshape(-1)


def interpn_regular(
    method: str,
    dims,
    starts,
    steps,
    vals,
    obs: Sequence,
    out: np.ndarray | None = None,
    *,
    linearize_extrapolation: bool = True,
    fma: bool | None = None,
    order: str = "reference",
    nthreads: int = 1,
    dtype=None,
) -> np.ndarray:
    """Regular-grid evaluation: linear / cubic / nearest.

    Mirrors ``interpn(dims, starts, steps, vals, [linearize_extrapolation], obs, out)``
    (multilinear/regular.rs:51, multicubic/regular.rs:52, nearest/regular.rs:41).
    Raises OracleError with the reference's message on failure; ``first_bad`` carries the index of
    the first unrepresentable point.
    """
    dtype = np.dtype(dtype or np.asarray(vals).dtype)
    sfx, ct = _suffix(dtype)
    dims_a = np.ascontiguousarray(dims, dtype=np.uint64).reshape(-1)
    starts_a, steps_a, vals_a = _prep(starts, dtype), _prep(steps, dtype), _prep(vals, dtype)
    obs_a = [_prep(o, dtype) for o in obs]
    if out is None:
        out = np.zeros(obs_a[0].size if obs_a else 0, dtype=dtype)
    assert out.dtype == dtype and out.flags.c_contiguous
    optrs, olens = _ptr_array(obs_a, ct)
    first_bad = C.c_size_t(0)
    fn = getattr(lib(), f"oracle_regular_{sfx}")
    fn.restype = C.c_int
    st = fn(
        C.c_int(METHODS[method]),
        dims_a.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(dims_a.size),
        starts_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(starts_a.size),
        steps_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(steps_a.size),
        vals_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(vals_a.size),
        C.c_int(int(linearize_extrapolation)),
        optrs, olens, C.c_size_t(len(obs_a)),
        out.ctypes.data_as(C.POINTER(ct)), C.c_size_t(out.size),
        C.c_int(int(DEFAULT_FMA if fma is None else fma)), C.c_int(ORDERS[order]), C.c_int(nthreads), C.byref(first_bad),
    )  # fmt: skip
    if st != 0:
        raise OracleError(st, first_bad.value if first_bad.value != NO_BAD else None)
    return out


def interpn_rectilinear(
    method: str,
    grids: Sequence,
    vals,
    obs: Sequence,
    out: np.ndarray | None = None,
    *,
    linearize_extrapolation: bool = True,
    fma: bool | None = None,
    order: str = "reference",
    nthreads: int = 1,
    dtype=None,
) -> np.ndarray:
    """Rectilinear-grid evaluation; mirrors ``interpn(grids, vals, [linearize], obs, out)``
    (multilinear/rectilinear.rs:49, multicubic/rectilinear.rs:54, nearest/rectilinear.rs:39)."""
    dtype = np.dtype(dtype or np.asarray(vals).dtype)
    sfx, ct = _suffix(dtype)
    grids_a = [_prep(g, dtype) for g in grids]
    vals_a = _prep(vals, dtype)
    obs_a = [_prep(o, dtype) for o in obs]
    if out is None:
        out = np.zeros(obs_a[0].size if obs_a else 0, dtype=dtype)
    assert out.dtype == dtype and out.flags.c_contiguous
    gptrs, glens = _ptr_array(grids_a, ct)
    optrs, olens = _ptr_array(obs_a, ct)
    fn = getattr(lib(), f"oracle_rectilinear_{sfx}")
    fn.restype = C.c_int
    st = fn(
        C.c_int(METHODS[method]),
        gptrs, glens, C.c_size_t(len(grids_a)),
        vals_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(vals_a.size),
        C.c_int(int(linearize_extrapolation)),
        optrs, olens, C.c_size_t(len(obs_a)),
        out.ctypes.data_as(C.POINTER(ct)), C.c_size_t(out.size),
        C.c_int(int(DEFAULT_FMA if fma is None else fma)), C.c_int(ORDERS[order]), C.c_int(nthreads),
    )  # fmt: skip
    if st != 0:
        raise OracleError(st)
    return out


def one_dim_regular(kind: str, start, step, vals, locs, out=None, *, fma: bool | None = None, dtype=None) -> np.ndarray:
    """``Kind1D::new(RegularGrid1D::new(start, step, vals)?).eval(locs, out)`` (one_dim/mod.rs:41-138)."""
    dtype = np.dtype(dtype or np.asarray(vals).dtype)
    sfx, ct = _suffix(dtype)
    vals_a, locs_a = _prep(vals, dtype), _prep(locs, dtype)
    if out is None:
        out = np.zeros(locs_a.size, dtype=dtype)
    first_bad = C.c_size_t(0)
    fn = getattr(lib(), f"oracle_one_dim_regular_{sfx}")
    fn.restype = C.c_int
    st = fn(
        C.c_int(KINDS_1D[kind]), ct(start), ct(step),
        vals_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(vals_a.size),
        locs_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(locs_a.size),
        out.ctypes.data_as(C.POINTER(ct)), C.c_size_t(out.size),
        C.c_int(int(DEFAULT_FMA if fma is None else fma)), C.byref(first_bad),
    )  # fmt: skip
    if st != 0:
        raise OracleError(st, first_bad.value if first_bad.value != NO_BAD else None)
    return out


def one_dim_rectilinear(kind: str, grid, vals, locs, out=None, *, fma: bool | None = None, dtype=None) -> np.ndarray:
    """``Kind1D::new(RectilinearGrid1D::new(grid, vals)?).eval(locs, out)`` (one_dim/mod.rs:142-187)."""
    dtype = np.dtype(dtype or np.asarray(vals).dtype)
    sfx, ct = _suffix(dtype)
    grid_a, vals_a, locs_a = _prep(grid, dtype), _prep(vals, dtype), _prep(locs, dtype)
    if out is None:
        out = np.zeros(locs_a.size, dtype=dtype)
    fn = getattr(lib(), f"oracle_one_dim_rectilinear_{sfx}")
    fn.restype = C.c_int
    st = fn(
        C.c_int(KINDS_1D[kind]),
        grid_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(grid_a.size),
        vals_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(vals_a.size),
        locs_a.ctypes.data_as(C.POINTER(ct)), C.c_size_t(locs_a.size),
        out.ctypes.data_as(C.POINTER(ct)), C.c_size_t(out.size),
        C.c_int(int(DEFAULT_FMA if fma is None else fma)),
    )  # fmt: skip
    if st != 0:
        raise OracleError(st)
    return out


def check_bounds_regular(dims, starts, steps, obs: Sequence, atol: float, out=None, *, dtype=None) -> np.ndarray:
    """multilinear/regular.rs:145-182."""
    dtype = np.dtype(dtype or np.asarray(starts).dtype)
    sfx, ct = _suffix(dtype)
    dims_a = np.ascontiguousarray(dims, dtype=np.uint64).reshape(-1)
    starts_a, steps_a = _prep(starts, dtype), _prep(steps, dtype)
    obs_a = [_prep(o, dtype) for o in obs]
    if out is None:
        out = np.zeros(dims_a.size, dtype=np.bool_)
    optrs, olens = _ptr_array(obs_a, ct)
    fn = getattr(lib(), f"oracle_check_bounds_regular_{sfx}")
    fn.restype = C.c_int
    st = fn(
        dims_a.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(dims_a.size),
        starts_a.ctypes.data_as(C.POINTER(ct)), steps_a.ctypes.data_as(C.POINTER(ct)),
        optrs, olens, C.c_size_t(len(obs_a)), ct(atol),
        out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(out.size),
    )  # fmt: skip
    if st != 0:
        raise OracleError(st)
    return out


def check_bounds_rectilinear(grids: Sequence, obs: Sequence, atol: float, out=None, *, dtype=None) -> np.ndarray:
    """multilinear/rectilinear.rs:109-134."""
    dtype = np.dtype(dtype or np.asarray(grids[0]).dtype)
    sfx, ct = _suffix(dtype)
    grids_a = [_prep(g, dtype) for g in grids]
    obs_a = [_prep(o, dtype) for o in obs]
    if out is None:
        out = np.zeros(len(grids_a), dtype=np.bool_)
    gptrs, glens = _ptr_array(grids_a, ct)
    optrs, olens = _ptr_array(obs_a, ct)
    fn = getattr(lib(), f"oracle_check_bounds_rectilinear_{sfx}")
    fn.restype = C.c_int
    st = fn(
        gptrs, glens, C.c_size_t(len(grids_a)),
        optrs, olens, C.c_size_t(len(obs_a)), ct(atol),
        out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(out.size),
    )  # fmt: skip
    if st != 0:
        raise OracleError(st)
    return out


def max_threads() -> int:
    lib().oracle_max_threads.restype = C.c_int
    return int(lib().oracle_max_threads())
