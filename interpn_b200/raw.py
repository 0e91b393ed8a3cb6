"""Drop-in replacements for ``interpn.raw`` — the 16 PyO3 functions of the reference
(/root/reference/src/python.rs:13-39, typed in src/interpn/raw.pyi) — bound to the sm_100a
C-ABI library instead of the Rust crate.

Same names, argument order and meaning; arrays must be 1-D, C-contiguous numpy arrays of the
function's dtype (the PyO3 extractors reject anything else with TypeError); every reference
``Err(msg)`` surfaces as ``AssertionError(msg)`` exactly like python.rs:78. Inputs are host
arrays: each call copies the grid and the query batch to the GPU, evaluates there, and copies
the result into ``out``. For repeated evaluation on one grid or for device-resident data use
``interpn_b200.Interpolator`` (the grid stays in HBM).

Arithmetic flavour. The reference exists in two builds that differ by a few ulp: the crate's default
features (every ``a*b + c`` rounded twice) and its ``fma`` feature — the build of the published Python
wheel (pyproject.toml:72). This package follows the CRATE by default (``libinterpn_b200.so``); a Python
user who wants the wheel's bits sets ``INTERPN_B200_ARITHMETIC=fma`` before the first import
(``libinterpn_b200_fma.so``, bit-identical to the wheel's own outputs on tests/golden/ref_docs.npz).
"""

from __future__ import annotations

import ctypes as C
from collections.abc import Sequence

import numpy as np

from . import _lib
from ._lib import lib

__all__ = [
    "interpn_linear_regular_f64",
    "interpn_linear_regular_f32",
    "interpn_linear_rectilinear_f64",
    "interpn_linear_rectilinear_f32",
    "interpn_nearest_regular_f64",
    "interpn_nearest_regular_f32",
    "interpn_nearest_rectilinear_f64",
    "interpn_nearest_rectilinear_f32",
    "interpn_cubic_regular_f64",
    "interpn_cubic_regular_f32",
    "interpn_cubic_rectilinear_f64",
    "interpn_cubic_rectilinear_f32",
    "check_bounds_regular_f64",
    "check_bounds_regular_f32",
    "check_bounds_rectilinear_f64",
    "check_bounds_rectilinear_f32",
]

_CT = {"f64": (np.dtype(np.float64), C.c_double), "f32": (np.dtype(np.float32), C.c_float)}


def _arr(a, dtype: np.dtype, name: str, writable: bool = False) -> np.ndarray:
    """What `PyReadonlyArray1<T>` + `.as_slice()` accept (python.rs:41-53, 60-64)."""
    if not isinstance(a, np.ndarray):
        raise TypeError(f"argument '{name}': expected a numpy array of {dtype}, got {type(a).__name__}")
    if a.dtype != dtype:
        raise TypeError(f"argument '{name}': expected dtype {dtype}, got {a.dtype}")
    if a.ndim != 1:
        raise TypeError(f"argument '{name}': expected a 1-dimensional array, got {a.ndim} dimensions")
    if not a.flags.c_contiguous:
        raise TypeError(f"argument '{name}': the given array is not contiguous")
    if writable and not a.flags.writeable:
        raise TypeError(f"argument '{name}': array is read-only")
    return a


def _dims(dims) -> np.ndarray:
    """`dims: Vec<usize>` (python.rs:59): any sequence of non-negative integers."""
    out = np.empty(len(dims), dtype=np.uint64)
    for i, d in enumerate(dims):
        d = int(d)
        if d < 0:
            raise OverflowError("can't convert negative int to unsigned")
        out[i] = d
    return out


def _ptrs(arrs: Sequence[np.ndarray], ct):
    n = len(arrs)
    if n > 8:
        # `unpack_vec_of_arr!` indexes a fixed [_; 8] (python.rs:46-50): the reference panics here.
        raise AssertionError("Dimension exceeds maximum (8). Use interpolator struct directly for higher dimensions.")
    ptrs = (C.POINTER(ct) * max(n, 1))()
    lens = (C.c_size_t * max(n, 1))()
    for i, a in enumerate(arrs):
        ptrs[i] = a.ctypes.data_as(C.POINTER(ct))
        lens[i] = a.size
    return ptrs, lens


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def _regular(method: str, sfx: str, dims, starts, steps, vals, obs, out, linearize=None):
    dt, ct = _CT[sfx]
    dims_a = _dims(dims)
    starts, steps, vals = _arr(starts, dt, "starts"), _arr(steps, dt, "steps"), _arr(vals, dt, "vals")
    obs = [_arr(o, dt, "obs") for o in obs]
    out = _arr(out, dt, "out", writable=True)
    optrs, olens = _ptrs(obs, ct)
    fn = getattr(lib, f"interpn_b200_{method}_regular_{sfx}")
    args = [
        dims_a.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(dims_a.size),
        _p(starts, ct), C.c_size_t(starts.size),
        _p(steps, ct), C.c_size_t(steps.size),
        _p(vals, ct), C.c_size_t(vals.size),
    ]  # fmt: skip
    if linearize is not None:
        args.append(C.c_int(1 if linearize else 0))
    args += [optrs, olens, C.c_size_t(len(obs)), _p(out, ct), C.c_size_t(out.size), None]
    _lib.check(fn(*args))


def _rectilinear(method: str, sfx: str, grids, vals, obs, out, linearize=None):
    dt, ct = _CT[sfx]
    grids = [_arr(g, dt, "grids") for g in grids]
    vals = _arr(vals, dt, "vals")
    obs = [_arr(o, dt, "obs") for o in obs]
    out = _arr(out, dt, "out", writable=True)
    gptrs, glens = _ptrs(grids, ct)
    optrs, olens = _ptrs(obs, ct)
    fn = getattr(lib, f"interpn_b200_{method}_rectilinear_{sfx}")
    args = [gptrs, glens, C.c_size_t(len(grids)), _p(vals, ct), C.c_size_t(vals.size)]
    if linearize is not None:
        args.append(C.c_int(1 if linearize else 0))
    args += [optrs, olens, C.c_size_t(len(obs)), _p(out, ct), C.c_size_t(out.size)]
    _lib.check(fn(*args))


def _bool_out(out) -> np.ndarray:
    if not isinstance(out, np.ndarray) or out.dtype != np.bool_ or out.ndim != 1 or not out.flags.c_contiguous:
        raise TypeError("argument 'out': expected a contiguous 1-dimensional numpy array of bool")
    return out


def _check_bounds_regular(sfx: str, dims, starts, steps, obs, atol, out):
    dt, ct = _CT[sfx]
    dims_a = _dims(dims)
    starts, steps = _arr(starts, dt, "starts"), _arr(steps, dt, "steps")
    obs = [_arr(o, dt, "obs") for o in obs]
    out = _bool_out(out)
    optrs, olens = _ptrs(obs, ct)
    fn = getattr(lib, f"interpn_b200_check_bounds_regular_{sfx}")
    _lib.check(
        fn(
            dims_a.ctypes.data_as(C.POINTER(C.c_size_t)), C.c_size_t(dims_a.size),
            _p(starts, ct), C.c_size_t(starts.size), _p(steps, ct), C.c_size_t(steps.size),
            optrs, olens, C.c_size_t(len(obs)), ct(float(atol)),
            out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(out.size),
        )  # fmt: skip
    )


def _check_bounds_rectilinear(sfx: str, grids, obs, atol, out):
    dt, ct = _CT[sfx]
    grids = [_arr(g, dt, "grids") for g in grids]
    obs = [_arr(o, dt, "obs") for o in obs]
    out = _bool_out(out)
    gptrs, glens = _ptrs(grids, ct)
    optrs, olens = _ptrs(obs, ct)
    fn = getattr(lib, f"interpn_b200_check_bounds_rectilinear_{sfx}")
    _lib.check(
        fn(
            gptrs, glens, C.c_size_t(len(grids)), optrs, olens, C.c_size_t(len(obs)), ct(float(atol)),
            out.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_size_t(out.size),
        )  # fmt: skip
    )


# ---- python.rs:55-85 ------------------------------------------------------------------------
def interpn_linear_regular_f64(dims, starts, steps, vals, obs, out) -> None:
    _regular("linear", "f64", dims, starts, steps, vals, obs, out)


def interpn_linear_regular_f32(dims, starts, steps, vals, obs, out) -> None:
    _regular("linear", "f32", dims, starts, steps, vals, obs, out)


# ---- python.rs:119-147 ----------------------------------------------------------------------
def interpn_linear_rectilinear_f64(grids, vals, obs, out) -> None:
    _rectilinear("linear", "f64", grids, vals, obs, out)


def interpn_linear_rectilinear_f32(grids, vals, obs, out) -> None:
    _rectilinear("linear", "f32", grids, vals, obs, out)


# ---- python.rs:149-178 ----------------------------------------------------------------------
def interpn_nearest_regular_f64(dims, starts, steps, vals, obs, out) -> None:
    _regular("nearest", "f64", dims, starts, steps, vals, obs, out)


def interpn_nearest_regular_f32(dims, starts, steps, vals, obs, out) -> None:
    _regular("nearest", "f32", dims, starts, steps, vals, obs, out)


# ---- python.rs:180-201 ----------------------------------------------------------------------
def interpn_nearest_rectilinear_f64(grids, vals, obs, out) -> None:
    _rectilinear("nearest", "f64", grids, vals, obs, out)


def interpn_nearest_rectilinear_f32(grids, vals, obs, out) -> None:
    _rectilinear("nearest", "f32", grids, vals, obs, out)


# ---- python.rs:228-260 ----------------------------------------------------------------------
def interpn_cubic_regular_f64(dims, starts, steps, vals, linearize_extrapolation, obs, out) -> None:
    _regular("cubic", "f64", dims, starts, steps, vals, obs, out, bool(linearize_extrapolation))


def interpn_cubic_regular_f32(dims, starts, steps, vals, linearize_extrapolation, obs, out) -> None:
    _regular("cubic", "f32", dims, starts, steps, vals, obs, out, bool(linearize_extrapolation))


# ---- python.rs:262-292 ----------------------------------------------------------------------
def interpn_cubic_rectilinear_f64(grids, vals, linearize_extrapolation, obs, out) -> None:
    _rectilinear("cubic", "f64", grids, vals, obs, out, bool(linearize_extrapolation))


def interpn_cubic_rectilinear_f32(grids, vals, linearize_extrapolation, obs, out) -> None:
    _rectilinear("cubic", "f32", grids, vals, obs, out, bool(linearize_extrapolation))


# ---- python.rs:87-117 -----------------------------------------------------------------------
def check_bounds_regular_f64(dims, starts, steps, obs, atol, out) -> None:
    _check_bounds_regular("f64", dims, starts, steps, obs, atol, out)


def check_bounds_regular_f32(dims, starts, steps, obs, atol, out) -> None:
    _check_bounds_regular("f32", dims, starts, steps, obs, atol, out)


# ---- python.rs:203-226 ----------------------------------------------------------------------
def check_bounds_rectilinear_f64(grids, obs, atol, out) -> None:
    _check_bounds_rectilinear("f64", grids, obs, atol, out)


def check_bounds_rectilinear_f32(grids, obs, atol, out) -> None:
    _check_bounds_rectilinear("f32", grids, obs, atol, out)
