"""1-D specialisations: the reference's `one_dim` module
(/root/reference/src/one_dim/{mod,linear,hold}.rs) evaluated on the GPU.

`eval_regular(kind, start, step, vals, locs, out)` is
`Kind::new(RegularGrid1D::new(start, step, vals)?).eval(locs, out)`;
`eval_rectilinear(kind, grid, vals, locs, out)` is the `RectilinearGrid1D` twin. `kind` is one
of "linear" (Linear1D), "linear_hold_last" (LinearHoldLast1D), "left" (Left1D), "right"
(Right1D), "nearest" (Nearest1D). Errors surface as AssertionError with the reference's
messages ("Length mismatch", "Unrepresentable number").
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import lib
from .raw import _CT, _arr, _p

KINDS = tuple(_lib.KINDS_1D)


def _sfx(dtype) -> str:
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"Unexpected data type: {dtype}")


def eval_regular(kind: str, start, step, vals: np.ndarray, locs: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    sfx = _sfx(vals.dtype)
    dt, ct = _CT[sfx]
    vals, locs = _arr(vals, dt, "vals"), _arr(locs, dt, "locs")
    if out is None:
        out = np.zeros_like(locs)
    out = _arr(out, dt, "out", writable=True)
    fn = getattr(lib, f"interpn_b200_one_dim_regular_{sfx}")
    _lib.check(
        fn(C.c_int(_lib.KINDS_1D[kind]), ct(float(start)), ct(float(step)), _p(vals, ct), C.c_size_t(vals.size),
           _p(locs, ct), C.c_size_t(locs.size), _p(out, ct), C.c_size_t(out.size), None)  # fmt: skip
    )
    return out


def eval_rectilinear(kind: str, grid: np.ndarray, vals: np.ndarray, locs: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
    sfx = _sfx(vals.dtype)
    dt, ct = _CT[sfx]
    grid, vals, locs = _arr(grid, dt, "grid"), _arr(vals, dt, "vals"), _arr(locs, dt, "locs")
    if out is None:
        out = np.zeros_like(locs)
    out = _arr(out, dt, "out", writable=True)
    fn = getattr(lib, f"interpn_b200_one_dim_rectilinear_{sfx}")
    _lib.check(
        fn(C.c_int(_lib.KINDS_1D[kind]), _p(grid, ct), C.c_size_t(grid.size), _p(vals, ct), C.c_size_t(vals.size),
           _p(locs, ct), C.c_size_t(locs.size), _p(out, ct), C.c_size_t(out.size))  # fmt: skip
    )
    return out
