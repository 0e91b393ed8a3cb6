"""The reference's Python-level API for the hot path, backed by the GPU library.

* ``interpn(obs, grids, vals, method=..., ...)`` — /root/reference/src/interpn/__init__.py:48-194
* ``MultilinearRegular`` / ``MultilinearRectilinear`` / ``MulticubicRegular`` /
  ``MulticubicRectilinear`` / ``NearestRegular`` / ``NearestRectilinear`` — the pydantic models of
  /root/reference/src/interpn/{multilinear,multicubic,nearest}_{regular,rectilinear}.py:
  same fields (so the JSON form round-trips between the two packages), same
  ``new / eval / eval_unchecked / check_bounds / ndims`` methods and assertion messages.

Unlike the reference (which rebuilds its stack-only Rust struct on every call), each model keeps
one grid-resident ``Interpolator`` alive after its first evaluation, so repeated ``eval`` calls
only move the query batch over PCIe.
"""

from __future__ import annotations

import json
import weakref
from collections.abc import Sequence
from functools import reduce
from typing import Annotated, Any, ClassVar, Literal

import numpy as np
from numpy.typing import NDArray
from pydantic import BaseModel, ConfigDict, Field, field_serializer, field_validator, model_validator

from . import raw
from .interpolator import Interpolator

# ---------------------------------------------------------------------------------------------
# JSON-compatible array wrappers (wire format of interpn/serialization.py:19-77:
# {"data": "<json list as string>", "dtype": "float64" | "float32"})
# ---------------------------------------------------------------------------------------------


def _array_model(name: str, np_dtype, tag: str):
    class _Arr(BaseModel):
        data: NDArray[np_dtype]  # type: ignore[valid-type]
        dtype: Literal[tag] = tag  # type: ignore[valid-type]

        model_config = ConfigDict(frozen=True, extra="forbid", arbitrary_types_allowed=True)

        @field_validator("data", mode="before")
        def _coerce(data: Any):
            if isinstance(data, str):
                data = json.loads(data)
            if isinstance(data, (list, np.ndarray)):
                return np.ascontiguousarray(np.asarray(data, dtype=np_dtype))
            raise TypeError

        @field_serializer("data", return_type=str)
        def _dump(data: Any) -> str:
            return json.dumps(data.tolist())

    _Arr.__name__ = _Arr.__qualname__ = name
    return _Arr


ArrayF64 = _array_model("ArrayF64", np.float64, "float64")
ArrayF32 = _array_model("ArrayF32", np.float32, "float32")
Array = Annotated[ArrayF32 | ArrayF64, Field(discriminator="dtype")]  # type: ignore[valid-type]


def _wrap(a: NDArray, dtype):
    return (ArrayF64 if dtype == np.float64 else ArrayF32)(data=np.asarray(a).flatten())


def _suffix(dtype) -> str:
    if dtype == np.float64:
        return "f64"
    if dtype == np.float32:
        return "f32"
    raise TypeError(f"Unexpected data type: {dtype}")


# ---------------------------------------------------------------------------------------------
# Interpolator models
# ---------------------------------------------------------------------------------------------


# Grid-resident interpolators of the live models, keyed by id(model). The device handle is deliberately NOT part of
# a model's state: like the reference's six classes, a model stays plain data — picklable, deep-copyable, equal to
# another model with the same fields — before and after its first evaluation. A copy or an unpickled model simply
# builds its own resident interpolator on first use; the entry is closed when its model is collected.
_RESIDENT: dict[int, Interpolator] = {}


def _drop_resident(key: int) -> None:
    it = _RESIDENT.pop(key, None)
    if it is not None:
        it.close()


class _Model(BaseModel):
    model_config = ConfigDict(frozen=True, extra="forbid", arbitrary_types_allowed=True)

    _method: ClassVar[str] = "linear"
    _max_dims: ClassVar[int] = 8

    vals: Array

    def _linearize(self) -> bool:
        return bool(getattr(self, "linearize_extrapolation", True))

    def _build(self) -> Interpolator:  # pragma: no cover - overridden
        raise NotImplementedError

    def _interp(self) -> Interpolator:
        key = id(self)
        it = _RESIDENT.get(key)
        if it is None:
            it = _RESIDENT[key] = self._build()
            weakref.finalize(self, _drop_resident, key)
        return it

    def eval(self, obs: list[NDArray], out: NDArray | None = None) -> NDArray:
        """Evaluate at the observation points (``obs`` = [x, y, ...] coordinate arrays), optionally
        into a preallocated ``out``. Inputs are not reallocated: non-contiguous arrays or a wrong
        dtype raise. Mirrors e.g. multilinear_regular.py:101-123."""
        out_inner = out if out is not None else np.zeros_like(obs[0])
        self.eval_unchecked(obs, out_inner)
        return out_inner

    def eval_unchecked(self, obs: list[NDArray], out: NDArray | None = None) -> NDArray:
        """Same as ``eval``; named for parity with the reference (e.g. multilinear_regular.py:125-166)."""
        dtype = self.vals.data.dtype
        _suffix(dtype)  # TypeError for anything but f32/f64
        out_inner = out if out is not None else np.zeros_like(obs[0])
        self._interp().eval(list(obs), out_inner)
        return out_inner


class _RegularModel(_Model):
    dims: list[int]
    starts: Array
    steps: Array

    @model_validator(mode="after")
    def _validate_model(self):
        ndims = self.ndims()
        assert 1 <= ndims <= self._max_dims, (
            f"Number of dimensions must be at least 1 and no more than {self._max_dims}"
        )
        assert self.starts.data.size == ndims, "Grid dimension mismatch"
        assert self.steps.data.size == ndims, "Grid dimension mismatch"
        assert self.vals.data.size == reduce(lambda acc, x: acc * x, self.dims), (
            "Size of value array does not match grid dims"
        )
        assert all(x > 0.0 for x in self.steps.data), "All grid steps must be positive and nonzero"
        assert all(x.data.dtype == self.vals.data.dtype for x in (self.starts, self.steps)), (
            "All grid inputs must be of the same data type (np.float32 or np.float64)"
        )
        assert all(x.data.data.contiguous for x in (self.starts, self.steps, self.vals)), (
            "Grid data must be contiguous"
        )
        return self

    def ndims(self) -> int:
        return len(self.dims)

    def _build(self) -> Interpolator:
        return Interpolator.regular(
            self._method, self.dims, self.starts.data, self.steps.data, self.vals.data, self._linearize()
        )

    def check_bounds(self, obs: list[NDArray], atol: float) -> NDArray[np.bool_]:
        """Per-dimension flag: True if any observation violates that axis' bounds by ``atol``."""
        out = np.array([False] * self.ndims())
        fn = getattr(raw, f"check_bounds_regular_{_suffix(self.vals.data.dtype)}")
        fn(self.dims, self.starts.data, self.steps.data, [np.asarray(x).flatten() for x in obs], atol, out)
        return out


class _RectilinearModel(_Model):
    grids: list[Array]

    @model_validator(mode="after")
    def _validate_model(self):
        ndims = self.ndims()
        dims = self.dims()
        assert 1 <= ndims <= self._max_dims, (
            f"Number of dimensions must be at least 1 and no more than {self._max_dims}"
        )
        assert self.vals.data.size == reduce(lambda acc, x: acc * x, dims), (
            "Size of value array does not match grid dims"
        )
        assert all(np.all(np.diff(x.data) > 0.0) for x in self.grids), (
            "All grids must be monotonically increasing"
        )
        assert all(x.data.dtype == self.vals.data.dtype for x in self.grids), (
            "All grid inputs must be of the same data type (np.float32 or np.float64)"
        )
        assert all(x.data.data.contiguous for x in self.grids) and self.vals.data.data.contiguous, (
            "Grid data must be contiguous"
        )
        return self

    def ndims(self) -> int:
        return len(self.grids)

    def dims(self) -> list[int]:
        return [x.data.size for x in self.grids]

    def _build(self) -> Interpolator:
        return Interpolator.rectilinear(self._method, [g.data for g in self.grids], self.vals.data, self._linearize())

    def check_bounds(self, obs: list[NDArray], atol: float) -> NDArray[np.bool_]:
        out = np.array([False] * self.ndims())
        fn = getattr(raw, f"check_bounds_rectilinear_{_suffix(self.vals.data.dtype)}")
        fn([g.data for g in self.grids], [np.asarray(x).flatten() for x in obs], atol, out)
        return out


class MultilinearRegular(_RegularModel):
    """Multilinear interpolation on a regular grid in up to 8 dimensions
    (multilinear_regular.py:24; Rust MultilinearRegular, multilinear/regular.rs:200)."""

    _method: ClassVar[str] = "linear"

    @classmethod
    def new(cls, dims: list[int], starts: NDArray, steps: NDArray, vals: NDArray) -> "MultilinearRegular":
        dt = vals.dtype
        return cls(dims=list(dims), starts=_wrap(starts, dt), steps=_wrap(steps, dt), vals=_wrap(vals, dt))


class NearestRegular(_RegularModel):
    """Nearest-neighbour interpolation on a regular grid in up to 6 dimensions (nearest_regular.py)."""

    _method: ClassVar[str] = "nearest"
    _max_dims: ClassVar[int] = 6

    @classmethod
    def new(cls, dims: list[int], starts: NDArray, steps: NDArray, vals: NDArray) -> "NearestRegular":
        dt = vals.dtype
        return cls(dims=list(dims), starts=_wrap(starts, dt), steps=_wrap(steps, dt), vals=_wrap(vals, dt))


class MulticubicRegular(_RegularModel):
    """Cubic Hermite interpolation on a regular grid in up to 8 dimensions (multicubic_regular.py)."""

    _method: ClassVar[str] = "cubic"
    linearize_extrapolation: bool

    @classmethod
    def new(cls, dims: list[int], starts: NDArray, steps: NDArray, vals: NDArray,
            linearize_extrapolation: bool = True) -> "MulticubicRegular":  # fmt: skip
        dt = vals.dtype
        return cls(dims=list(dims), starts=_wrap(starts, dt), steps=_wrap(steps, dt), vals=_wrap(vals, dt),
                   linearize_extrapolation=linearize_extrapolation)  # fmt: skip


class MultilinearRectilinear(_RectilinearModel):
    """Multilinear interpolation on a rectilinear grid in up to 8 dimensions (multilinear_rectilinear.py)."""

    _method: ClassVar[str] = "linear"

    @classmethod
    def new(cls, grids: list[NDArray], vals: NDArray) -> "MultilinearRectilinear":
        dt = vals.dtype
        return cls(grids=[_wrap(g, dt) for g in grids], vals=_wrap(vals, dt))


class NearestRectilinear(_RectilinearModel):
    """Nearest-neighbour interpolation on a rectilinear grid in up to 6 dimensions (nearest_rectilinear.py)."""

    _method: ClassVar[str] = "nearest"
    _max_dims: ClassVar[int] = 6

    @classmethod
    def new(cls, grids: list[NDArray], vals: NDArray) -> "NearestRectilinear":
        dt = vals.dtype
        return cls(grids=[_wrap(g, dt) for g in grids], vals=_wrap(vals, dt))


class MulticubicRectilinear(_RectilinearModel):
    """Cubic Hermite interpolation on a rectilinear grid in up to 8 dimensions (multicubic_rectilinear.py)."""

    _method: ClassVar[str] = "cubic"
    linearize_extrapolation: bool

    @classmethod
    def new(cls, grids: list[NDArray], vals: NDArray, linearize_extrapolation: bool = True) -> "MulticubicRectilinear":
        dt = vals.dtype
        return cls(grids=[_wrap(g, dt) for g in grids], vals=_wrap(vals, dt),
                   linearize_extrapolation=linearize_extrapolation)  # fmt: skip


# ---------------------------------------------------------------------------------------------
# interpn() convenience function
# ---------------------------------------------------------------------------------------------


def _check_regular(grids: Sequence[NDArray]) -> bool:
    """O(grid) regularity test (interpn/__init__.py:197-203): every spacing equals the first."""
    for grid in grids:
        d = np.diff(grid)
        if not np.all(d == d[0]):
            return False
    return True


def interpn(
    obs: Sequence[NDArray],
    grids: Sequence[NDArray],
    vals: NDArray,
    *,
    method: Literal["linear", "cubic", "nearest"] = "linear",
    out: NDArray | None = None,
    linearize_extrapolation: bool = True,
    assume_regular: bool = False,
    check_bounds: bool = False,
    bounds_atol: float = 1e-8,
) -> NDArray:
    """Evaluate an N-dimensional grid at the observation points (interpn/__init__.py:48-194).

    Same keyword surface and dispatch as the reference. Two of its quirks are fixed rather than
    reproduced: a preallocated multi-element ``out`` is accepted, and
    ``check_bounds`` works for any dimensionality.
    """
    out = out if out is not None else np.zeros_like(obs[0])
    outshape = out.shape
    out = out.reshape(-1)

    obs = [np.ascontiguousarray(np.asarray(x).ravel()) for x in obs]
    grids = [np.ascontiguousarray(np.asarray(x).ravel()) for x in grids]
    vals = np.ascontiguousarray(np.asarray(vals).ravel())

    dtype = vals.dtype
    assert dtype in [np.float64, np.float32], "`interpn` defined only for float32 and float64 data"
    sfx = _suffix(dtype)
    if method not in ("linear", "cubic", "nearest"):
        raise ValueError(f"Unsupported interpolation configuration: {dtype}, {assume_regular}, {method}")

    is_regular = assume_regular or _check_regular(grids)
    if is_regular:
        dims = [len(g) for g in grids]
        starts = np.array([g[0] for g in grids], dtype=dtype)
        steps = np.array([g[1] - g[0] for g in grids], dtype=dtype)

    if check_bounds:
        outb = np.zeros(len(grids), dtype=bool)
        if is_regular:
            getattr(raw, f"check_bounds_regular_{sfx}")(dims, starts, steps, obs, bounds_atol, outb)
        else:
            getattr(raw, f"check_bounds_rectilinear_{sfx}")(grids, obs, bounds_atol, outb)
        if any(outb):
            raise ValueError("Observation points violate interpolator bounds")

    if is_regular:
        fn = getattr(raw, f"interpn_{method}_regular_{sfx}")
        if method == "cubic":
            fn(dims, starts, steps, vals, linearize_extrapolation, obs, out)
        else:
            fn(dims, starts, steps, vals, obs, out)
    else:
        fn = getattr(raw, f"interpn_{method}_rectilinear_{sfx}")
        if method == "cubic":
            fn(grids, vals, linearize_extrapolation, obs, out)
        else:
            fn(grids, vals, obs, out)

    return out.reshape(outshape)
