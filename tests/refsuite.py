"""Restatement of the reference's own test suite for the interpolation hot path.

Each ``check_*`` function re-expresses one reference test (cited by file:line, relative to
/root/reference) against an *engine* — any object with the methods of :class:`Engine`. The same
checks are run against the CPU oracle (tests/test_oracle_reference_suite.py, no GPU) and the CUDA
library through its reference-shaped Python API (tests/test_gpu_reference_suite.py, -m gpu).

The reference's tests hold no stored vectors: they are analytic/known-answer checks,
so re-running them is how the oracle is pinned. The reference's test RNG (ChaCha12 `StdRng`,
src/testing.rs:7-25) cannot be reproduced without the `rand` crate; jittered grids use numpy's
seeded generator instead (same distribution: uniform [0,1)).
"""

from __future__ import annotations

import itertools
from typing import Protocol, Sequence

import numpy as np


class Engine(Protocol):
    def regular(self, method: str, dims, starts, steps, vals, obs, linearize: bool = True) -> np.ndarray: ...
    def rectilinear(self, method: str, grids, vals, obs, linearize: bool = True) -> np.ndarray: ...
    def one_dim_regular(self, kind: str, start, step, vals, locs) -> np.ndarray: ...
    def one_dim_rectilinear(self, kind: str, grid, vals, locs) -> np.ndarray: ...
    def check_bounds_regular(self, dims, starts, steps, obs, atol) -> np.ndarray: ...
    def check_bounds_rectilinear(self, grids, obs, atol) -> np.ndarray: ...


# --------------------------------------------------------------------------------------------
# Helpers restating src/utils.rs
# --------------------------------------------------------------------------------------------


def linspace(start: float, stop: float, n: int, dtype=np.float64) -> np.ndarray:
    """src/utils.rs:8-14: dx = (stop-start)/(n-1); x[i] = start + i*dx (no endpoint fix-up)."""
    dt = np.dtype(dtype).type
    dx = (dt(stop) - dt(start)) / dt(n - 1)
    return np.array([dt(start) + dt(i) * dx for i in range(n)], dtype=dtype)


def meshgrid(axes: Sequence[np.ndarray]) -> np.ndarray:
    """src/utils.rs:17-25: C-ordered cartesian product, shape (prod, ndims)."""
    return np.array(list(itertools.product(*[list(a) for a in axes])), dtype=np.float64).reshape(-1, len(axes))


def rng_fixed_seed() -> np.random.Generator:
    return np.random.default_rng(20260117)


def jittered_axis(rng, i: int, n: int) -> np.ndarray:
    """linspace + (u-0.5)/10 jitter, e.g. multilinear/rectilinear.rs:422-429."""
    x = linspace(-5.0 * i, 5.0 * (i + 1), n)
    x = x + (rng.random(n) - 0.5) / 10.0
    assert np.all(np.diff(x) > 0)
    return x


def hat_func(x: float) -> float:
    return x if x <= 1.0 else 2.0 - x


def _obs_from(axes: Sequence[np.ndarray]):
    g = meshgrid(axes)
    return g, [np.ascontiguousarray(g[:, i]) for i in range(len(axes))]


# --------------------------------------------------------------------------------------------
# Multilinear
# --------------------------------------------------------------------------------------------


def check_linear_regular_field(e: Engine, ndims: int):
    """multilinear/regular.rs:437-477 (N=1..6) and regular_recursive.rs:401-441 (N=1..8)."""
    dims = [2] * ndims
    xs = [linspace(-5.0 * i, 5.0 * (i + 1), dims[i]) for i in range(ndims)]
    u = meshgrid(xs).sum(axis=1)
    starts = np.array([x[0] for x in xs])
    steps = np.array([x[1] - x[0] for x in xs])
    gridobs, obs = _obs_from([linspace(-7.0 * i, 7.0 * (i + 1), 3) for i in range(ndims)])
    uobs = gridobs.sum(axis=1)
    out = e.regular("linear", dims, starts, steps, u, obs)
    assert np.all(np.abs(out - uobs) < 1e-12)


def check_linear_regular_hat(e: Engine):
    """multilinear/regular.rs:480-495 — assert_eq! (bit-exact)."""
    y = np.array([hat_func(float(i)) for i in range(3)])
    obs = linspace(-2.0, 4.0, 100)
    out = e.regular("linear", [3], np.array([0.0]), np.array([1.0]), y, [obs])
    expect = np.array([hat_func(v) for v in obs])
    assert np.array_equal(out, expect)


def check_linear_rect_2d_small(e: Engine):
    """multilinear/rectilinear.rs:380-407."""
    x = linspace(-1.0, 1.0, 3)
    y = np.array([0.5, 0.6])
    z = meshgrid([x, y]).sum(axis=1)
    gridobs, obs = _obs_from([linspace(-10.0, 10.0, 5), linspace(-10.0, 10.0, 5)])
    out = e.rectilinear("linear", [x, y], z, obs)
    assert np.all(np.abs(out - gridobs.sum(axis=1)) < 1e-12)


def check_linear_rect_field(e: Engine, ndims: int):
    """multilinear/rectilinear.rs:413-456 (N=1..6), rectilinear_recursive.rs:379-422 (N=1..8)."""
    rng = rng_fixed_seed()
    xs = [jittered_axis(rng, i, 2) for i in range(ndims)]
    u = meshgrid(xs).sum(axis=1)
    gridobs, obs = _obs_from([linspace(-7.0 * i, 7.0 * (i + 1), 3) for i in range(ndims)])
    out = e.rectilinear("linear", xs, u, obs)
    assert np.all(np.abs(out - gridobs.sum(axis=1)) < 1e-12)


def check_linear_rect_hat(e: Engine):
    """multilinear/rectilinear.rs:459-476 — assert_eq! (bit-exact)."""
    x = np.array([0.0, 1.0, 2.0])
    y = np.array([hat_func(v) for v in x])
    obs = linspace(-2.0, 4.0, 100)
    out = e.rectilinear("linear", [x], y, [obs])
    assert np.array_equal(out, np.array([hat_func(v) for v in obs]))


# --------------------------------------------------------------------------------------------
# Multicubic
# --------------------------------------------------------------------------------------------


def _cubic_field_case(ndims: int, fn, rect: bool, nobs_extra: int = 2, ngrid: int = 4, obs_span=7.0):
    rng = rng_fixed_seed()
    dims = [ngrid] * ndims
    if rect:
        xs = [jittered_axis(rng, i, dims[i]) for i in range(ndims)]
    else:
        xs = [linspace(-5.0 * i, 5.0 * (i + 1), dims[i]) for i in range(ndims)]
    u = fn(meshgrid(xs))
    gridobs, obs = _obs_from(
        [linspace(-obs_span * i, obs_span * (i + 1), dims[i] + nobs_extra) for i in range(ndims)]
    )
    return dims, xs, u, obs, fn(gridobs)


def _eval_cubic(e: Engine, rect: bool, dims, xs, u, obs, linearize: bool):
    if rect:
        return e.rectilinear("cubic", xs, u, obs, linearize)
    starts = np.array([x[0] for x in xs])
    steps = np.array([x[1] - x[0] for x in xs])
    return e.regular("cubic", dims, starts, steps, u, obs, linearize)


def check_cubic_linear_field(e: Engine, ndims: int, rect: bool):
    """multicubic/regular.rs:634-676 (<1e-12), rectilinear.rs:557-604 (<1e-10), and recursive twins
    regular_recursive.rs:621-664, rectilinear_recursive.rs:551-599; both linearize settings."""
    dims, xs, u, obs, uobs = _cubic_field_case(ndims, lambda g: g.sum(axis=1), rect)
    tol = 1e-10 if rect else 1e-12
    for linearize in (False, True):
        out = _eval_cubic(e, rect, dims, xs, u, obs, linearize)
        assert np.all(np.abs(out - uobs) < tol)


def check_cubic_quadratic_field(e: Engine, ndims: int, rect: bool):
    """multicubic/regular.rs:680-730, rectilinear.rs:608-666 (+ recursive twins): a quadratic is
    reproduced to 1e-10 under interpolation and (non-linearized) extrapolation."""
    dims, xs, u, obs, uobs = _cubic_field_case(ndims, lambda g: (g * g).sum(axis=1), rect)
    out = _eval_cubic(e, rect, dims, xs, u, obs, False)
    assert np.all(np.abs(out - uobs) < 1e-10)


def check_cubic_sine(e: Engine, ndims: int, rect: bool):
    """multicubic/regular.rs:736-792, rectilinear.rs:672-736: sine within 2e-2*ndims (interp only)."""
    fn = lambda g: np.sin(g * 6.28 / 10.0).sum(axis=1)  # noqa: E731
    dims, xs, u, obs, uobs = _cubic_field_case(
        ndims, fn, rect, nobs_extra=1 if rect else 2, ngrid=10, obs_span=5.0
    )
    out = _eval_cubic(e, rect, dims, xs, u, obs, False)
    assert np.all(np.abs(out - uobs) < 2e-2 * ndims)


# --------------------------------------------------------------------------------------------
# Nearest
# --------------------------------------------------------------------------------------------


def nearest_regular_index(value: float, start: float, step: float, dim: int) -> int:
    """nearest/regular.rs:324-337 (in-test restatement of the index rule)."""
    floc = np.floor((value - start) / step)
    dimmax = max(dim - 2, 0)
    origin = int(min(max(int(floc), 0), dimmax))
    index_zero = start + step * float(origin)
    dt = (value - index_zero) / step
    return origin if dt <= 0.5 else min(origin + 1, dim - 1)


def nearest_rectilinear_index(value: float, grid: np.ndarray) -> int:
    """nearest/rectilinear.rs:274-283."""
    iloc = int(np.searchsorted(grid, value, side="left")) - 1
    dimmax = max(len(grid) - 2, 0)
    origin = min(max(iloc, 0), dimmax)
    x0, x1 = grid[origin], grid[origin + 1]
    dt = (value - x0) / (x1 - x0)
    return origin if dt <= 0.5 else origin + 1


def check_nearest_regular_field(e: Engine, ndims: int):
    """nearest/regular.rs:343-398."""
    dims = [2] * ndims
    xs = [linspace(-5.0 * i, 5.0 * (i + 1), dims[i]) for i in range(ndims)]
    u = meshgrid(xs).sum(axis=1)
    starts = np.array([x[0] for x in xs])
    steps = np.array([x[1] - x[0] for x in xs])
    gridobs, obs = _obs_from([linspace(-7.0 * i, 7.0 * (i + 1), 3) for i in range(ndims)])
    expected = np.array(
        [
            sum(
                starts[d] + steps[d] * float(nearest_regular_index(p[d], starts[d], steps[d], dims[d]))
                for d in range(ndims)
            )
            for p in gridobs
        ]
    )
    out = e.regular("nearest", dims, starts, steps, u, obs)
    assert np.all(np.abs(out - expected) < 1e-12)


def check_nearest_regular_hat(e: Engine):
    """nearest/regular.rs:401-417 — assert_eq!."""
    y = np.array([hat_func(float(i)) for i in range(3)])
    obs = linspace(-2.0, 4.0, 100)
    out = e.regular("nearest", [3], np.array([0.0]), np.array([1.0]), y, [obs])
    expect = np.array([y[nearest_regular_index(v, 0.0, 1.0, 3)] for v in obs])
    assert np.array_equal(out, expect)


def check_nearest_rect_2d_small(e: Engine):
    """nearest/rectilinear.rs:286-312."""
    x = linspace(-1.0, 1.0, 3)
    y = np.array([0.5, 0.6])
    z = meshgrid([x, y]).sum(axis=1)
    gridobs, obs = _obs_from([linspace(-10.0, 10.0, 5), linspace(-10.0, 10.0, 5)])
    out = e.rectilinear("nearest", [x, y], z, obs)
    expected = np.array([x[nearest_rectilinear_index(p[0], x)] + y[nearest_rectilinear_index(p[1], y)] for p in gridobs])
    assert np.all(np.abs(out - expected) < 1e-12)


def check_nearest_rect_field(e: Engine, ndims: int):
    """nearest/rectilinear.rs:318-370."""
    rng = rng_fixed_seed()
    xs = [jittered_axis(rng, i, 2) for i in range(ndims)]
    u = meshgrid(xs).sum(axis=1)
    gridobs, obs = _obs_from([linspace(-7.0 * i, 7.0 * (i + 1), 3) for i in range(ndims)])
    expected = np.array(
        [sum(xs[d][nearest_rectilinear_index(p[d], xs[d])] for d in range(ndims)) for p in gridobs]
    )
    out = e.rectilinear("nearest", xs, u, obs)
    assert np.all(np.abs(out - expected) < 1e-12)


def check_nearest_rect_hat(e: Engine):
    """nearest/rectilinear.rs:373-391 — assert_eq!."""
    x = np.array([0.0, 1.0, 2.0])
    y = np.array([hat_func(v) for v in x])
    obs = linspace(-2.0, 4.0, 100)
    out = e.rectilinear("nearest", [x], y, [obs])
    expect = np.array([y[nearest_rectilinear_index(v, x)] for v in obs])
    assert np.array_equal(out, expect)


# --------------------------------------------------------------------------------------------
# one_dim
# --------------------------------------------------------------------------------------------


def _one_dim_case():
    rng = rng_fixed_seed()
    n = 77
    vals = rng.random(n)
    start, stop = -3.14, 314.0
    x_reg = linspace(start, stop, n)
    x_rect = np.sort(rng.random(n)) * (stop - start) + start
    locs = rng.random(3 * n) * 2.0 * (stop - start) + 2.0 * start
    return n, vals, x_reg, x_rect, locs


def check_one_dim_linear(e: Engine):
    """one_dim/linear.rs:96-179."""
    n, vals, x_reg, x_rect, locs = _one_dim_case()
    step = x_reg[1] - x_reg[0]
    results = [
        (x_reg, e.one_dim_regular("linear", x_reg[0], step, vals, locs), False),
        (x_rect, e.one_dim_rectilinear("linear", x_rect, vals, locs), False),
        (x_reg, e.one_dim_regular("linear_hold_last", x_reg[0], step, vals, locs), True),
        (x_rect, e.one_dim_rectilinear("linear_hold_last", x_rect, vals, locs), True),
    ]
    for xs, ys, hold in results:
        for loc, y in zip(locs, ys):
            j = min(max(int(np.searchsorted(xs, loc, side="left")) - 1, 0), n - 2)
            xl, xr, yl, yr = xs[j], xs[j + 1], vals[j], vals[j + 1]
            slope = (yr - yl) / (xr - xl)
            dx = loc - xl
            if xs[0] <= loc <= xs[n - 1]:
                assert min(yl, yr) <= y <= max(yl, yr)
                assert xl <= loc <= xr
            elif loc > xs[n - 1] and hold:
                assert abs((y - vals[n - 1]) / vals[n - 1]) < 1e-12
                continue
            elif loc < xs[0] and hold:
                assert abs((y - vals[0]) / vals[0]) < 1e-12
                continue
            y_expected = yl + slope * dx
            assert abs((y - y_expected) / y_expected) < 1e-12


def check_one_dim_hold(e: Engine):
    """one_dim/hold.rs:118-179 — assert_eq! on Left/Right/Nearest."""
    n, vals, x_reg, _, locs = _one_dim_case()
    step = x_reg[1] - x_reg[0]
    y_l = e.one_dim_regular("left", x_reg[0], step, vals, locs)
    y_r = e.one_dim_regular("right", x_reg[0], step, vals, locs)
    y_n = e.one_dim_regular("nearest", x_reg[0], step, vals, locs)
    for i, loc in enumerate(locs):
        j = min(max(int(np.searchsorted(x_reg, loc, side="left")) - 1, 0), n - 2)
        xl, xr, yl, yr = x_reg[j], x_reg[j + 1], vals[j], vals[j + 1]
        if x_reg[0] <= loc <= x_reg[n - 1]:
            assert xl <= loc <= xr
            assert y_l[i] == yl and y_r[i] == yr
        elif loc > x_reg[n - 1]:
            assert y_l[i] == yr and y_r[i] == yr
        else:
            assert y_l[i] == yl and y_r[i] == yl
        assert y_n[i] == (yl if (loc - xl) <= (xr - loc) else yr)


# --------------------------------------------------------------------------------------------
# Python integration tests (test/test_*.py): evaluation at the grid nodes of z = x + 2y
# --------------------------------------------------------------------------------------------


def _py_case(dtype, nx, ny):
    x = np.linspace(0.0, 10.0, nx).astype(dtype)
    y = np.linspace(20.0, 30.0, ny).astype(dtype)
    xg, yg = np.meshgrid(x, y, indexing="ij")
    z = (xg + 2.0 * yg).astype(dtype)
    dims = [x.size, y.size]
    starts = np.array([x[0], y[0]]).astype(dtype)
    steps = np.array([x[1] - x[0], y[1] - y[0]]).astype(dtype)
    obs = [xg.flatten().astype(dtype), yg.flatten().astype(dtype)]
    return x, y, z.flatten(), dims, starts, steps, obs


def check_py_linear_regular_nodes(e: Engine, dtype):
    """test/test_multilinear_regular.py:6-48 — exact equality at nodes."""
    x, y, zf, dims, starts, steps, obs = _py_case(dtype, 5, 3)
    out = e.regular("linear", dims, starts, steps, zf, obs)
    assert out.dtype == dtype and np.array_equal(out, zf)


def check_py_linear_rect_nodes(e: Engine, dtype):
    """test/test_multilinear_rectilinear.py:6-41 — exact equality at nodes."""
    x, y, zf, *_, obs = _py_case(dtype, 5, 3)
    out = e.rectilinear("linear", [x, y], zf, obs)
    assert out.dtype == dtype and np.array_equal(out, zf)


def check_py_cubic_regular_nodes(e: Engine, dtype):
    """test/test_multicubic_regular.py:6-100 — tol 1e-12 (f64) / 1e-6 (f32), linearize False."""
    tol = 1e-12 if dtype == np.float64 else 1e-6
    x, y, zf, dims, starts, steps, obs = _py_case(dtype, 7, 5)
    out = e.regular("cubic", dims, starts, steps, zf, obs, False)
    assert np.all(np.abs(out - zf) / np.maximum(np.abs(zf), 1.0) < tol)


def check_py_cubic_rect_nodes(e: Engine, dtype):
    """test/test_multicubic_rectilinear.py:6-43 — exact equality at nodes, linearize False."""
    x, y, zf, *_, obs = _py_case(dtype, 5, 4)
    out = e.rectilinear("cubic", [x, y], zf, obs, False)
    assert out.dtype == dtype and np.array_equal(out, zf)


def check_py_nearest_regular(e: Engine, dtype):
    """test/test_nearest_regular.py:13-61 — assert_array_equal vs the Python restatement."""
    x = np.linspace(0.0, 6.0, 4).astype(dtype)
    y = np.linspace(-3.0, 3.0, 3).astype(dtype)
    xg, yg = np.meshgrid(x, y, indexing="ij")
    z = (xg - 2.0 * yg).astype(dtype)
    dims = [x.size, y.size]
    starts = np.array([x[0], y[0]]).astype(dtype)
    steps = np.array([x[1] - x[0], y[1] - y[0]]).astype(dtype)
    obs = [np.array([0.1, 1.6, 2.9, 5.0], dtype=dtype), np.array([-3.0, -1.2, 0.4, 2.4], dtype=dtype)]
    out = e.regular("nearest", dims, starts, steps, z.flatten(), obs)

    def idx(value, start, step, size):  # test/test_nearest_regular.py:5-10
        loc = int(max(0, min(np.floor((value - start) / step), size - 2)))
        dt = (value - (start + step * loc)) / step
        return min(loc if dt <= 0.5 else loc + 1, size - 1)

    expected = np.array(
        [
            z[idx(float(a), float(starts[0]), float(steps[0]), dims[0]), idx(float(b), float(starts[1]), float(steps[1]), dims[1])]
            for a, b in zip(obs[0], obs[1])
        ],
        dtype=dtype,
    )
    assert np.array_equal(out, expected)


def check_py_nearest_rect(e: Engine, dtype):
    """test/test_nearest_rectilinear.py:14-60."""
    x = np.array([0.0, 1.0, 3.5, 4.0], dtype=dtype)
    y = np.array([-2.0, -0.5, 0.1], dtype=dtype)
    xg, yg = np.meshgrid(x, y, indexing="ij")
    z = (xg + yg**2).astype(dtype)
    obs = [np.array([0.2, 2.8, 3.8], dtype=dtype), np.array([-1.5, -0.2, 0.4], dtype=dtype)]
    out = e.rectilinear("nearest", [x, y], z.flatten(), obs)

    def idx(value, grid):  # test/test_nearest_rectilinear.py:5-11
        i = int(max(0, min(np.searchsorted(grid, value, side="right") - 1, grid.size - 2)))
        dt = (value - grid[i]) / (grid[i + 1] - grid[i])
        return i if dt <= 0.5 else i + 1

    expected = np.array([z[idx(a, x), idx(b, y)] for a, b in zip(obs[0], obs[1])], dtype=dtype)
    assert np.array_equal(out, expected)


def check_py_check_bounds(e: Engine, dtype):
    """test/test_multilinear_regular.py:70-81, test_multilinear_rectilinear.py (same block),
    test/test_interpn.py:8-58."""
    x, y, zf, dims, starts, steps, obs = _py_case(dtype, 5, 3)
    inside = [np.array([5.0], dtype=dtype), np.array([25.0], dtype=dtype)]
    outside = [np.array([-5.0], dtype=dtype), np.array([-25.0], dtype=dtype)]
    atol = dtype(1e-6)
    assert not e.check_bounds_regular(dims, starts, steps, inside, atol).any()
    assert e.check_bounds_regular(dims, starts, steps, outside, atol).any()
    assert not e.check_bounds_rectilinear([x, y], inside, atol).any()
    assert e.check_bounds_rectilinear([x, y], outside, atol).any()
    grid = np.linspace(-1.0, 1.0, 5).astype(dtype)
    st, sp = np.array([grid[0]], dtype=dtype), np.array([grid[1] - grid[0]], dtype=dtype)
    assert not e.check_bounds_regular([5], st, sp, [np.array([-0.5, 0.5], dtype=dtype)], dtype(1e-8)).any()
    assert e.check_bounds_regular([5], st, sp, [np.array([-0.5, 1.5], dtype=dtype)], dtype(1e-8)).all()
    g2 = np.array([-1.0, -0.25, 0.5, 2.0], dtype=dtype)
    assert not e.check_bounds_rectilinear([g2], [np.array([-0.5, 1.0], dtype=dtype)], dtype(1e-8)).any()
    assert e.check_bounds_rectilinear([g2], [np.array([-1.5, 0.25], dtype=dtype)], dtype(1e-8)).all()
