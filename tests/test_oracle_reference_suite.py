"""Pins the CPU oracle against every known-answer / analytic test the reference holds for the
interpolation hot path, in both arithmetic modes (strict = crate default,
fma = `--features=fma`, mirroring .github/workflows/test-rust.yml:32-36) and in both evaluation
orders (flattened / recursive). CPU only."""

import numpy as np
import pytest

from tests import refsuite as rs


class OracleEngine:
    def __init__(self, oracle, fma=False, order="reference"):
        self.o, self.fma, self.order = oracle, fma, order

    def regular(self, method, dims, starts, steps, vals, obs, linearize=True):
        return self.o.interpn_regular(
            method, dims, starts, steps, vals, obs, linearize_extrapolation=linearize, fma=self.fma, order=self.order
        )

    def rectilinear(self, method, grids, vals, obs, linearize=True):
        return self.o.interpn_rectilinear(
            method, grids, vals, obs, linearize_extrapolation=linearize, fma=self.fma, order=self.order
        )

    def one_dim_regular(self, kind, start, step, vals, locs):
        return self.o.one_dim_regular(kind, start, step, vals, locs, fma=self.fma)

    def one_dim_rectilinear(self, kind, grid, vals, locs):
        return self.o.one_dim_rectilinear(kind, grid, vals, locs, fma=self.fma)

    def check_bounds_regular(self, dims, starts, steps, obs, atol):
        return self.o.check_bounds_regular(dims, starts, steps, obs, atol)

    def check_bounds_rectilinear(self, grids, obs, atol):
        return self.o.check_bounds_rectilinear(grids, obs, atol)


MODES = [(False, "reference"), (True, "reference"), (False, "flattened"), (False, "recursive"), (True, "recursive")]


@pytest.fixture(params=MODES, ids=lambda m: f"{'fma' if m[0] else 'strict'}-{m[1]}")
def engine(request, oracle):
    return OracleEngine(oracle, *request.param)


@pytest.mark.parametrize("ndims", range(1, 9))
def test_linear_regular_field(engine, ndims):
    rs.check_linear_regular_field(engine, ndims)


def test_linear_regular_hat(engine):
    rs.check_linear_regular_hat(engine)


def test_linear_rect_2d_small(engine):
    rs.check_linear_rect_2d_small(engine)


@pytest.mark.parametrize("ndims", range(1, 9))
def test_linear_rect_field(engine, ndims):
    rs.check_linear_rect_field(engine, ndims)


def test_linear_rect_hat(engine):
    rs.check_linear_rect_hat(engine)


@pytest.mark.parametrize("rect", [False, True], ids=["regular", "rectilinear"])
@pytest.mark.parametrize("ndims", range(1, 6))
def test_cubic_linear_field(engine, ndims, rect):
    rs.check_cubic_linear_field(engine, ndims, rect)


@pytest.mark.parametrize("rect", [False, True], ids=["regular", "rectilinear"])
@pytest.mark.parametrize("ndims", range(1, 6))
def test_cubic_quadratic_field(engine, ndims, rect):
    rs.check_cubic_quadratic_field(engine, ndims, rect)


@pytest.mark.parametrize("rect", [False, True], ids=["regular", "rectilinear"])
@pytest.mark.parametrize("ndims", [1, 2])
def test_cubic_sine(engine, ndims, rect):
    rs.check_cubic_sine(engine, ndims, rect)


@pytest.mark.parametrize("ndims", range(1, 7))
def test_nearest_regular_field(engine, ndims):
    rs.check_nearest_regular_field(engine, ndims)


def test_nearest_regular_hat(engine):
    rs.check_nearest_regular_hat(engine)


def test_nearest_rect_2d_small(engine):
    rs.check_nearest_rect_2d_small(engine)


@pytest.mark.parametrize("ndims", range(1, 7))
def test_nearest_rect_field(engine, ndims):
    rs.check_nearest_rect_field(engine, ndims)


def test_nearest_rect_hat(engine):
    rs.check_nearest_rect_hat(engine)


def test_one_dim_linear(engine):
    rs.check_one_dim_linear(engine)


def test_one_dim_hold(engine):
    rs.check_one_dim_hold(engine)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_python_suite(engine, dtype):
    rs.check_py_linear_regular_nodes(engine, dtype)
    rs.check_py_linear_rect_nodes(engine, dtype)
    rs.check_py_cubic_regular_nodes(engine, dtype)
    rs.check_py_cubic_rect_nodes(engine, dtype)
    rs.check_py_nearest_regular(engine, dtype)
    rs.check_py_nearest_rect(engine, dtype)
    rs.check_py_check_bounds(engine, dtype)


def test_cubic_6d_rect_linear_field_recursive(oracle):
    """rectilinear_recursive.rs:551-599 runs N=1..=6; N=6 (6^6 points x 4^6 vertices) once, strict."""
    rs.check_cubic_linear_field(OracleEngine(oracle, False, "reference"), 6, True)


# ---- oracle self-checks (SURVEY.md §8c "Oracle self-check") -----------------------------------


def _random_case(rng, ndims, n, lo, hi, dtype=np.float64):
    dims = [int(rng.integers(lo, hi)) for _ in range(ndims)]
    grids = [np.sort(rng.random(d) * 10.0 - 5.0).astype(dtype) for d in dims]
    for g in grids:
        assert np.all(np.diff(g) > 0)
    vals = rng.standard_normal(int(np.prod(dims))).astype(dtype)
    obs = [(rng.random(n) * 14.0 - 7.0).astype(dtype) for _ in range(ndims)]
    starts = np.array([g[0] for g in grids], dtype=dtype)
    steps = np.array([(g[-1] - g[0]) / (len(g) - 1) for g in grids], dtype=dtype)
    return dims, grids, starts, steps, vals, obs


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("method,maxn", [("linear", 7), ("cubic", 5)])
def test_flattened_and_recursive_orders_agree_bitwise(oracle, method, maxn, dtype):
    """Strict mode: both evaluation orders visit the same arithmetic DAG -> identical bits."""
    rng = np.random.default_rng(7)
    lo = 4 if method == "cubic" else 2
    for ndims in range(1, maxn + 1):
        dims, grids, starts, steps, vals, obs = _random_case(rng, ndims, 300, lo, lo + 3, dtype)
        for lin in (False, True):
            a = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=lin, order="flattened")
            b = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=lin, order="recursive")
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
            a = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=lin, order="flattened")
            b = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=lin, order="recursive")
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


@pytest.mark.parametrize("fma", [False, True])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("method,maxn", [("linear", 6), ("cubic", 4)])
def test_const_dimension_fast_path_equals_the_runtime_dimension_twin(oracle, method, maxn, dtype, fma):
    """order="reference" runs the flattened range (linear N<=6, cubic N<=4) with the dimensionality as a compile-time
    constant (the timed CPU baseline, like the crate's const-generic structs); order="flattened" keeps the runtime-N
    loops. Same operations, same order: identical bits in both arithmetic modes."""
    rng = np.random.default_rng(17)
    lo = 4 if method == "cubic" else 2
    for ndims in range(1, maxn + 1):
        dims, grids, starts, steps, vals, obs = _random_case(rng, ndims, 500, lo, lo + 3, dtype)
        for lin in (False, True):
            a = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=lin, fma=fma, order="reference")
            b = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=lin, fma=fma, order="flattened")
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))
            a = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=lin, fma=fma, order="reference")
            b = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=lin, fma=fma, order="flattened")
            assert np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_threads_match_serial(oracle):
    rng = np.random.default_rng(11)
    dims, grids, starts, steps, vals, obs = _random_case(rng, 3, 10007, 5, 9)
    a = oracle.interpn_regular("cubic", dims, starts, steps, vals, obs, nthreads=1)
    b = oracle.interpn_regular("cubic", dims, starts, steps, vals, obs, nthreads=5)
    assert np.array_equal(a, b)
    a = oracle.interpn_rectilinear("linear", grids, vals, obs, nthreads=1)
    b = oracle.interpn_rectilinear("linear", grids, vals, obs, nthreads=3)
    assert np.array_equal(a, b)


def test_fma_mode_stays_within_a_few_ulp_on_smooth_data(oracle):
    """CHANGELOG.md:114-118: fma changes roundoff by 0-4 eps inside the grid."""
    rng = np.random.default_rng(3)
    x = np.linspace(0.0, 1.0, 12)
    g = np.meshgrid(x, x, x, indexing="ij")
    vals = (1.0 + g[0] + 2 * g[1] * g[1] + np.sin(g[2])).ravel()
    obs = [rng.random(2000) * 0.8 + 0.1 for _ in range(3)]
    starts, steps = np.zeros(3), np.full(3, x[1] - x[0])
    a = oracle.interpn_regular("cubic", [12] * 3, starts, steps, vals, obs, fma=False)
    b = oracle.interpn_regular("cubic", [12] * 3, starts, steps, vals, obs, fma=True)
    assert np.max(np.abs(a - b) / np.abs(a)) < 16 * np.finfo(np.float64).eps
