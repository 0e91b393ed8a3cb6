"""The reference's own test suite (tests/refsuite.py) run against the CUDA library through the
C ABI with host buffers. Needs a B200: `pytest -m gpu`."""

import numpy as np
import pytest

from tests import refsuite as rs
from tests.gpu_engine import GpuEngine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    return GpuEngine()


@pytest.mark.parametrize("ndims", range(1, 9))
def test_linear_regular_field(engine, ndims):
    rs.check_linear_regular_field(engine, ndims)


def test_linear_hats_and_small(engine):
    rs.check_linear_regular_hat(engine)
    rs.check_linear_rect_hat(engine)
    rs.check_linear_rect_2d_small(engine)


@pytest.mark.parametrize("ndims", range(1, 9))
def test_linear_rect_field(engine, ndims):
    rs.check_linear_rect_field(engine, ndims)


@pytest.mark.parametrize("rect", [False, True], ids=["regular", "rectilinear"])
@pytest.mark.parametrize("ndims", range(1, 7))
def test_cubic_linear_field(engine, ndims, rect):
    if ndims == 6 and not rect:
        # the reference stops at N=5 for the regular grid (regular_recursive.rs:623 `1..6`); its 1e-12
        # bound does not hold at N=6 for the reference arithmetic itself (the oracle gives 5.6e-12)
        pytest.skip("not asserted by the reference")
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
def test_nearest_fields(engine, ndims):
    rs.check_nearest_regular_field(engine, ndims)
    rs.check_nearest_rect_field(engine, ndims)


def test_nearest_hats_and_small(engine):
    rs.check_nearest_regular_hat(engine)
    rs.check_nearest_rect_hat(engine)
    rs.check_nearest_rect_2d_small(engine)


def test_one_dim(engine):
    rs.check_one_dim_linear(engine)
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
