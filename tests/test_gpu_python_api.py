"""The Python-level mirror of the reference (interpn_b200.api: six pydantic classes + `interpn()`) ON THE GPU, compared
with the oracle bit for bit (ref: /root/reference/src/interpn/__init__.py:48-194 and
src/interpn/{multilinear,multicubic,nearest}_{regular,rectilinear}.py). Needs a B200: `pytest -m gpu`."""

import copy
import pickle

import numpy as np
import pytest

from tests.test_gpu_parity import assert_same_bits, random_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import interpn_b200

    return interpn_b200


CLASSES = [("linear", "MultilinearRegular", "MultilinearRectilinear"), ("cubic", "MulticubicRegular", "MulticubicRectilinear"),
           ("nearest", "NearestRegular", "NearestRectilinear")]  # fmt: skip


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("ndims", [1, 3, 5])
@pytest.mark.parametrize("method,reg_cls,rect_cls", CLASSES)
def test_classes_eval_matches_the_oracle(ib, oracle, method, reg_cls, rect_cls, ndims, dtype):
    rng = np.random.default_rng(50 + ndims)
    n = 5000
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, 4, {1: 30, 3: 9, 5: 5}[ndims], dtype)
    lins = (False, True) if method == "cubic" else (True,)
    for lin in lins:
        extra = (lin,) if method == "cubic" else ()
        want_reg = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=lin, nthreads=4)
        want_rect = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=lin, nthreads=4)
        reg = getattr(ib, reg_cls).new(dims, starts, steps, vals, *extra)
        rect = getattr(ib, rect_cls).new(grids, vals, *extra)
        for model, want in ((reg, want_reg), (rect, want_rect)):
            assert_same_bits(model.eval(obs), want, f"{type(model).__name__}.eval")
            out = np.full(n, -3.0, dtype=dtype)
            assert model.eval(obs, out) is out  # preallocated output, second call on the resident grid
            assert_same_bits(out, want)
            out2 = np.zeros(n, dtype=dtype)
            model.eval_unchecked(obs, out2)
            assert_same_bits(out2, want)
            # plain data after evaluation: JSON round trip, pickle and deepcopy each evaluate to the same bits
            for twin in (type(model).model_validate_json(model.model_dump_json()), pickle.loads(pickle.dumps(model)), copy.deepcopy(model)):
                assert_same_bits(twin.eval(obs), want)
            assert model.ndims() == ndims
        # check_bounds through the classes (multilinear/regular.rs:145-182)
        atol = dtype(1e-6)
        got = reg.check_bounds(obs, atol)
        assert list(got) == list(oracle.check_bounds_regular(dims, starts, steps, obs, float(atol)))
        got = rect.check_bounds(obs, atol)
        assert list(got) == list(oracle.check_bounds_rectilinear(grids, obs, float(atol)))
        inside = [np.full(7, g[1], dtype=dtype) for g in grids]
        assert not any(rect.check_bounds(inside, atol)) and not any(reg.check_bounds([np.full(7, s, dtype=dtype) for s in starts], atol))


def test_classes_reject_what_the_reference_rejects(ib):
    reg = ib.MultilinearRegular.new([4, 4], np.zeros(2), np.ones(2), np.arange(16.0))
    with pytest.raises(TypeError):  # dtype mismatch: no silent reallocation (multilinear_regular.py:101-123)
        reg.eval([np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.float32)])
    with pytest.raises(AssertionError, match="Dimension mismatch"):
        reg.eval([np.zeros(3)])
    with pytest.raises((AssertionError, ValueError, TypeError)):
        reg.eval([np.zeros(6)[::2], np.zeros(3)])  # non-contiguous
    out = np.zeros(3)
    with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
        reg.eval([np.array([0.5, np.nan, 0.5]), np.zeros(3)], out)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("method", ["linear", "cubic", "nearest"])
def test_interpn_function_dispatch(ib, oracle, method, dtype):
    """interpn(): regular-grid detection, assume_regular, out=, N-d shaped obs/vals, linearize flag, check_bounds."""
    rng = np.random.default_rng(8)
    x = np.linspace(0.0, 8.0, 9).astype(dtype)  # exactly regular in both dtypes
    y = np.linspace(-2.0, 2.0, 5).astype(dtype)
    yr = np.array([-2.0, -1.5, 0.0, 0.25, 2.0], dtype=dtype)
    vals = rng.standard_normal((9, 5)).astype(dtype)
    ox = (rng.random((40, 25)) * 10 - 1).astype(dtype)
    oy = (rng.random((40, 25)) * 5 - 2.5).astype(dtype)
    flat = [np.ascontiguousarray(ox.ravel()), np.ascontiguousarray(oy.ravel())]
    for lin in (True, False):
        got = ib.interpn(obs=[ox, oy], grids=[x, y], vals=vals, method=method, linearize_extrapolation=lin)
        assert got.shape == ox.shape
        starts = np.array([x[0], y[0]], dtype=dtype)
        steps = np.array([x[1] - x[0], y[1] - y[0]], dtype=dtype)
        want = oracle.interpn_regular(method, [9, 5], starts, steps, vals.ravel(), flat, linearize_extrapolation=lin)
        assert_same_bits(np.ascontiguousarray(got.ravel()), want, "regular dispatch")
        got = ib.interpn(obs=[ox, oy], grids=[x, yr], vals=vals, method=method, linearize_extrapolation=lin)
        want = oracle.interpn_rectilinear(method, [x, yr], vals.ravel(), flat, linearize_extrapolation=lin)
        assert_same_bits(np.ascontiguousarray(got.ravel()), want, "rectilinear dispatch")
        # assume_regular=True takes starts/steps from the first two nodes even of an irregular axis
        got = ib.interpn(obs=flat, grids=[x, yr], vals=vals, method=method, linearize_extrapolation=lin, assume_regular=True)
        steps_r = np.array([x[1] - x[0], yr[1] - yr[0]], dtype=dtype)
        want = oracle.interpn_regular(method, [9, 5], np.array([x[0], yr[0]], dtype=dtype), steps_r, vals.ravel(), flat, linearize_extrapolation=lin)
        assert_same_bits(got, want, "assume_regular")
    out = np.full(ox.size, -1.0, dtype=dtype)
    ret = ib.interpn(obs=flat, grids=[x, y], vals=vals, method=method, out=out)
    assert np.shares_memory(ret, out) and not np.any(out == -1.0)
    inside = [np.array([1.0, 7.5], dtype=dtype), np.array([-1.0, 1.5], dtype=dtype)]
    assert ib.interpn(obs=inside, grids=[x, yr], vals=vals, method=method, check_bounds=True).shape == (2,)
    with pytest.raises(ValueError, match="violate interpolator bounds"):
        ib.interpn(obs=[np.array([1.0, 9.5], dtype=dtype), inside[1]], grids=[x, yr], vals=vals, method=method, check_bounds=True)
    with pytest.raises(ValueError, match="violate interpolator bounds"):
        ib.interpn(obs=[inside[0], np.array([-2.5, 0.0], dtype=dtype)], grids=[x, y], vals=vals, method=method, check_bounds=True)
    with pytest.raises(ValueError):
        ib.interpn(obs=inside, grids=[x, y], vals=vals, method="quintic")
    with pytest.raises(AssertionError):
        ib.interpn(obs=inside, grids=[x, y], vals=vals.astype(np.float16), method=method)
