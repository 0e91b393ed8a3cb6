"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/interpn_b200.h
declares, its host-side validation reproduces the reference's error messages and precedence, and a
compute call without a GPU fails loudly instead of falling back. No kernel runs here."""

import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "interpn_b200.h")


def _has_gpu() -> bool:
    import interpn_b200

    return interpn_b200.device_count() > 0


def declared_symbols() -> set[str]:
    """Function names declared by the header after macro expansion (gcc -E)."""
    src = subprocess.run(["gcc", "-E", "-P", HEADER], check=True, capture_output=True, text=True).stdout
    return set(re.findall(r"\b(interpn_b200_\w+)\s*\(", src))


def test_library_exports_every_declared_symbol():
    import interpn_b200._lib as L

    names = declared_symbols()
    assert len(names) == 48, sorted(names)
    # both arithmetic flavours of the library export the same ABI and say which one they are
    here = os.path.dirname(L.LIB_PATH)
    for fname, flavour in (("libinterpn_b200.so", 0), ("libinterpn_b200_fma.so", 1)):
        lib = ctypes.CDLL(os.path.join(here, fname))
        missing = [n for n in sorted(names) if not hasattr(lib, n)]
        assert not missing, (fname, missing)
        assert lib.interpn_b200_arithmetic() == flavour
    for sfx in ("f64", "f32"):
        for m in ("linear", "cubic", "nearest"):
            for g in ("regular", "rectilinear"):
                assert f"interpn_b200_{m}_{g}_{sfx}" in names
        assert f"interpn_b200_check_bounds_regular_{sfx}" in names
        assert f"interpn_b200_one_dim_rectilinear_{sfx}" in names


def test_strerror_carries_the_reference_literals():
    import interpn_b200._lib as L

    expect = {
        1: "Dimension mismatch",
        2: "All grids must have at least two entries",
        3: "All grids must have at least 2 entries",
        4: "All grids must have at least four entries",
        5: "All grids must have at least 4 entries",
        6: "All grids must be monotonically increasing",
        7: "Unrepresentable coordinate value",
        8: "Dimension exceeds maximum (8). Use interpolator struct directly for higher dimensions.",
        9: "Dimension exceeds maximum (6).",
        10: "Length mismatch",
        11: "Unrepresentable number",
    }
    for code, msg in expect.items():
        assert L.lib.interpn_b200_strerror(code).decode() == msg
    assert L.lib.interpn_b200_strerror(0).decode() == ""


def test_raw_module_has_the_sixteen_reference_bindings():
    import interpn_b200

    ref = [
        f"interpn_{m}_{g}_{t}" for m in ("linear", "nearest", "cubic") for g in ("regular", "rectilinear") for t in ("f64", "f32")
    ] + [f"check_bounds_{g}_{t}" for g in ("regular", "rectilinear") for t in ("f64", "f32")]
    assert sorted(interpn_b200.raw.__all__) == sorted(ref)
    for name in ref:
        assert callable(getattr(interpn_b200.raw, name))


F = np.float64


def _reg(method, dims, starts, steps, vals, obs, out, lin=True):
    import interpn_b200 as ib

    fn = getattr(ib.raw, f"interpn_{method}_regular_f64")
    args = (dims, np.asarray(starts, F), np.asarray(steps, F), np.asarray(vals, F))
    args += ((lin,) if method == "cubic" else ()) + ([np.asarray(o, F) for o in obs], out)
    return fn(*args)


def _rect(method, grids, vals, obs, out, lin=True):
    import interpn_b200 as ib

    fn = getattr(ib.raw, f"interpn_{method}_rectilinear_f64")
    args = ([np.asarray(g, F) for g in grids], np.asarray(vals, F))
    args += ((lin,) if method == "cubic" else ()) + ([np.asarray(o, F) for o in obs], out)
    return fn(*args)


@pytest.mark.parametrize("method", ["linear", "cubic", "nearest"])
def test_regular_validation_messages(method, oracle):
    """Same inputs -> same error as the oracle's restatement of the reference dispatchers
    (multilinear/regular.rs:51-117, multicubic/regular.rs:52-136, nearest/regular.rs:41-101)."""
    mind = 4 if method == "cubic" else 2
    good_dims = [mind, mind + 1]
    nv = good_dims[0] * good_dims[1]
    out = np.zeros(3)
    obs = [np.zeros(3), np.zeros(3)]
    cases = {
        "short starts": (good_dims, [0.0], [1.0, 1.0], np.zeros(nv), obs, out),
        "missing obs axis": (good_dims, [0.0, 0.0], [1.0, 1.0], np.zeros(nv), obs[:1], out),
        "wrong nvals": (good_dims, [0.0, 0.0], [1.0, 1.0], np.zeros(nv + 1), obs, out),
        "degenerate axis": ([mind - 1, mind + 1], [0.0, 0.0], [1.0, 1.0], np.zeros((mind - 1) * (mind + 1)), obs, out),
        "zero step": (good_dims, [0.0, 0.0], [1.0, 0.0], np.zeros(nv), obs, out),
        "nan step": (good_dims, [0.0, 0.0], [np.nan, 1.0], np.zeros(nv), obs, out),
        "obs length": (good_dims, [0.0, 0.0], [1.0, 1.0], np.zeros(nv), [np.zeros(3), np.zeros(4)], out),
        "too many dims": ([mind] * 9, [0.0] * 9, [1.0] * 9, np.zeros(mind**9 if mind == 2 else 1), [np.zeros(3)] * 9, out),
        "zero dims": ([], [], [], np.zeros(1), [], out),
    }
    for label, (dims, starts, steps, vals, o, ou) in cases.items():
        if len(o) > 8:
            continue  # the binding layer rejects >8 arrays before the C ABI (python.rs:46-50 panics)
        with pytest.raises(oracle.OracleError) as want:
            oracle.interpn_regular(method, dims, starts, steps, vals, o, out=ou.copy(), dtype=F)
        with pytest.raises(AssertionError) as got:
            _reg(method, dims, starts, steps, vals, o, ou)
        assert str(got.value) == str(want.value), label
    # nine dims, eight obs arrays: multilinear/nearest check lengths first (regular.rs:60-62), multicubic
    # matches on ndims first (multicubic/regular.rs:64-65)
    expect = r"Dimension exceeds maximum \(8\)" if method == "cubic" else "Dimension mismatch"
    with pytest.raises(AssertionError, match=expect):
        _reg(method, [mind] * 9, [0.0] * 9, [1.0] * 9, np.zeros(1), [np.zeros(3)] * 8, out)


@pytest.mark.parametrize("method", ["linear", "cubic", "nearest"])
def test_rectilinear_validation_messages(method, oracle):
    mind = 4 if method == "cubic" else 2
    gx = np.arange(mind, dtype=F)
    gy = np.arange(mind + 1, dtype=F) * 2.0
    nv = gx.size * gy.size
    out = np.zeros(3)
    obs = [np.zeros(3), np.zeros(3)]
    cases = {
        "missing obs axis": ([gx, gy], np.zeros(nv), obs[:1], out),
        "wrong nvals": ([gx, gy], np.zeros(nv - 1), obs, out),
        "degenerate axis": ([gx[: mind - 1], gy], np.zeros((mind - 1) * gy.size), obs, out),
        "not monotonic": ([gx[::-1].copy(), gy], np.zeros(nv), obs, out),
        "equal first two": ([np.array([0.0, 0.0, 1.0, 2.0, 3.0])[: max(mind, 2) + 1], gy], np.zeros((max(mind, 2) + 1) * gy.size), obs, out),
        "obs length": ([gx, gy], np.zeros(nv), [np.zeros(3), np.zeros(2)], out),
        "zero dims": ([], np.zeros(1), [], out),
    }
    for label, (grids, vals, o, ou) in cases.items():
        with pytest.raises(oracle.OracleError) as want:
            oracle.interpn_rectilinear(method, grids, vals, o, out=ou.copy(), dtype=F)
        with pytest.raises(AssertionError) as got:
            _rect(method, grids, vals, o, ou)
        assert str(got.value) == str(want.value), label


def test_binding_layer_rejects_what_pyo3_rejects():
    import interpn_b200 as ib

    dims, starts, steps, vals = [2, 2], np.zeros(2), np.ones(2), np.zeros(4)
    obs, out = [np.zeros(3), np.zeros(3)], np.zeros(3)
    with pytest.raises(TypeError):  # wrong dtype
        ib.raw.interpn_linear_regular_f64(dims, starts.astype(np.float32), steps, vals, obs, out)
    with pytest.raises(TypeError):  # not contiguous
        ib.raw.interpn_linear_regular_f64(dims, starts, steps, vals, [np.zeros(6)[::2], obs[1]], out)
    with pytest.raises(TypeError):  # not an array
        ib.raw.interpn_linear_regular_f64(dims, starts, steps, vals, [[0.0, 0.0, 0.0], obs[1]], out)
    with pytest.raises(TypeError):  # 2-D
        ib.raw.interpn_linear_regular_f64(dims, starts, steps, vals.reshape(2, 2), obs, out)
    with pytest.raises(OverflowError):
        ib.raw.interpn_linear_regular_f64([-2, 2], starts, steps, vals, obs, out)


def test_one_dim_and_check_bounds_validation():
    import interpn_b200 as ib

    with pytest.raises(AssertionError, match="Length mismatch"):
        ib.one_dim.eval_regular("linear", 0.0, 1.0, np.zeros(5), np.zeros(3), np.zeros(4))
    with pytest.raises(AssertionError, match="Length mismatch"):
        ib.one_dim.eval_rectilinear("left", np.arange(4.0), np.zeros(5), np.zeros(3))
    with pytest.raises(AssertionError, match="Length mismatch"):
        ib.one_dim.eval_rectilinear("left", np.arange(1.0), np.zeros(1), np.zeros(3))
    with pytest.raises(AssertionError, match="Dimension mismatch"):
        ib.raw.check_bounds_regular_f64([3, 3], np.zeros(2), np.ones(2), [np.zeros(3)], 1e-8, np.zeros(2, dtype=bool))
    with pytest.raises(AssertionError, match="Dimension mismatch"):
        ib.raw.check_bounds_rectilinear_f64([np.arange(3.0)], [np.zeros(3)], 1e-8, np.zeros(2, dtype=bool))


def test_compute_without_a_gpu_fails_loudly():
    """No CPU fallback: a valid call on a box without an sm_100 device raises InterpnDeviceError."""
    import interpn_b200 as ib

    if _has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(ib.InterpnDeviceError, match="no CPU fallback"):
        ib.raw.interpn_linear_regular_f64([2, 2], np.zeros(2), np.ones(2), np.zeros(4), [np.zeros(3), np.zeros(3)], np.zeros(3))
    with pytest.raises(ib.InterpnDeviceError):
        ib.Interpolator.regular("cubic", [4], np.zeros(1), np.ones(1), np.zeros(4))
    with pytest.raises(ib.InterpnDeviceError):
        ib.one_dim.eval_regular("nearest", 0.0, 1.0, np.zeros(4), np.zeros(2))
    assert ib.launch_count() == 0


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under interpn_b200/ may reference it."""
    pkg = os.path.join(ROOT, "interpn_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle|liboracle|interpn_oracle|#include.*oracle", text, re.M):
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_python_classes_mirror_the_reference_surface():
    import interpn_b200 as ib

    m = ib.MulticubicRegular.new([4, 5], np.zeros(2), np.ones(2), np.arange(20.0))
    assert m.linearize_extrapolation is True and m.ndims() == 2
    m2 = ib.MulticubicRegular.model_validate_json(m.model_dump_json())
    assert np.array_equal(m2.vals.data, m.vals.data) and m2.dims == [4, 5]
    r = ib.NearestRectilinear.new([np.arange(3.0), np.arange(4.0)], np.arange(12.0))
    assert r.dims() == [3, 4]
    with pytest.raises(Exception, match="monotonically increasing"):
        ib.MultilinearRectilinear.new([np.array([0.0, 2.0, 1.0])], np.arange(3.0))
    with pytest.raises(Exception, match="no more than 6"):
        ib.NearestRegular.new([2] * 7, np.zeros(7), np.ones(7), np.zeros(128))
    with pytest.raises(Exception, match="positive and nonzero"):
        ib.MultilinearRegular.new([2], np.zeros(1), np.zeros(1), np.zeros(2))
    f32 = ib.MultilinearRegular.new([2], np.zeros(1, np.float32), np.ones(1, np.float32), np.zeros(2, np.float32))
    assert f32.vals.data.dtype == np.float32 and json_dtype(f32) == "float32"


def json_dtype(model) -> str:
    import json

    return json.loads(model.model_dump_json())["vals"]["dtype"]


def test_models_stay_plain_data_after_evaluation(monkeypatch):
    """the grid-resident device handle (a ctypes pointer) must not live in a model's state — the reference's
    models are plain data and can be pickled, deep-copied and compared at any time."""
    import copy
    import ctypes
    import gc
    import pickle

    import interpn_b200 as ib
    from interpn_b200 import api

    closed = []

    class FakeResident:
        def __init__(self):
            self.handle = ctypes.c_void_p(0xDEAD)  # what makes a real Interpolator unpicklable

        def eval(self, obs, out):
            out[:] = 1.0
            return out

        def close(self):
            closed.append(self)

    for cls in (ib.MultilinearRegular, ib.MulticubicRegular, ib.NearestRegular):
        monkeypatch.setattr(cls, "_build", lambda self: FakeResident())
    m = ib.MulticubicRegular.new([4, 5], np.zeros(2), np.ones(2), np.arange(20.0))
    fresh = ib.MulticubicRegular.new([4, 5], np.zeros(2), np.ones(2), np.arange(20.0))
    assert m.eval([np.zeros(3), np.zeros(3)]).tolist() == [1.0, 1.0, 1.0]
    assert id(m) in api._RESIDENT and id(fresh) not in api._RESIDENT
    c = copy.deepcopy(m)
    c2 = m.model_copy(deep=True)
    p = pickle.loads(pickle.dumps(m))
    for other in (c, c2, p):
        assert id(other) not in api._RESIDENT  # a copy builds its own resident interpolator lazily
        assert other.model_dump_json() == m.model_dump_json() == fresh.model_dump_json()
        assert other.eval([np.zeros(2), np.zeros(2)]).tolist() == [1.0, 1.0]
    assert m.__pydantic_private__ == fresh.__pydantic_private__  # nothing hidden distinguishes an evaluated model
    key = id(m)
    del m
    gc.collect()
    assert key not in api._RESIDENT and len(closed) == 1  # the handle is released with its model
