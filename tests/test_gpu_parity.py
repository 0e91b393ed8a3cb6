"""Parity of the CUDA path against the CPU oracle (strict arithmetic = crate default features).

Bar (north_star): cell indices and nearest-neighbour results bit-exact; f64 linear/cubic within
4 ULP or 1e-13 relative; f32 within 4 ULP or 5e-5 relative. The kernels reproduce the reference's
operation order without contraction, so these tests assert the stronger property — identical
bits — for every method, and keep the ULP bound only as the documented fallback tolerance.

Needs a B200: `pytest -m gpu`. Every call goes through the C ABI (ctypes -> libinterpn_b200.so).
"""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ULP_TOL = 4  # documented tolerance; the assertions below require 0


@pytest.fixture(scope="module")
def ib():
    import interpn_b200

    return interpn_b200


def bits(a: np.ndarray) -> np.ndarray:
    return a.view(np.uint64 if a.dtype == np.float64 else np.uint32)


def assert_same_bits(got: np.ndarray, want: np.ndarray, what: str = ""):
    assert got.dtype == want.dtype and got.shape == want.shape
    same = bits(got) == bits(want)
    both_nan = np.isnan(got) & np.isnan(want)
    ok = same | both_nan
    if not ok.all():
        bad = np.flatnonzero(~ok)
        i = bad[0]
        raise AssertionError(
            f"{what}: {bad.size}/{got.size} results differ from the oracle; first at {i}: "
            f"gpu={got[i]!r} oracle={want[i]!r}"
        )


def random_case(rng, ndims, n, min_dim, max_dim, dtype):
    dims = [int(rng.integers(min_dim, max_dim + 1)) for _ in range(ndims)]
    grids = []
    for d in dims:
        g = np.cumsum(rng.random(d) + 0.05) + rng.normal() * 3.0
        grids.append(np.ascontiguousarray(g.astype(dtype)))
        assert np.all(np.diff(grids[-1]) > 0)
    starts = np.array([g[0] for g in grids], dtype=dtype)
    steps = np.array([(g[-1] - g[0]) / (len(g) - 1) for g in grids], dtype=dtype)
    vals = rng.standard_normal(int(np.prod(dims))).astype(dtype)
    obs = []
    for d, g in enumerate(grids):
        lo, hi = float(g[0]), float(g[-1])
        span = hi - lo
        x = lo - 0.3 * span + 1.6 * span * rng.random(n)  # ~37 % outside the grid on each axis
        # plant exact nodes (regular and rectilinear), exact mid-points and exact end points
        k = rng.integers(0, dims[d], size=n)
        reg_nodes = (starts[d] + steps[d] * k.astype(dtype)).astype(np.float64)
        rect_nodes = g[k].astype(np.float64)
        sel = rng.random(n)
        x = np.where(sel < 0.05, reg_nodes, x)
        x = np.where((sel >= 0.05) & (sel < 0.10), rect_nodes, x)
        k2 = np.minimum(k, dims[d] - 2)
        mid = (g[k2].astype(np.float64) + g[k2 + 1].astype(np.float64)) * 0.5
        x = np.where((sel >= 0.10) & (sel < 0.13), mid, x)
        obs.append(np.ascontiguousarray(x.astype(dtype)))
    return dims, grids, starts, steps, vals, obs


CASES = []
for _dtype in (np.float64, np.float32):
    for _n in range(1, 9):
        CASES.append(("linear", _n, _dtype))
    for _n in range(1, 7):
        CASES.append(("cubic", _n, _dtype))
    for _n in range(1, 7):
        CASES.append(("nearest", _n, _dtype))


@pytest.mark.parametrize("layout", ["plain", "window"])
@pytest.mark.parametrize("method,ndims,dtype", CASES, ids=lambda v: getattr(v, "__name__", str(v)))
def test_random_grids_bit_exact(ib, oracle, monkeypatch, method, ndims, dtype, layout):
    """`window` forces the derived gather layouts (window copy for multilinear, transposed window +
    quad-cooperative kernels for multicubic N = 2..4) onto these small grids; `plain` disables them."""
    if layout == "window":
        if method == "nearest" or ndims > (6 if method == "linear" else 4):
            pytest.skip("no window layout for this method / dimensionality")
        monkeypatch.setenv("INTERPN_B200_WINDOW_MIN_KB", "0")
    else:
        monkeypatch.setenv("INTERPN_B200_WINDOW_MB", "0")
    rng = np.random.default_rng(1000 * ndims + len(method))
    min_dim = 4 if method == "cubic" else 2
    max_dim = {1: 40, 2: 20, 3: 12, 4: 8, 5: 6, 6: 5, 7: 4, 8: 4}[ndims]
    n = 20000 if (method != "cubic" or ndims <= 4) else (3000 if ndims == 5 else 800)
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, min_dim, max(max_dim, min_dim), dtype)
    sfx = "f64" if dtype == np.float64 else "f32"
    for linearize in ((False, True) if method == "cubic" else (True,)):
        # regular
        out = np.zeros(n, dtype=dtype)
        fn = getattr(ib.raw, f"interpn_{method}_regular_{sfx}")
        args = (dims, starts, steps, vals) + ((linearize,) if method == "cubic" else ()) + (obs, out)
        fn(*args)
        want = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=linearize, nthreads=4)
        assert_same_bits(out, want, f"{method} regular N={ndims} {sfx} lin={linearize}")
        # rectilinear
        out = np.zeros(n, dtype=dtype)
        fn = getattr(ib.raw, f"interpn_{method}_rectilinear_{sfx}")
        args = (grids, vals) + ((linearize,) if method == "cubic" else ()) + (obs, out)
        fn(*args)
        want = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=linearize, nthreads=4)
        assert_same_bits(out, want, f"{method} rectilinear N={ndims} {sfx} lin={linearize}")


SWEEP_CASES = [("linear", n, dt) for dt in (np.float64, np.float32) for n in (2, 3, 4, 6, 8)] + [
    ("cubic", n, dt) for dt in (np.float64, np.float32) for n in (2, 3, 4)
]


@pytest.mark.parametrize("method,ndims,dtype", SWEEP_CASES, ids=lambda v: getattr(v, "__name__", str(v)))
def test_bin_swept_evaluation_bit_exact(ib, oracle, monkeypatch, method, ndims, dtype):
    """The bin-swept kernels (sweep.cuh: grids beyond L2) forced onto small grids: tiny slabs so the key
    spans several dimensions, tiny tiles so every CTA sorts several of them, a ragged last tile, and
    unrepresentable points in the middle of the batch."""
    monkeypatch.setenv("INTERPN_B200_SWEEP_MIN_MB", "0")
    monkeypatch.setenv("INTERPN_B200_SWEEP_MIN_POINTS", "0")
    monkeypatch.setenv("INTERPN_B200_SWEEP_MIN_ROWS", "0")
    monkeypatch.setenv("INTERPN_B200_SWEEP_SLAB_KB", "0")
    monkeypatch.setenv("INTERPN_B200_SWEEP_CHUNK", "150000")
    if ndims % 2 == 0:
        monkeypatch.setenv("INTERPN_B200_WINDOW_MIN_KB", "0")  # sweep over the window layouts as well
    rng = np.random.default_rng(77 * ndims + len(method))
    min_dim = {2: 8, 3: 5}.get(ndims, 4 if method == "cubic" else 2)
    max_dim = {2: 40, 3: 14, 4: 9, 6: 5, 8: 4}[ndims]
    n = 700_001 if ndims <= 4 else 50_001
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, min_dim, max(max_dim, min_dim), dtype)
    sfx = "f64" if dtype == np.float64 else "f32"
    before = ib.swept_launch_count()
    lin = bool(ndims % 2)
    extra = (lin,) if method == "cubic" else ()
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_{method}_regular_{sfx}")(dims, starts, steps, vals, *extra, obs, out)
    assert_same_bits(out, oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=lin, nthreads=8))
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_{method}_rectilinear_{sfx}")(grids, vals, *extra, obs, out)
    assert_same_bits(out, oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=lin, nthreads=8))
    assert ib.swept_launch_count() >= before + 2, "the bin-swept kernels did not run"
    # reference failure semantics survive the reordering: the smallest failing index is reported and
    # everything before it is written
    bad = [n // 3, n // 3 + 5000]
    obs2 = [o.copy() for o in obs]
    for b in bad:
        obs2[0][b] = np.nan
    out = np.full(n, -7.0, dtype=dtype)
    with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
        getattr(ib.raw, f"interpn_{method}_regular_{sfx}")(dims, starts, steps, vals, *extra, obs2, out)
    want = oracle.interpn_regular(method, dims, starts, steps, vals, [o[: bad[0]] for o in obs], linearize_extrapolation=lin, nthreads=8)
    assert_same_bits(out[: bad[0]], want)
    assert np.all(out[bad[0] :] == -7.0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ndims", [3, 4, 5])
def test_slab_passes_bit_exact(ib, oracle, monkeypatch, ndims, dtype):
    """The slab passes of the multilinear kernels (kernels.cuh linear_slab_kernel: grids a little beyond L2, C3-linear)
    forced onto small grids: several passes, a ragged last tile, points exactly on the pass edges, below and above the
    grid, and unrepresentable points reported from different passes."""
    monkeypatch.setenv("INTERPN_B200_WINDOW_MB", "0")
    monkeypatch.setenv("INTERPN_B200_SLAB_MIN_KB", "0")
    monkeypatch.setenv("INTERPN_B200_SLAB_MIN_POINTS", "0")
    rng = np.random.default_rng(991 + ndims)
    n = 300_007
    lo, hi = {3: (9, 16), 4: (6, 9), 5: (5, 6)}[ndims]
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, lo, hi, dtype)
    obs[0][:64] = np.resize(grids[0], 64)  # exact nodes of axis 0: the pass edges among them
    obs[0][64:128] = np.resize(starts[0] + steps[0] * np.arange(dims[0], dtype=dtype), 64)
    sfx = "f64" if dtype == np.float64 else "f32"
    pass_kb = max(1, -(-vals.nbytes // (3 * 1024)))
    passes = -(-vals.nbytes // (pass_kb * 1024))
    assert 2 <= passes <= dims[0] - 1
    monkeypatch.setenv("INTERPN_B200_SLAB_PASS_KB", str(pass_kb))
    before = ib.launch_count()
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_linear_regular_{sfx}")(dims, starts, steps, vals, obs, out)
    assert_same_bits(out, oracle.interpn_regular("linear", dims, starts, steps, vals, obs, nthreads=8))
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_linear_rectilinear_{sfx}")(grids, vals, obs, out)
    assert_same_bits(out, oracle.interpn_rectilinear("linear", grids, vals, obs, nthreads=8))
    assert ib.launch_count() >= before + 2 * passes, "the slab-pass kernel did not run"
    bad = [n // 3, n // 3 + 5000, n // 2]
    obs2 = [o.copy() for o in obs]
    obs2[0][bad[0]] = np.nan
    obs2[1][bad[1]] = np.inf
    obs2[0][bad[2]] = -np.inf
    out = np.full(n, -7.0, dtype=dtype)
    with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
        getattr(ib.raw, f"interpn_linear_regular_{sfx}")(dims, starts, steps, vals, obs2, out)
    want = oracle.interpn_regular("linear", dims, starts, steps, vals, [o[: bad[0]] for o in obs], nthreads=8)
    assert_same_bits(out[: bad[0]], want)
    assert np.all(out[bad[0] :] == -7.0)
    # rectilinear grids never fail (NaN and infinities land in the end cells, like slice::partition_point)
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_linear_rectilinear_{sfx}")(grids, vals, obs2, out)
    assert_same_bits(out, oracle.interpn_rectilinear("linear", grids, vals, obs2, nthreads=8))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ndims", [3, 4, 5, 6])
def test_hypercube_layout_bit_exact(ib, oracle, monkeypatch, ndims, dtype):
    """The hypercube layout of the multilinear kernels (kernels.cuh linear_hyper_kernel / linear_hyper3_kernel: N = 3..6 grids beyond L2, C3-linear and C4)
    forced onto small grids: every cell incl. the last one of each axis, points on nodes, outside the grid, and
    unrepresentable points (regular) / NaN and infinities (rectilinear)."""
    monkeypatch.setenv("INTERPN_B200_HYPER_MIN_KB", "0")
    rng = np.random.default_rng(4471 + ndims)
    n = 200_003
    lo, hi = {3: (5, 12), 4: (4, 8), 5: (3, 6), 6: (3, 5)}[ndims]
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, lo, hi, dtype)
    for d in range(ndims):  # exact nodes of every axis, the last one among them
        obs[d][d * 64 : d * 64 + 64] = np.resize(grids[d], 64)
    sfx = "f64" if dtype == np.float64 else "f32"
    before = ib.launch_count()
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_linear_regular_{sfx}")(dims, starts, steps, vals, obs, out)
    assert_same_bits(out, oracle.interpn_regular("linear", dims, starts, steps, vals, obs, nthreads=8))
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_linear_rectilinear_{sfx}")(grids, vals, obs, out)
    assert_same_bits(out, oracle.interpn_rectilinear("linear", grids, vals, obs, nthreads=8))
    assert ib.launch_count() >= before + 4, "expected a layout build and an evaluation launch per call"
    bad = [n // 3, n // 3 + 5000, n // 2]
    obs2 = [o.copy() for o in obs]
    obs2[0][bad[0]] = np.nan
    obs2[1][bad[1]] = np.inf
    obs2[0][bad[2]] = -np.inf
    out = np.full(n, -7.0, dtype=dtype)
    with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
        getattr(ib.raw, f"interpn_linear_regular_{sfx}")(dims, starts, steps, vals, obs2, out)
    want = oracle.interpn_regular("linear", dims, starts, steps, vals, [o[: bad[0]] for o in obs], nthreads=8)
    assert_same_bits(out[: bad[0]], want)
    assert np.all(out[bad[0] :] == -7.0)
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_linear_rectilinear_{sfx}")(grids, vals, obs2, out)
    assert_same_bits(out, oracle.interpn_rectilinear("linear", grids, vals, obs2, nthreads=8))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_cubic_four_node_axes(ib, oracle, dtype):
    """Axes with exactly 4 nodes: origin is always 0 and every saturation class is reachable
    (SURVEY.md appendix A)."""
    rng = np.random.default_rng(5)
    dims, grids, starts, steps, vals, obs = random_case(rng, 3, 30000, 4, 4, dtype)
    sfx = "f64" if dtype == np.float64 else "f32"
    for lin in (False, True):
        out = np.zeros(30000, dtype=dtype)
        getattr(ib.raw, f"interpn_cubic_regular_{sfx}")(dims, starts, steps, vals, lin, obs, out)
        assert_same_bits(out, oracle.interpn_regular("cubic", dims, starts, steps, vals, obs, linearize_extrapolation=lin))
        out = np.zeros(30000, dtype=dtype)
        getattr(ib.raw, f"interpn_cubic_rectilinear_{sfx}")(grids, vals, lin, obs, out)
        assert_same_bits(out, oracle.interpn_rectilinear("cubic", grids, vals, obs, linearize_extrapolation=lin))


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ndims", [1, 2, 3, 4])
@pytest.mark.parametrize("method", ["linear", "cubic"])
def test_plateau_grids_keep_the_sign_of_zero(ib, oracle, method, ndims, dtype):
    """Grid values drawn from {-1, -0.0, 0, 1}: neighbouring values are often equal and many results are exactly
    zero. The permuted end-cell formulas of the four-points-per-quad cubic kernels (cubic_quad4.cuh) evaluate v0-v2
    where the reference evaluates -(v2-v0); the two differ in the sign of a zero, which neg_zero_if restores — so the
    comparison stays bit for bit, including -0 against +0."""
    rng = np.random.default_rng(77 + ndims)
    n = 40000
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, 4, {1: 30, 2: 12, 3: 8, 4: 6}[ndims], dtype)
    vals = rng.choice(np.array([-1.0, -0.0, 0.0, 1.0]), size=vals.size, p=[0.2, 0.2, 0.4, 0.2]).astype(dtype)
    sfx = "f64" if dtype == np.float64 else "f32"
    for linearize in ((False, True) if method == "cubic" else (True,)):
        extra = (linearize,) if method == "cubic" else ()
        out = np.full(n, 7.0, dtype=dtype)
        getattr(ib.raw, f"interpn_{method}_regular_{sfx}")(dims, starts, steps, vals, *extra, obs, out)
        want = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=linearize, nthreads=4)
        if ndims <= 2:
            assert (want == 0).sum() > n // 100  # the case is exercised
        assert_same_bits(out, want, f"plateau {method} regular N={ndims} {sfx} lin={linearize}")
        out = np.full(n, 7.0, dtype=dtype)
        getattr(ib.raw, f"interpn_{method}_rectilinear_{sfx}")(grids, vals, *extra, obs, out)
        want = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=linearize, nthreads=4)
        assert_same_bits(out, want, f"plateau {method} rectilinear N={ndims} {sfx} lin={linearize}")


WORKLOADS = [
    ("x_linear3d_reg100", np.float64, 200_000),
    ("x_linear4d_reg32", np.float64, 200_000),
    ("x_linear4d_rect32", np.float64, 200_000),
    ("x_cubic4d_reg32", np.float64, 100_000),
    ("x_cubic3d_rect100", np.float64, 200_000),
    ("x_cubic4d_rect32", np.float64, 60_000),
    ("x_cubic4d_reg32", np.float32, 100_000),
    ("x_cubic3d_rect100", np.float32, 200_000),
    ("c1_linear3d_reg20", np.float64, 200_000),
    ("c2_cubic3d_reg100", np.float64, 200_000),
    ("c3_linear4d_rect64", np.float64, 200_000),
    ("c3_linear4d_rect64", np.float64, 1_100_003),  # above the slab passes' point threshold
    ("x_linear3d_reg256", np.float64, 1_100_003),  # regular grids a little beyond L2: slab passes
    ("x_linear4d_reg64", np.float64, 1_100_003),
    ("x_linear5d_reg26", np.float32, 1_100_003),  # 48 MB in f32: the direct kernel; f64 below
    ("x_linear5d_reg26", np.float64, 1_100_003),
    ("c3_cubic4d_rect64", np.float64, 60_000),
    ("c4_linear6d_reg24", np.float64, 100_000),
    ("c5_nearest2d_reg1024", np.float64, 300_000),
    ("c5_nearest2d_reg1024", np.float32, 300_000),
    ("c5_nearest3d_reg128", np.float64, 300_000),
    ("c5_nearest3d_reg128", np.float32, 300_000),
    ("c5_nearest2d_rect1024", np.float64, 300_000),
    ("c5_nearest2d_rect1024", np.float32, 300_000),
    ("c5_nearest3d_rect128", np.float64, 300_000),
    ("c5_nearest3d_rect128", np.float32, 300_000),
]


@pytest.mark.parametrize("name,dtype,n", WORKLOADS, ids=lambda v: getattr(v, "__name__", str(v)))
def test_baseline_configs_bit_exact(ib, oracle, name, dtype, n):
    """All five BASELINE.json configurations (grids at full size, query batch at a size the oracle
    finishes in seconds), through the resident-interpolator C ABI with host buffers."""
    from interpn_b200 import workloads as W

    w = W.get(name, dtype)
    vals = w.vals("np")
    obs = w.queries(12345, n, "np")
    if w.rect:
        interp = ib.Interpolator.rectilinear(w.method, w.grids, vals, w.linearize)
        want = oracle.interpn_rectilinear(w.method, w.grids, vals, obs, linearize_extrapolation=w.linearize, nthreads=8)
    else:
        interp = ib.Interpolator.regular(w.method, w.dims, w.starts, w.steps, vals, w.linearize)
        want = oracle.interpn_regular(w.method, w.dims, w.starts, w.steps, vals, obs, linearize_extrapolation=w.linearize, nthreads=8)
    with interp:
        got = interp.eval(obs)
    assert_same_bits(got, want, name)


def test_multi_chunk_host_pipeline(ib, oracle):
    """n larger than one 32 MiB chunk per array: the three-slot H2D/kernel/D2H pipeline must give the
    same bits as one serial pass."""
    rng = np.random.default_rng(21)
    n = 9_500_000
    dims = [64, 48]
    starts = np.array([-1.0, 2.0])
    steps = np.array([0.25, 0.5])
    vals = rng.standard_normal(dims[0] * dims[1])
    obs = [rng.random(n) * 18.0 - 2.0, rng.random(n) * 26.0 + 1.0]
    out = np.zeros(n)
    ib.raw.interpn_linear_regular_f64(dims, starts, steps, vals, obs, out)
    want = oracle.interpn_regular("linear", dims, starts, steps, vals, obs, nthreads=8)
    assert_same_bits(out, want)
    obs32 = [o.astype(np.float32) for o in obs]
    out32 = np.zeros(n, dtype=np.float32)
    ib.raw.interpn_nearest_regular_f32(dims, starts.astype(np.float32), steps.astype(np.float32), vals.astype(np.float32), obs32, out32)
    want32 = oracle.interpn_regular("nearest", dims, starts.astype(np.float32), steps.astype(np.float32), vals.astype(np.float32), obs32, nthreads=8)
    assert_same_bits(out32, want32)


@pytest.mark.parametrize("method", ["linear", "cubic", "nearest"])
@pytest.mark.parametrize("badval", [np.nan, np.inf, -np.inf, 1e300])
def test_unrepresentable_point_semantics(ib, oracle, method, badval):
    """Regular grids: Err("Unrepresentable coordinate value"), earlier outputs written, later ones
    untouched (ref: multilinear/regular.rs:276-280, 418) — identical to the serial reference loop."""
    rng = np.random.default_rng(3)
    n = 5000
    dims = [6, 5]
    starts = np.array([0.0, 0.0])
    steps = np.array([1.0, 1.0])
    vals = rng.standard_normal(30)
    obs = [rng.random(n) * 5.0, rng.random(n) * 4.0]
    obs[1][3210] = badval
    obs[0][4000] = np.nan  # a later failure must not win
    sentinel = -777.0
    out = np.full(n, sentinel)
    args = (dims, starts, steps, vals) + ((True,) if method == "cubic" else ()) + (obs, out)
    with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
        getattr(ib.raw, f"interpn_{method}_regular_f64")(*args)
    want = np.full(n, sentinel)
    with pytest.raises(oracle.OracleError) as ei:
        oracle.interpn_regular(method, dims, starts, steps, vals, obs, out=want, nthreads=1)
    assert ei.value.first_bad == 3210
    assert_same_bits(out, want)
    assert np.all(out[3210:] == sentinel) and not np.any(out[:3210] == sentinel)


@pytest.mark.parametrize("method", ["linear", "cubic", "nearest"])
def test_rectilinear_nan_and_inf_queries_do_not_fail(ib, oracle, method):
    """Rectilinear grids never fail per point: NaN -> partition_point 0 (ref: multilinear/rectilinear.rs:363)."""
    rng = np.random.default_rng(4)
    grids = [np.cumsum(rng.random(7) + 0.1), np.cumsum(rng.random(6) + 0.1)]
    vals = rng.standard_normal(42)
    obs = [rng.random(200) * 5.0, rng.random(200) * 4.0]
    obs[0][5], obs[1][9], obs[0][17], obs[1][17] = np.nan, np.inf, -np.inf, np.nan
    out = np.zeros(200)
    args = (grids, vals) + ((True,) if method == "cubic" else ()) + (obs, out)
    getattr(ib.raw, f"interpn_{method}_rectilinear_f64")(*args)
    want = oracle.interpn_rectilinear(method, grids, vals, obs)
    assert np.array_equal(np.isnan(out), np.isnan(want))
    ok = ~np.isnan(want)
    assert_same_bits(out[ok], want[ok])


def test_resident_interpolator_device_path_matches_host_path(ib, oracle):
    torch = pytest.importorskip("torch")
    from interpn_b200 import workloads as W

    w = W.get("c2_cubic3d_reg100")
    vals = w.vals("np")
    n = 300_000
    obs = w.queries(0, n, "np")
    with ib.Interpolator.regular("cubic", w.dims, w.starts, w.steps, vals, True) as interp:
        host = interp.eval(obs)
        dev_obs = [torch.from_numpy(o).cuda() for o in obs]
        dev_out = interp.eval_torch(dev_obs)
        interp.status(torch.cuda.current_stream().cuda_stream)
        assert_same_bits(dev_out.cpu().numpy(), host)
        # a device-side failure is latched and reported once, with the smallest index
        dev_obs[1][777] = float("nan")
        dev_obs[2][99] = float("inf")
        interp.eval_torch(dev_obs)
        with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
            interp.status(torch.cuda.current_stream().cuda_stream)
        assert interp.first_bad == 99
        interp.status(torch.cuda.current_stream().cuda_stream)  # cleared
        # zero-copy through __cuda_array_interface__ (what CuPy / Numba arrays expose; torch tensors do too)
        dev_obs = [torch.from_numpy(o).cuda() for o in obs]
        cai_out = torch.full((n,), -1.0, dtype=torch.float64, device="cuda")
        interp.eval_cuda_arrays(dev_obs, cai_out, torch.cuda.current_stream().cuda_stream)
        interp.status(torch.cuda.current_stream().cuda_stream)
        assert_same_bits(cai_out.cpu().numpy(), host)
        with pytest.raises(TypeError):
            interp.eval_cuda_arrays([o.float() for o in dev_obs], cai_out)
        with pytest.raises(AssertionError, match="Dimension mismatch"):
            interp.eval_cuda_arrays([dev_obs[0][:10], dev_obs[1], dev_obs[2]], cai_out)
    want = oracle.interpn_regular("cubic", w.dims, w.starts, w.steps, vals, obs, nthreads=8)
    assert_same_bits(host, want)


@pytest.mark.parametrize("method,ndims,rect", [("linear", 3, False), ("linear", 4, True), ("linear", 1, False), ("linear", 7, False),
                                               ("nearest", 3, False), ("nearest", 2, True), ("cubic", 3, False), ("cubic", 2, True)])
def test_fused_fields_equal_separate_calls(ib, oracle, method, ndims, rect):
    """Several fields over one grid and one query batch: one cell location per point for multilinear
    and nearest, field by field for the rest — either way every field equals its own evaluation (and the oracle) bit for
    bit, including ten fields (more than one fused launch holds) and the refusal of a different grid."""
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(31 + ndims)
    n = 30000
    max_dim = {1: 40, 2: 20, 3: 12, 4: 8, 7: 4}[ndims]
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, 4, max_dim, np.float64)
    nf = 10 if method == "nearest" else 3
    fields = [rng.standard_normal(vals.size) for _ in range(nf)]
    def make(v, g=grids, d=dims):
        if rect:
            return ib.Interpolator.rectilinear(method, g, v, True)
        return ib.Interpolator.regular(method, d, starts, steps, v, True)
    interps = [make(v) for v in fields]
    dev_obs = [torch.from_numpy(o).cuda() for o in obs]
    outs = ib.Interpolator.eval_fields_torch(interps, dev_obs)
    interps[0].status(torch.cuda.current_stream().cuda_stream)
    for k, v in enumerate(fields):
        if rect:
            want = oracle.interpn_rectilinear(method, grids, v, obs, linearize_extrapolation=True, nthreads=4)
        else:
            want = oracle.interpn_regular(method, dims, starts, steps, v, obs, linearize_extrapolation=True, nthreads=4)
        assert_same_bits(outs[k].cpu().numpy(), want, f"field {k}")
        assert_same_bits(interps[k].eval(obs), want, f"field {k} alone")
    # a point no regular grid can represent is latched on the first interpolator
    if not rect:
        dev_obs[0][123] = float("nan")
        ib.Interpolator.eval_fields_torch(interps, dev_obs, outs)
        with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
            interps[0].status(torch.cuda.current_stream().cuda_stream)
        assert interps[0].first_bad == 123
    # another grid is refused
    other_grids = [g * 1.5 for g in grids]
    if rect:
        other = ib.Interpolator.rectilinear(method, other_grids, fields[0], True)
    else:
        other = ib.Interpolator.regular(method, dims, starts, steps * 2.0, fields[0], True)
    with pytest.raises(AssertionError, match="Dimension mismatch"):
        ib.Interpolator.eval_fields_torch([interps[0], other], dev_obs)
    for it in interps + [other]:
        it.close()


@pytest.mark.parametrize("method,ndims,rect,hyper", [("linear", 3, False, False), ("cubic", 3, False, False), ("cubic", 3, True, False),
                                                     ("cubic", 4, True, False), ("linear", 4, True, True), ("linear", 3, False, True)])  # fmt: skip
def test_vals_from_device_and_uninitialised_storage(ib, monkeypatch, method, ndims, rect, hyper):
    """Values given on the device, and values written into uninitialised resident storage afterwards (the multi-GPU
    broadcast): `vals_updated` must rebuild whatever derived layout the grid has — patch, coefficient table (regular and
    rectilinear), hypercube blocks — so all three interpolators agree bit for bit."""
    torch = pytest.importorskip("torch")
    if hyper:
        monkeypatch.setenv("INTERPN_B200_HYPER_MIN_KB", "0")
    rng = np.random.default_rng(8 + ndims)
    dims = [9, 7, 5, 6][:ndims]
    starts, steps = np.zeros(ndims), np.ones(ndims)
    grids = [np.cumsum(rng.random(d) + 0.2) for d in dims]
    vals = rng.standard_normal(int(np.prod(dims)))
    obs = [rng.random(5000) * (g[-1] - g[0]) * 1.2 + g[0] - 0.1 * (g[-1] - g[0]) for g in (grids if rect else [np.arange(d, dtype=float) for d in dims])]

    def make(v, **kw):
        if rect:
            return ib.Interpolator.rectilinear(method, grids, v, True, **kw)
        return ib.Interpolator.regular(method, dims, starts, steps, v, True, **kw)

    a = make(vals)
    b = make(torch.from_numpy(vals).cuda())
    c = make(None, dtype=np.float64)
    assert c.vals_len == vals.size and c.vals_ptr != 0
    # fill c's storage the way the multi-GPU broadcast does: through a tensor aliasing the resident copy
    c.vals_tensor().copy_(torch.from_numpy(vals).cuda())
    c.vals_updated(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    ra, rb, rc = a.eval(obs), b.eval(obs), c.eval(obs)
    assert_same_bits(ra, rb)
    assert_same_bits(ra, rc)
    # new values through the same storage: the derived layout must follow
    vals2 = rng.standard_normal(vals.size)
    c.vals_tensor().copy_(torch.from_numpy(vals2).cuda())
    c.vals_updated(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    d = make(vals2)
    assert_same_bits(c.eval(obs), d.eval(obs))
    for x in (a, b, c, d):
        x.close()


def test_full_size_c2_properties(ib, oracle):
    """BASELINE.json config 2 at its full size (1e8 points, 100^3 grid, 10 % out of bounds), inputs
    generated on the device. Size-independent checks: (1) bit-exact against the oracle on a strided
    sample and on the first/last blocks; (2) with a linear field as `vals`, cubic interpolation with
    linearized extrapolation reproduces the field at every one of the 1e8 points to 1e-9 absolute
    (the property the reference asserts at 1e-12 on O(1) coordinates; coordinates here reach 125)."""
    torch = pytest.importorskip("torch")
    from interpn_b200 import workloads as W

    dev = torch.device("cuda:0")
    w = W.get("c2_cubic3d_reg100")
    n = w.n_full
    vals = w.vals("torch", dev)
    obs = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
    blk = 1 << 24
    for lo in range(0, n, blk):
        cnt = min(blk, n - lo)
        q = w.queries(lo, cnt, "torch", dev)
        for d in range(3):
            obs[d][lo : lo + cnt] = q[d]
    stream = torch.cuda.current_stream().cuda_stream
    with ib.Interpolator.regular("cubic", w.dims, w.starts, w.steps, vals, True) as interp:
        out = interp.eval_torch(obs)
        interp.status(stream)
    vals_h = vals.cpu().numpy()
    for sl in (slice(0, 100_000), slice(n - 100_000, n), slice(0, n, 997)):
        o = [x[sl].contiguous().cpu().numpy() for x in obs]
        want = oracle.interpn_regular("cubic", w.dims, w.starts, w.steps, vals_h, o, linearize_extrapolation=True, nthreads=8)
        assert_same_bits(out[sl].contiguous().cpu().numpy(), want, f"c2 sample {sl}")
    # property: linear field
    ax = [torch.from_numpy(g).to(dev) for g in w.grids]
    field = (1.5 * ax[0][:, None, None] - 0.25 * ax[1][None, :, None] + 3.0 * ax[2][None, None, :] + 7.0).contiguous().reshape(-1)
    with ib.Interpolator.regular("cubic", w.dims, w.starts, w.steps, field, True) as interp:
        out = interp.eval_torch(obs, out)
        interp.status(stream)
    expect = 1.5 * obs[0] - 0.25 * obs[1] + 3.0 * obs[2] + 7.0
    err = (out - expect).abs().max().item()
    assert err < 1e-9, err
