"""Every compiled instantiation executes at least once, bit-compared with the oracle (VERDICT r1 "untested
instantiations": rows a16, a19, a22, a23, a24, f1 of SURVEY.md §8).

* multicubic N = 7, 8 regular + rectilinear (ref: multicubic/regular_recursive.rs:47-131, rectilinear_recursive.rs:46-...)
* the 64-bit index kernels (`I = long long`; ref: lib.rs:119-144 indexes with usize): forced onto small grids with
  INTERPN_B200_INDEX64=1, and one real grid of more than 2^31 f32 values
* one_dim: 5 kinds x regular/rectilinear x f32/f64 on 1e6 random locations incl. out-of-range, nodes, NaN and
  infinities (ref: one_dim/mod.rs:85-187, linear.rs:96-179, hold.rs:118-179), many CTAs, error path
* check_bounds on 1e7 points: violations planted in the last CTA only, the `<= -atol` / `>= atol` edges, NaN
  (ref: multilinear/regular.rs:145-182, multilinear/rectilinear.rs:109-134)

Needs a B200: `pytest -m gpu`. All calls go through the C ABI.
"""

import numpy as np
import pytest

from tests.test_gpu_parity import assert_same_bits, random_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import interpn_b200

    return interpn_b200


# ------------------------------------------------------------------------------------------------ cubic N = 7, 8
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("ndims", [7, 8])
def test_cubic_seven_and_eight_dimensions_bit_exact(ib, oracle, ndims, dtype):
    rng = np.random.default_rng(4200 + ndims)
    n = 200
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, 4, 4, dtype)
    sfx = "f64" if dtype == np.float64 else "f32"
    before = ib.launch_count()
    for lin in (False, True):
        out = np.zeros(n, dtype=dtype)
        getattr(ib.raw, f"interpn_cubic_regular_{sfx}")(dims, starts, steps, vals, lin, obs, out)
        want = oracle.interpn_regular("cubic", dims, starts, steps, vals, obs, linearize_extrapolation=lin, nthreads=8)
        assert_same_bits(out, want, f"cubic regular N={ndims} {sfx} lin={lin}")
        out = np.zeros(n, dtype=dtype)
        getattr(ib.raw, f"interpn_cubic_rectilinear_{sfx}")(grids, vals, lin, obs, out)
        want = oracle.interpn_rectilinear("cubic", grids, vals, obs, linearize_extrapolation=lin, nthreads=8)
        assert_same_bits(out, want, f"cubic rectilinear N={ndims} {sfx} lin={lin}")
    assert ib.launch_count() >= before + 4


def test_cubic_nine_dimensions_is_refused_like_the_reference(ib):
    """multicubic/regular.rs:98-129: `_ => Err("Dimension exceeds maximum (8)...")`."""
    dims = [4] * 9
    vals = np.zeros(4**9)
    obs = [np.zeros(3)] * 9
    with pytest.raises(AssertionError, match="Dimension exceeds maximum"):
        ib.raw.interpn_cubic_regular_f64(dims, np.zeros(9), np.ones(9), vals, True, obs, np.zeros(3))


# ------------------------------------------------------------------------------------------------ 64-bit index path
INDEX64_CASES = [(m, n, dt) for dt in (np.float64, np.float32) for m, dims_ in (("linear", (1, 2, 3, 4, 6, 8)), ("nearest", (1, 2, 3, 6)), ("cubic", (1, 3, 5)))
                 for n in dims_]  # fmt: skip


@pytest.mark.parametrize("method,ndims,dtype", INDEX64_CASES, ids=lambda v: getattr(v, "__name__", str(v)))
def test_forced_64bit_index_kernels_bit_exact(ib, oracle, monkeypatch, method, ndims, dtype):
    """INTERPN_B200_INDEX64=1 sends every grid through the kernels that index `vals` with 64-bit arithmetic
    (linear_kernel<.., long long>, nearest_kernel<.., long long>, the plain cubic kernel), with window copies, the
    bin-swept path, slab passes and fused fields switched off, exactly as for a grid of 2^31 values or more."""
    monkeypatch.setenv("INTERPN_B200_INDEX64", "1")
    rng = np.random.default_rng(6400 + 10 * ndims + len(method))
    min_dim = 4 if method == "cubic" else 2
    max_dim = {1: 40, 2: 20, 3: 12, 4: 8, 5: 5, 6: 5, 8: 4}[ndims]
    n = 20000 if (method != "cubic" or ndims <= 3) else 2000
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, min_dim, max(max_dim, min_dim), dtype)
    sfx = "f64" if dtype == np.float64 else "f32"
    extra = (True,) if method == "cubic" else ()
    swept0 = ib.swept_launch_count()
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_{method}_regular_{sfx}")(dims, starts, steps, vals, *extra, obs, out)
    assert_same_bits(out, oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=True, nthreads=4))
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_{method}_rectilinear_{sfx}")(grids, vals, *extra, obs, out)
    assert_same_bits(out, oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=True, nthreads=4))
    assert ib.swept_launch_count() == swept0
    # failure semantics are the same kernels' business
    if method != "cubic":
        obs2 = [o.copy() for o in obs]
        obs2[-1][n // 2] = np.nan
        out = np.full(n, -7.0, dtype=dtype)
        with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
            getattr(ib.raw, f"interpn_{method}_regular_{sfx}")(dims, starts, steps, vals, obs2, out)
        assert np.all(out[n // 2 :] == -7.0) and not np.any(out[: n // 2] == -7.0)


def test_forced_64bit_index_env_really_changes_the_kernel(ib, monkeypatch):
    """Launch accounting: the hook must disable the bin-swept path (its kernels index with int)."""
    rng = np.random.default_rng(1)
    dims, grids, starts, steps, vals, obs = random_case(rng, 4, 300_000, 6, 9, np.float64)
    for k, v in (("INTERPN_B200_SWEEP_MIN_MB", "0"), ("INTERPN_B200_SWEEP_MIN_POINTS", "0"), ("INTERPN_B200_SWEEP_MIN_ROWS", "0"),
                 ("INTERPN_B200_SWEEP_SLAB_KB", "0")):  # fmt: skip
        monkeypatch.setenv(k, v)
    out = np.zeros(300_000)
    s0 = ib.swept_launch_count()
    ib.raw.interpn_linear_regular_f64(dims, starts, steps, vals, obs, out)
    assert ib.swept_launch_count() > s0
    monkeypatch.setenv("INTERPN_B200_INDEX64", "1")
    out64 = np.zeros(300_000)
    s0 = ib.swept_launch_count()
    ib.raw.interpn_linear_regular_f64(dims, starts, steps, vals, obs, out64)
    assert ib.swept_launch_count() == s0
    assert_same_bits(out64, out)


def test_real_grid_beyond_2_pow_31_values(ib):
    """A REAL grid of 2^31 + 2^26 f32 values (8.9 GB of the 180 GB): nearest, multilinear and multicubic. The oracle cannot
    hold it in seconds, so the check is the size-independent one: vals[i] = a linear field of the node indices,
    evaluated through float64-exact host arithmetic — nearest must return exactly the field at the rounded node
    (bit-exact: the values are small integers), linear/cubic the field itself at interior points within f32 tolerance.
    Queries concentrate on the last rows of dimension 0, where flat indices exceed 2^31."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    d0, d1, d2 = 2112, 1024, 1024  # 2 214 592 512 values > 2^31
    assert d0 * d1 * d2 > 2**31
    # field: f(i,j,k) = (i mod 64) + 2*(j mod 32) + 3*(k mod 16): exactly representable, piecewise linear
    i = torch.arange(d0, device=dev, dtype=torch.float32) % 64
    j = torch.arange(d1, device=dev, dtype=torch.float32) % 32
    k = torch.arange(d2, device=dev, dtype=torch.float32) % 16
    vals = (i[:, None, None] + 2 * j[None, :, None] + 3 * k[None, None, :]).contiguous().reshape(-1)
    del i, j, k
    starts = np.zeros(3, dtype=np.float32)
    steps = np.ones(3, dtype=np.float32)
    n = 200_000
    rng = np.random.default_rng(9)
    # cells whose 4-point cubic footprint stays inside one linear piece: index mod 16 in [1, 13] on every axis
    def inner(dim, lo):
        base = rng.integers(lo // 16, dim // 16, size=n) * 16
        return base + rng.integers(1, 13, size=n) + rng.random(n)
    q = [inner(d0, 2048).astype(np.float32), inner(d1, 0).astype(np.float32), inner(d2, 0).astype(np.float32)]
    flat = np.floor(q[0].astype(np.float64)) * d1 * d2
    assert (flat >= 2**31).all()
    qd = [x.astype(np.float64) for x in q]
    field = (qd[0] % 64) + 2 * (qd[1] % 32) + 3 * (qd[2] % 16)
    near = (np.rint(np.nextafter(qd[0], -np.inf)) % 64) + 2 * (np.rint(np.nextafter(qd[1], -np.inf)) % 32) + 3 * (np.rint(np.nextafter(qd[2], -np.inf)) % 16)
    for method in ("nearest", "linear", "cubic"):
        with ib.Interpolator.regular(method, [d0, d1, d2], starts, steps, vals, True) as it:
            got = it.eval(q).astype(np.float64)
        if method == "nearest":
            # ties (x.5 exactly) are measure-zero here; dt <= 0.5 -> lower node (nearest/regular.rs:277-287)
            frac = qd[0] - np.floor(qd[0]), qd[1] - np.floor(qd[1]), qd[2] - np.floor(qd[2])
            want = ((np.floor(qd[0]) + (frac[0] > 0.5)) % 64) + 2 * ((np.floor(qd[1]) + (frac[1] > 0.5)) % 32) + 3 * ((np.floor(qd[2]) + (frac[2] > 0.5)) % 16)
            assert np.array_equal(got, want)
        else:
            assert np.max(np.abs(got - field)) < 2e-4, method
    del near


# ------------------------------------------------------------------------------------------------ one_dim at scale
KINDS = ["linear", "linear_hold_last", "left", "right", "nearest"]


def _one_dim_locs(rng, grid, n, dtype):
    lo, hi = float(grid[0]), float(grid[-1])
    span = hi - lo
    x = lo - 0.25 * span + 1.5 * span * rng.random(n)
    sel = rng.random(n)
    k = rng.integers(0, grid.size, size=n)
    x = np.where(sel < 0.1, grid[k].astype(np.float64), x)  # exact nodes
    k2 = np.minimum(k, grid.size - 2)
    x = np.where((sel >= 0.1) & (sel < 0.15), 0.5 * (grid[k2].astype(np.float64) + grid[k2 + 1].astype(np.float64)), x)  # ties
    x = x.astype(dtype)
    x[:4] = [grid[0], grid[-1], np.nextafter(grid[0], -np.inf, dtype=dtype), np.nextafter(grid[-1], np.inf, dtype=dtype)]
    return np.ascontiguousarray(x)


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("kind", KINDS)
def test_one_dim_at_scale_bit_exact(ib, oracle, kind, dtype):
    rng = np.random.default_rng(1100 + KINDS.index(kind))
    n = 1_000_003  # > 148 * 8 CTAs of 256 threads: the grid-stride loop runs
    nv = 257
    start, step = dtype(-3.25), dtype(0.0625)
    vals = rng.standard_normal(nv).astype(dtype)
    reg_grid = (start + step * np.arange(nv, dtype=dtype)).astype(dtype)
    locs = _one_dim_locs(rng, reg_grid, n, dtype)
    got = ib.one_dim.eval_regular(kind, start, step, vals, locs)
    assert_same_bits(got, oracle.one_dim_regular(kind, start, step, vals, locs), f"one_dim regular {kind}")
    grid = np.ascontiguousarray((np.cumsum(rng.random(nv) + 0.05) - 7.0).astype(dtype))
    assert np.all(np.diff(grid) > 0)
    locs = _one_dim_locs(rng, grid, n, dtype)
    locs[100], locs[101], locs[102] = np.nan, np.inf, -np.inf  # rectilinear grids never fail (partition_point)
    got = ib.one_dim.eval_rectilinear(kind, grid, vals, locs)
    want = oracle.one_dim_rectilinear(kind, grid, vals, locs)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert_same_bits(got[ok], want[ok], f"one_dim rectilinear {kind}")


@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
@pytest.mark.parametrize("bad", [np.nan, np.inf, -np.inf])
def test_one_dim_regular_error_path(ib, oracle, bad, dtype):
    """RegularGrid1D::index_of: `<isize as NumCast>::from(...).ok_or("Unrepresentable number")` (one_dim/mod.rs:105-112);
    the serial loop stops at the first failing location: earlier outputs written, later untouched."""
    rng = np.random.default_rng(5)
    n = 400_000
    vals = rng.standard_normal(50).astype(dtype)
    locs = (rng.random(n) * 60 - 5).astype(dtype)
    locs[333_333] = bad
    locs[390_000] = np.nan
    out = np.full(n, -9.0, dtype=dtype)
    with pytest.raises(AssertionError, match="Unrepresentable number"):
        ib.one_dim.eval_regular("linear", dtype(0.0), dtype(1.0), vals, locs, out)
    want = np.full(n, -9.0, dtype=dtype)
    with pytest.raises(oracle.OracleError, match="Unrepresentable number"):
        oracle.one_dim_regular("linear", dtype(0.0), dtype(1.0), vals, locs, want)
    assert_same_bits(out, want)
    assert np.all(out[333_333:] == -9.0)
    # constructor errors (one_dim/mod.rs:53, 86-88, 150)
    with pytest.raises(AssertionError, match="Length mismatch"):
        ib.one_dim.eval_rectilinear("left", np.arange(5, dtype=dtype), np.zeros(6, dtype=dtype), locs[:10].copy())


# ------------------------------------------------------------------------------------------------ check_bounds at scale
@pytest.mark.parametrize("dtype", [np.float64, np.float32], ids=["f64", "f32"])
def test_check_bounds_at_scale(ib, oracle, dtype):
    rng = np.random.default_rng(77)
    n = 10_000_019
    dims = [11, 7, 5]
    starts = np.array([-1.0, 2.0, 0.5], dtype=dtype)
    steps = np.array([0.5, 0.25, 2.0], dtype=dtype)
    grids = [np.ascontiguousarray((starts[d] + steps[d] * np.arange(dims[d])).astype(dtype)) for d in range(3)]
    hi = [g[-1] for g in grids]
    atol = dtype(0.125)  # exactly representable: the edge cases below are exact in both dtypes
    inside = [np.ascontiguousarray((starts[d] + (hi[d] - starts[d]) * rng.random(n)).astype(dtype)) for d in range(3)]

    def both(obs, expect):
        got_r = np.zeros(3, dtype=bool)
        getattr(ib.raw, f"check_bounds_regular_{'f64' if dtype == np.float64 else 'f32'}")(dims, starts, steps, obs, atol, got_r)
        got_x = np.zeros(3, dtype=bool)
        getattr(ib.raw, f"check_bounds_rectilinear_{'f64' if dtype == np.float64 else 'f32'}")(grids, obs, atol, got_x)
        want_r = oracle.check_bounds_regular(dims, starts, steps, obs, float(atol))
        want_x = oracle.check_bounds_rectilinear(grids, obs, float(atol))
        assert list(got_r) == list(want_r) == expect, (got_r, want_r, expect)
        assert list(got_x) == list(want_x) == expect, (got_x, want_x, expect)

    both(inside, [False, False, False])
    # one violation, in the very last element (last CTA, last warp, ragged tail), on axis 1 only
    obs = [o.copy() for o in inside]
    obs[1][-1] = hi[1] + dtype(1.0)
    both(obs, [False, True, False])
    # the comparison is `(x - lo) <= -atol || (x - hi) >= atol` (multilinear/regular.rs:168-171): the edge itself violates,
    # one ulp inside does not
    obs = [o.copy() for o in inside]
    obs[0][n // 2] = starts[0] - atol  # == -atol -> bad
    obs[2][n - 7] = np.nextafter(hi[2] + atol, -np.inf, dtype=dtype)  # just below +atol -> fine
    both(obs, [True, False, False])
    obs = [o.copy() for o in inside]
    obs[0][5] = np.nextafter(starts[0] - atol, np.inf, dtype=dtype)  # just above -atol -> fine
    obs[2][n - 300] = hi[2] + atol  # == atol -> bad
    both(obs, [False, False, True])
    # NaN compares false on both sides: never a violation; infinities are
    obs = [o.copy() for o in inside]
    obs[0][123_456] = np.nan
    both(obs, [False, False, False])
    obs[1][9_999_999] = -np.inf
    obs[2][0] = np.inf
    both(obs, [False, True, True])


# ------------------------------------------------------------------------------------------------ extreme axes
@pytest.mark.parametrize("method", ["linear", "cubic", "nearest"])
@pytest.mark.parametrize("kind", ["huge_span", "subnormal_span"])
def test_rectilinear_axes_whose_span_leaves_the_bucket_tables_range(ib, oracle, method, kind):
    """f32 axes that are strictly increasing and finite, but whose span overflows (or is subnormal) in f32: the bucket
    tables of the rectilinear search would compute garbage buckets, so such grids must take the plain bisection and still
    match the reference bit for bit (capi.cu rect_new)."""
    rng = np.random.default_rng(12)
    n = 20000
    if kind == "huge_span":
        ax = np.linspace(-3.0e38, 3.0e38, 9).astype(np.float32)
        q = (rng.random(n) * 6.4e38 - 3.2e38).astype(np.float32)
    else:
        ax = (np.arange(9, dtype=np.float64) * 3e-43).astype(np.float32)  # subnormal nodes, subnormal span
        q = (rng.random(n) * 2.6e-42 - 1e-43).astype(np.float32)
    assert np.all(np.diff(ax.astype(np.float64)) > 0) and np.all(np.isfinite(ax))
    grids = [ax, np.linspace(0.0, 1.0, 6).astype(np.float32)]
    vals = rng.standard_normal(9 * 6).astype(np.float32)
    obs = [q, rng.random(n).astype(np.float32) * 1.2 - 0.1]
    obs[0][:9] = ax
    extra = (True,) if method == "cubic" else ()
    out = np.zeros(n, dtype=np.float32)
    getattr(ib.raw, f"interpn_{method}_rectilinear_f32")(grids, vals, *extra, obs, out)
    want = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=True, nthreads=4)
    both_nan = np.isnan(out) & np.isnan(want)
    assert np.all((out.view(np.uint32) == want.view(np.uint32)) | both_nan)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("method,ndims", [("linear", 2), ("linear", 4), ("linear", 6), ("nearest", 2), ("nearest", 3),
                                          ("cubic", 2), ("cubic", 3), ("cubic", 4)])  # fmt: skip
def test_rectilinear_axes_in_global_memory_bit_exact(ib, oracle, monkeypatch, method, ndims, dtype):
    """Every rectilinear kernel exists twice: with the axes blob (axes, bucket / cell tables) staged in shared memory and
    with the blob read from global memory (blobs beyond the 96 KB budget). INTERPN_B200_AXES_SMEM_KB=0 forces the second
    instantiation onto small grids — incl. the quad-cooperative cubic kernels and the hypercube multilinear kernels — and a
    2-D grid with 3 000-node axes takes it for real."""
    monkeypatch.setenv("INTERPN_B200_AXES_SMEM_KB", "0")
    monkeypatch.setenv("INTERPN_B200_HYPER_MIN_KB", "0")
    rng = np.random.default_rng(8123 + ndims)
    n = 60_007
    lo = 4 if method == "cubic" else 2
    hi = {2: 40, 3: 12, 4: 8, 6: 5}[ndims]
    dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, lo, hi, dtype)
    sfx = "f64" if dtype == np.float64 else "f32"
    extra = (True,) if method == "cubic" else ()
    out = np.zeros(n, dtype=dtype)
    getattr(ib.raw, f"interpn_{method}_rectilinear_{sfx}")(grids, vals, *extra, obs, out)
    assert_same_bits(out, oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=True, nthreads=8))
    if ndims == 2:  # axes too long for shared memory without any hook
        monkeypatch.delenv("INTERPN_B200_AXES_SMEM_KB")
        big = [np.cumsum(rng.random(3000) + 0.05).astype(dtype) for _ in range(2)]
        bvals = rng.standard_normal(3000 * 3000).astype(dtype)
        bobs = [(rng.random(n) * (g[-1] - g[0]) * 1.1 + g[0] - 0.05 * (g[-1] - g[0])).astype(dtype) for g in big]
        out = np.zeros(n, dtype=dtype)
        getattr(ib.raw, f"interpn_{method}_rectilinear_{sfx}")(big, bvals, *extra, bobs, out)
        assert_same_bits(out, oracle.interpn_rectilinear(method, big, bvals, bobs, linearize_extrapolation=True, nthreads=8))
