#!/usr/bin/env python
"""Workload for compute-sanitizer (tools/gpu.sh sanitize): one small evaluation of every kernel family through the C ABI,
including the kernels that exchange data through shared memory behind __syncwarp() (cubic_quad4.cuh), the counting sort
of the bin-swept path (sweep.cuh) and the warp-compacting slab passes. Sizes are small because racecheck runs ~100x slower
than native. Every result is also compared with the oracle, so a sanitizer-clean run is a correct run."""

import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import interpn_b200 as ib  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.test_gpu_parity import assert_same_bits, random_case  # noqa: E402


def run(method, ndims, dtype, n, env=None, maxdim=None):
    old = {}
    for k, v in (env or {}).items():
        old[k] = os.environ.get(k)
        os.environ[k] = v
    try:
        rng = np.random.default_rng(ndims * 7 + len(method))
        lo = 4 if method == "cubic" else 2
        hi = maxdim or {1: 30, 2: 14, 3: 9, 4: 7, 5: 5, 6: 4, 7: 4, 8: 4}[ndims]
        dims, grids, starts, steps, vals, obs = random_case(rng, ndims, n, lo, max(lo, hi), dtype)
        sfx = "f64" if dtype == np.float64 else "f32"
        extra = (True,) if method == "cubic" else ()
        out = np.zeros(n, dtype=dtype)
        getattr(ib.raw, f"interpn_{method}_regular_{sfx}")(dims, starts, steps, vals, *extra, obs, out)
        assert_same_bits(out, oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=True, nthreads=4))
        out = np.zeros(n, dtype=dtype)
        getattr(ib.raw, f"interpn_{method}_rectilinear_{sfx}")(grids, vals, *extra, obs, out)
        assert_same_bits(out, oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=True, nthreads=4))
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


SWEEP = {"INTERPN_B200_SWEEP_MIN_MB": "0", "INTERPN_B200_SWEEP_MIN_POINTS": "0", "INTERPN_B200_SWEEP_MIN_ROWS": "0",
         "INTERPN_B200_SWEEP_SLAB_KB": "0", "INTERPN_B200_SWEEP_CHUNK": "20000"}  # fmt: skip
SLAB = {"INTERPN_B200_WINDOW_MB": "0", "INTERPN_B200_SLAB_MIN_KB": "0", "INTERPN_B200_SLAB_MIN_POINTS": "0", "INTERPN_B200_SLAB_PASS_KB": "2"}
WIN = {"INTERPN_B200_WINDOW_MIN_KB": "0"}
HYPER = {"INTERPN_B200_HYPER_MIN_KB": "0"}
PLAIN = {"INTERPN_B200_WINDOW_MB": "0"}

n0 = ib.launch_count()
cases = 0
for dtype in (np.float64, np.float32):
    for method, nds in (("linear", (1, 2, 3, 4, 6, 8)), ("cubic", (1, 2, 3, 4, 5)), ("nearest", (1, 2, 3, 6))):
        for nd in nds:
            n = 3001 if not (method == "cubic" and nd >= 4) else 601
            run(method, nd, dtype, n, WIN)   # window / patch / coefficient layouts, quad4 kernels
            run(method, nd, dtype, n, PLAIN)  # straight from vals
            cases += 2
    for method, nd in (("linear", 3), ("linear", 6), ("cubic", 2), ("cubic", 3), ("cubic", 4)):
        run(method, nd, dtype, 40_001 if nd < 6 else 9001, dict(SWEEP, **WIN))  # counting sort + dynamic block schedule
        cases += 1
    for nd in (3, 4, 5):
        run("linear", nd, dtype, 30_011, SLAB, maxdim={3: 14, 4: 8, 5: 6}[nd])  # warp-compacting slab passes
        cases += 1
    for nd in (3, 4, 5, 6):
        run("linear", nd, dtype, 9001, HYPER)  # quad-cooperative hypercube kernels (shared-memory transposition behind __syncwarp)
        cases += 1
    run("linear", 3, dtype, 3001, {"INTERPN_B200_INDEX64": "1"})
    run("nearest", 3, dtype, 3001, {"INTERPN_B200_INDEX64": "1"})
    cases += 2
    # one_dim + check_bounds
    rng = np.random.default_rng(3)
    vals = rng.standard_normal(40).astype(dtype)
    grid = np.cumsum(rng.random(40) + 0.1).astype(dtype)
    locs = (rng.random(5003) * 24 - 2).astype(dtype)
    for kind in ("linear", "linear_hold_last", "left", "right", "nearest"):
        assert_same_bits(ib.one_dim.eval_regular(kind, dtype(0.0), dtype(0.5), vals, locs), oracle.one_dim_regular(kind, dtype(0.0), dtype(0.5), vals, locs))
        assert_same_bits(ib.one_dim.eval_rectilinear(kind, grid, vals, locs), oracle.one_dim_rectilinear(kind, grid, vals, locs))
    flags = np.zeros(1, dtype=bool)
    getattr(ib.raw, f"check_bounds_rectilinear_{'f64' if dtype == np.float64 else 'f32'}")([grid], [locs], dtype(1e-6), flags)
    assert flags[0]
print(f"sanitize_run: {cases} cases x2 grids bit-identical to the oracle, {ib.launch_count() - n0} kernel launches")
