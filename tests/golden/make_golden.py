#!/usr/bin/env python
"""Generates tests/golden/interpn_golden.npz: small, fully specified input/output vectors for
every (method, grid kind, dtype) of the hot path, in strict (crate default features) arithmetic.

The vectors are produced by the CPU oracle (oracle/), because the reference itself (pure Rust)
cannot be built or imported in this image; they pin the oracle against regressions and give a
maintainer with a Rust toolchain something to diff the crate against:
each case stores exactly the arguments of the reference's `interpn(...)` call and its `out`.

    python tests/golden/make_golden.py        # rewrites the .npz (deterministic)
"""

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = []  # (name, method, rect, ndims, dtype, linearize)
for dt in ("f64", "f32"):
    for method, nd_list in (("linear", (1, 2, 3, 4, 6, 8)), ("cubic", (1, 2, 3, 4, 5)), ("nearest", (1, 2, 3, 6))):
        for nd in nd_list:
            for rect in (False, True):
                for lin in ((False, True) if method == "cubic" else (True,)):
                    CASES.append((f"{method}_{'rect' if rect else 'reg'}_{nd}d_{dt}_lin{int(lin)}", method, rect, nd, dt, lin))


def build_case(seed, method, rect, nd, dt, n=96):
    rng = np.random.default_rng(seed)
    dtype = np.float64 if dt == "f64" else np.float32
    lo = 4 if method == "cubic" else 2
    hi = {1: 12, 2: 9, 3: 7, 4: 6, 5: 5, 6: 4, 8: 3}[nd]
    dims = [int(rng.integers(lo, max(lo, hi) + 1)) for _ in range(nd)]
    grids = [np.ascontiguousarray((np.cumsum(rng.random(d) + 0.1) + rng.normal()).astype(dtype)) for d in dims]
    starts = np.array([g[0] for g in grids], dtype=dtype)
    steps = np.array([(g[-1] - g[0]) / (len(g) - 1) for g in grids], dtype=dtype)
    vals = rng.standard_normal(int(np.prod(dims))).astype(dtype)
    obs = []
    for d, g in enumerate(grids):
        span = float(g[-1] - g[0])
        x = float(g[0]) - 0.3 * span + 1.6 * span * rng.random(n)
        k = rng.integers(0, dims[d], size=n)
        sel = rng.random(n)
        x = np.where(sel < 0.1, (starts[d] + steps[d] * k.astype(dtype)).astype(np.float64), x)   # regular nodes
        x = np.where((sel >= 0.1) & (sel < 0.2), g[k].astype(np.float64), x)                      # rectilinear nodes
        k2 = np.minimum(k, dims[d] - 2)
        x = np.where((sel >= 0.2) & (sel < 0.25), 0.5 * (g[k2].astype(np.float64) + g[k2 + 1].astype(np.float64)), x)  # ties
        obs.append(np.ascontiguousarray(x.astype(dtype)))
    return dims, grids, starts, steps, vals, obs


def main():
    from oracle import oracle

    oracle.build()
    out = {}
    for seed, (name, method, rect, nd, dt, lin) in enumerate(CASES):
        dims, grids, starts, steps, vals, obs = build_case(seed, method, rect, nd, dt)
        if rect:
            want = oracle.interpn_rectilinear(method, grids, vals, obs, linearize_extrapolation=lin)
            for d, g in enumerate(grids):
                out[f"{name}/grid{d}"] = g
        else:
            want = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=lin)
            out[f"{name}/dims"] = np.asarray(dims, dtype=np.int64)
            out[f"{name}/starts"] = starts
            out[f"{name}/steps"] = steps
        out[f"{name}/vals"] = vals
        for d, o in enumerate(obs):
            out[f"{name}/obs{d}"] = o
        out[f"{name}/out"] = want
    np.savez_compressed(os.path.join(HERE, "interpn_golden.npz"), **out)
    print(f"{len(CASES)} cases written")


if __name__ == "__main__":
    main()
