#!/usr/bin/env python
"""Extracts tests/golden/ref_docs.npz: inputs AND outputs of the reference wheel itself.

The reference's documentation figures embed, as full-precision Plotly JSON, the arrays that its own
example scripts fed to and received from the published `interpn` wheel (an `fma`-feature build,
`/root/reference/pyproject.toml:72`):

* `docs/1d_quality_of_fit_{Regular,Rectilinear}.html` and `docs/2d_quality_of_fit_{...}.html`, written by
  `/root/reference/examples/cubic_comparison.py:50-87` (1-D: `MulticubicRegular.new(dims, starts, steps, y,
  linearize_extrapolation=False).eval([x])` / `MulticubicRectilinear.new([xdata], ydata, ...)` for a
  quadratic, a sine and a step; traces "Data" = grid, "InterpN" = query x and result y) and `:221-257`
  (2-D: 7x7 grid of x^2+y^2, 30x30 queries on [-5,5]^2; traces "Sampled data" = grid nodes, heatmap
  "InterpN" = result transposed);
* `docs/nearest_quality_of_fit.html`, written by `/root/reference/examples/nearest_comparison.py:66-82`
  (`NearestRectilinear.new([xdata, ydata], zmesh).eval(...)` on a 25x18 irregular grid, 160x160 queries;
  traces "Grid samples" and heatmap "InterpN").

These are the only outputs of the real crate held anywhere in the reference tree; this script runs in the
build container (where /root/reference exists) and commits them as a fixture, so the oracle and the CUDA `fma`
library can be compared with them on the GPU box, where the reference is absent.

Grid VALUES are not in the figures for the 2-D cases (only node positions); they are recomputed with the
example's own expression. For x^2+y^2 this is exactly rounded arithmetic. For the nearest case the expression
calls sin/cos, so the script additionally proves the recomputation: every reference output must be, bit for bit,
one of the recomputed grid values (it is: the queries hit all 450 nodes).

    python tests/golden/extract_ref_docs.py      # rewrites tests/golden/ref_docs.npz (deterministic)
"""

import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
DOCS = os.environ.get("INTERPN_REFERENCE_DOCS", "/root/reference/docs")


def traces(name):
    s = open(os.path.join(DOCS, name)).read()
    i = s.index("Plotly.newPlot(")
    j = s.index("[", i)
    data, _ = json.JSONDecoder().raw_decode(s[j:])
    return data


def f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def main():
    out = {}
    meta = []

    def add(name, method, rect, linearize, vals, obs, result, grids=None, dims=None, starts=None, steps=None, source=""):
        meta.append({"name": name, "method": method, "rect": rect, "linearize": linearize, "ndims": len(obs), "source": source})
        out[f"{name}/vals"] = f64(vals).reshape(-1)
        out[f"{name}/out"] = f64(result).reshape(-1)
        for d, o in enumerate(obs):
            out[f"{name}/obs{d}"] = f64(o).reshape(-1)
        if rect:
            for d, g in enumerate(grids):
                out[f"{name}/grid{d}"] = f64(g)
        else:
            out[f"{name}/dims"] = np.asarray(dims, dtype=np.int64)
            out[f"{name}/starts"] = f64(starts)
            out[f"{name}/steps"] = f64(steps)

    # ---- 1-D cubic: three functions per figure, subplot column i uses axes x{i+1}/y{i+1} (row 1)
    for kind in ("Regular", "Rectilinear"):
        tr = traces(f"1d_quality_of_fit_{kind}.html")
        for col, fn in enumerate(("quadratic", "sine", "step")):
            ax = "x" if col == 0 else f"x{col + 1}"
            data = next(t for t in tr if t.get("name") == "Data" and t["xaxis"] == ax)
            res = next(t for t in tr if t.get("name") == "InterpN" and t["xaxis"] == ax)
            xdata, ydata = f64(data["x"]), f64(data["y"])
            name = f"cubic1d_{kind.lower()}_{fn}"
            src = f"docs/1d_quality_of_fit_{kind}.html; examples/cubic_comparison.py:62-87"
            if kind == "Regular":
                assert np.array_equal(xdata, np.arange(-2.0, 2.5, 0.5))
                add(name, "cubic", False, False, ydata, [res["x"]], res["y"], dims=[xdata.size], starts=[-2.0], steps=[0.5], source=src)
            else:
                add(name, "cubic", True, False, ydata, [res["x"]], res["y"], grids=[xdata], source=src)

    # ---- 2-D cubic: 7x7 grid, meshgrid(indexing="ij") flattened in the "Sampled data" trace
    for kind in ("Regular", "Rectilinear"):
        tr = traces(f"2d_quality_of_fit_{kind}.html")
        samp = next(t for t in tr if t.get("name") == "Sampled data")
        xm, ym = f64(samp["x"]).reshape(7, 7), f64(samp["y"]).reshape(7, 7)
        xdata, ydata = xm[:, 0].copy(), ym[0, :].copy()
        assert np.array_equal(xm, np.broadcast_to(xdata[:, None], (7, 7))) and np.array_equal(ym, np.broadcast_to(ydata[None, :], (7, 7)))
        zmesh = xm**2 + ym**2
        heat = next(t for t in tr if t.get("name") == "InterpN" and t.get("type") == "heatmap")
        xi, yi = f64(heat["x"]), f64(heat["y"])
        z = f64(heat["z"]).T  # the figure holds z_interpn.T
        xim, yim = np.meshgrid(xi, yi, indexing="ij")
        name = f"cubic2d_{kind.lower()}"
        src = f"docs/2d_quality_of_fit_{kind}.html; examples/cubic_comparison.py:221-257"
        if kind == "Regular":
            steps = [xm[1, 0] - xm[0, 0], ym[0, 1] - ym[0, 0]]
            add(name, "cubic", False, False, zmesh, [xim.flatten(), yim.flatten()], z, dims=[7, 7], starts=[-3.0, -3.0], steps=steps, source=src)
        else:
            add(name, "cubic", True, False, zmesh, [xim.flatten(), yim.flatten()], z, grids=[xdata, ydata], source=src)

    # ---- 2-D rectilinear nearest
    tr = traces("nearest_quality_of_fit.html")
    samp = next(t for t in tr if t.get("name") == "Grid samples")
    xm, ym = f64(samp["x"]).reshape(25, 18), f64(samp["y"]).reshape(25, 18)
    xdata, ydata = xm[:, 0].copy(), ym[0, :].copy()
    zmesh = np.sin(xm) + 0.5 * np.cos(2.0 * ym) + 0.15 * xm * ym  # nearest_comparison.py:21-23
    heat = next(t for t in tr if t.get("name") == "InterpN" and t.get("type") == "heatmap")
    z = f64(heat["z"]).T
    # prove the sin/cos recomputation: the reference's outputs are grid values, bit for bit
    grid_bits = set(zmesh.view(np.uint64).reshape(-1).tolist())
    out_bits = set(z.view(np.uint64).reshape(-1).tolist())
    assert out_bits <= grid_bits, "recomputed grid values differ from the ones the reference saw"
    assert len(out_bits) == len(grid_bits) == 450, "queries must pin every node's value"
    xim, yim = np.meshgrid(f64(heat["x"]), f64(heat["y"]), indexing="ij")
    add("nearest2d_rectilinear", "nearest", True, True, zmesh, [xim.flatten(), yim.flatten()], z, grids=[xdata, ydata],
        source="docs/nearest_quality_of_fit.html; examples/nearest_comparison.py:66-82")  # fmt: skip

    out["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "ref_docs.npz"), **out)
    npts = sum(out[f"{m['name']}/out"].size for m in meta)
    print(f"{len(meta)} cases, {npts} reference-evaluated points written to ref_docs.npz")


if __name__ == "__main__":
    main()
