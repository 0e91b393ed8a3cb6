"""Engine adapter (see tests/refsuite.py) that drives the CUDA library through its
reference-shaped Python API (`interpn_b200.raw`, `interpn_b200.one_dim`), i.e. through the C ABI
with host buffers."""

import numpy as np


class GpuEngine:
    def __init__(self):
        import interpn_b200

        self.ib = interpn_b200

    @staticmethod
    def _sfx(dtype):
        return "f64" if np.dtype(dtype) == np.float64 else "f32"

    def regular(self, method, dims, starts, steps, vals, obs, linearize=True):
        vals = np.ascontiguousarray(vals)
        dt = vals.dtype
        obs = [np.ascontiguousarray(o, dtype=dt) for o in obs]
        out = np.zeros(obs[0].size, dtype=dt)
        fn = getattr(self.ib.raw, f"interpn_{method}_regular_{self._sfx(dt)}")
        starts, steps = np.ascontiguousarray(starts, dtype=dt), np.ascontiguousarray(steps, dtype=dt)
        if method == "cubic":
            fn(list(dims), starts, steps, vals, linearize, obs, out)
        else:
            fn(list(dims), starts, steps, vals, obs, out)
        return out

    def rectilinear(self, method, grids, vals, obs, linearize=True):
        vals = np.ascontiguousarray(vals)
        dt = vals.dtype
        grids = [np.ascontiguousarray(g, dtype=dt) for g in grids]
        obs = [np.ascontiguousarray(o, dtype=dt) for o in obs]
        out = np.zeros(obs[0].size, dtype=dt)
        fn = getattr(self.ib.raw, f"interpn_{method}_rectilinear_{self._sfx(dt)}")
        if method == "cubic":
            fn(grids, vals, linearize, obs, out)
        else:
            fn(grids, vals, obs, out)
        return out

    def one_dim_regular(self, kind, start, step, vals, locs):
        vals = np.ascontiguousarray(vals)
        return self.ib.one_dim.eval_regular(kind, start, step, vals, np.ascontiguousarray(locs, dtype=vals.dtype))

    def one_dim_rectilinear(self, kind, grid, vals, locs):
        vals = np.ascontiguousarray(vals)
        dt = vals.dtype
        return self.ib.one_dim.eval_rectilinear(
            kind, np.ascontiguousarray(grid, dtype=dt), vals, np.ascontiguousarray(locs, dtype=dt)
        )

    def check_bounds_regular(self, dims, starts, steps, obs, atol):
        dt = np.asarray(starts).dtype
        out = np.zeros(len(dims), dtype=bool)
        fn = getattr(self.ib.raw, f"check_bounds_regular_{self._sfx(dt)}")
        fn(list(dims), np.ascontiguousarray(starts), np.ascontiguousarray(steps),
           [np.ascontiguousarray(o, dtype=dt).reshape(-1) for o in obs], atol, out)  # fmt: skip
        return out

    def check_bounds_rectilinear(self, grids, obs, atol):
        dt = np.asarray(grids[0]).dtype
        out = np.zeros(len(grids), dtype=bool)
        fn = getattr(self.ib.raw, f"check_bounds_rectilinear_{self._sfx(dt)}")
        fn([np.ascontiguousarray(g) for g in grids], [np.ascontiguousarray(o, dtype=dt).reshape(-1) for o in obs], atol, out)
        return out
