"""Deterministic synthetic grids and query batches for the BASELINE.json configurations.

Everything is derived from a counter-based generator, u(seed, i) = (splitmix64(seed*2^40 + i) >> 11) * 2^-53,
so the CPU (numpy) and the GPU (torch on the device) regenerate bit-identical inputs for any index
range without transfers. Used by bench.py and by the parity tests; no arithmetic
of the interpolation path lives here.

Configurations (names used in bench.py's `config.workload`):
  c1_linear3d_reg20     3-D multilinear, regular 20^3 f64, queries U(-0.99, 0.99)          (benches/bench_cpu.py:652-656)
  c2_cubic3d_reg100     3-D multicubic, regular 100^3 f64, linearize, 10 % out of bounds    (benches/bench.rs:299,538)
  c3_linear4d_rect64    4-D multilinear, rectilinear 64^4 f64 with jittered axes
  c3_cubic4d_rect64     4-D multicubic on the same grid
  c4_linear6d_reg24     6-D multilinear, regular 24^6 f64 (1.53 GB of values)
  c5_nearest2d_reg1024 / c5_nearest3d_reg128 / c5_nearest2d_rect1024 / c5_nearest3d_rect128 (f32 and f64)
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_M64 = (1 << 64) - 1
_GOLDEN = 0x9E3779B97F4A7C15
_MUL1 = 0xBF58476D1CE4E5B9
_MUL2 = 0x94D049BB133111EB


def _to_i64(u: int) -> int:
    u &= _M64
    return u - (1 << 64) if u >= (1 << 63) else u


def splitmix64_np(x: np.ndarray) -> np.ndarray:
    """splitmix64 finaliser on a uint64 array (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = x + np.uint64(_GOLDEN)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(_MUL1)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(_MUL2)
        return z ^ (z >> np.uint64(31))


def uniform_np(seed: int, start: int, count: int) -> np.ndarray:
    """float64 in [0, 1) for counters start .. start+count."""
    idx = np.arange(start, start + count, dtype=np.uint64) + np.uint64((seed << 40) & _M64)
    return (splitmix64_np(idx) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


def hash_np(seed: int, start: int, count: int) -> np.ndarray:
    idx = np.arange(start, start + count, dtype=np.uint64) + np.uint64((seed << 40) & _M64)
    return splitmix64_np(idx)


def _lsr(z, k: int):
    """Logical shift right on a torch int64 tensor."""
    return (z >> k) & ((1 << (64 - k)) - 1)


def hash_torch(seed: int, start: int, count: int, device):
    """Same bits as hash_np, as an int64 torch tensor on `device`."""
    import torch

    z = torch.arange(start, start + count, dtype=torch.int64, device=device)
    z = z + _to_i64(((seed << 40) & _M64) + _GOLDEN)
    z = (z ^ _lsr(z, 30)) * _to_i64(_MUL1)
    z = (z ^ _lsr(z, 27)) * _to_i64(_MUL2)
    return z ^ _lsr(z, 31)


def uniform_torch(seed: int, start: int, count: int, device):
    import torch

    return _lsr(hash_torch(seed, start, count, device), 11).to(torch.float64) * (1.0 / (1 << 53))


def _uniform(xp: str, seed: int, start: int, count: int, device=None):
    return uniform_np(seed, start, count) if xp == "np" else uniform_torch(seed, start, count, device)


def _hash_mod(xp: str, seed: int, start: int, count: int, mod: int, device=None):
    """(hash >> 33) % mod as a small non-negative integer array."""
    if xp == "np":
        return ((hash_np(seed, start, count) >> np.uint64(33)) % np.uint64(mod)).astype(np.int64)
    return _lsr(hash_torch(seed, start, count, device), 33) % mod


@dataclass
class Workload:
    name: str
    method: str                      # "linear" | "cubic" | "nearest"
    rect: bool
    dtype: np.dtype
    dims: list[int]
    n_full: int                      # query points of the BASELINE.json configuration
    linearize: bool = True
    oob_fraction: float = 0.0        # fraction of points with one axis pushed outside the grid
    starts: np.ndarray | None = None
    steps: np.ndarray | None = None
    grids: list[np.ndarray] = field(default_factory=list)   # axes (always filled, also for regular grids)
    lo: np.ndarray | None = None     # query box
    hi: np.ndarray | None = None
    vals_seed: int = 1
    vals_lo: float = 0.0
    vals_hi: float = 1.0
    obs_seed: int = 2
    plant_special: bool = False      # overwrite a sparse subset with exact nodes / exact ties (nearest)

    @property
    def ndims(self) -> int:
        return len(self.dims)

    @property
    def nvals(self) -> int:
        return int(np.prod(self.dims, dtype=np.int64))

    def vals(self, xp: str = "np", device=None):
        """Grid values, flat C-order."""
        out_parts = []
        step = 1 << 24
        for lo in range(0, self.nvals, step):
            cnt = min(step, self.nvals - lo)
            u = _uniform(xp, self.vals_seed, lo, cnt, device)
            out_parts.append(self.vals_lo + (self.vals_hi - self.vals_lo) * u)
        if xp == "np":
            return np.concatenate(out_parts).astype(self.dtype)
        import torch

        return torch.cat(out_parts).to(torch.float32 if self.dtype == np.float32 else torch.float64)

    def queries(self, start: int, count: int, xp: str = "np", device=None):
        """Coordinates of query points start .. start+count, one array per dimension."""
        obs = []
        if self.oob_fraction > 0:
            ucls = _uniform(xp, self.obs_seed + 101, start, count, device)
            is_oob = ucls < self.oob_fraction
            axis = _hash_mod(xp, self.obs_seed + 102, start, count, self.ndims, device)
            side = _hash_mod(xp, self.obs_seed + 103, start, count, 2, device)
            far = _uniform(xp, self.obs_seed + 104, start, count, device)
        for d in range(self.ndims):
            u = _uniform(xp, self.obs_seed + 7 * d, start, count, device)
            lo, hi = float(self.lo[d]), float(self.hi[d])
            x = lo + (hi - lo) * u
            if self.oob_fraction > 0:
                span = hi - lo
                below = lo - 0.25 * span * far
                above = hi + 0.25 * span * far
                if xp == "np":
                    pushed = np.where(side == 0, below, above)
                    x = np.where(is_oob & (axis == d), pushed, x)
                else:
                    import torch

                    pushed = torch.where(side == 0, below, above)
                    x = torch.where(is_oob & (axis == d), pushed, x)
            if self.plant_special:
                x = self._plant(x, d, start, count, xp, device)
            if xp == "np":
                obs.append(np.ascontiguousarray(x.astype(self.dtype)))
            else:
                import torch

                obs.append(x.to(torch.float32 if self.dtype == np.float32 else torch.float64).contiguous())
        return obs

    def _plant(self, x, d: int, start: int, count: int, xp: str, device):
        """Every 64th point sits exactly on a grid node, every 64th+1 exactly half-way between two
        nodes (the nearest-neighbour tie) on axis d."""
        g = self.grids[d].astype(np.float64)
        node = _hash_mod(xp, self.obs_seed + 200 + d, start, count, len(g) - 1, device)
        if xp == "np":
            i = np.arange(start, start + count, dtype=np.int64)
            gx = g[node]
            mid = g[node] + 0.5 * (g[node + 1] - g[node])
            x = np.where(i % 64 == 0, gx, x)
            return np.where(i % 64 == 1, mid, x)
        import torch

        gt = torch.from_numpy(g).to(device)
        i = torch.arange(start, start + count, dtype=torch.int64, device=device)
        gx = gt[node]
        mid = gt[node] + 0.5 * (gt[node + 1] - gt[node])
        x = torch.where(i % 64 == 0, gx, x)
        return torch.where(i % 64 == 1, mid, x)


def _regular(name, method, dims, starts, steps, dtype, n_full, **kw) -> Workload:
    dtype = np.dtype(dtype)
    starts = np.asarray(starts, dtype=dtype)
    steps = np.asarray(steps, dtype=dtype)
    grids = [(starts[d] + steps[d] * np.arange(dims[d], dtype=dtype)).astype(dtype) for d in range(len(dims))]
    lo = np.array([g[0] for g in grids], dtype=np.float64)
    hi = np.array([g[-1] for g in grids], dtype=np.float64)
    kw.setdefault("lo", lo)
    kw.setdefault("hi", hi)
    return Workload(name, method, False, dtype, list(dims), n_full, starts=starts, steps=steps, grids=grids, **kw)


def _jittered_axes(ndims: int, size: int, dtype, span=100.0, seed0=4) -> list[np.ndarray]:
    h = span / (size - 1)
    axes = []
    for d in range(ndims):
        x = np.linspace(0.0, span, size) + (uniform_np(seed0 + d, 0, size) - 0.5) * 0.8 * h
        x = x.astype(dtype)
        assert np.all(np.diff(x) > 0)
        axes.append(np.ascontiguousarray(x))
    return axes


def _rectilinear(name, method, ndims, size, dtype, n_full, **kw) -> Workload:
    dtype = np.dtype(dtype)
    grids = _jittered_axes(ndims, size, dtype)
    lo = np.array([g[0] for g in grids], dtype=np.float64)
    hi = np.array([g[-1] for g in grids], dtype=np.float64)
    kw.setdefault("lo", lo)
    kw.setdefault("hi", hi)
    return Workload(name, method, True, dtype, [size] * ndims, n_full, grids=grids, **kw)


def get(name: str, dtype=np.float64) -> Workload:
    """Build a named configuration (see module docstring)."""
    if name == "c1_linear3d_reg20":
        x = np.linspace(-1.0, 1.0, 20)
        return _regular(name, "linear", [20] * 3, [x[0]] * 3, [x[1] - x[0]] * 3, dtype, 1_000_000,
                        vals_lo=-1.0, vals_hi=1.0, lo=np.full(3, -0.99), hi=np.full(3, 0.99))  # fmt: skip
    if name == "c2_cubic3d_reg100":
        return _regular(name, "cubic", [100] * 3, [0.0] * 3, [100.0 / 99.0] * 3, dtype, 100_000_000,
                        linearize=True, oob_fraction=0.10)  # fmt: skip
    if name == "c3_linear4d_rect64":
        return _rectilinear(name, "linear", 4, 64, dtype, 100_000_000)
    if name == "c3_cubic4d_rect64":
        return _rectilinear(name, "cubic", 4, 64, dtype, 100_000_000, linearize=True)
    if name == "c4_linear6d_reg24":
        return _regular(name, "linear", [24] * 6, [0.0] * 6, [1.0] * 6, dtype, 1_000_000_000)
    if name == "c5_nearest2d_reg1024":
        return _regular(name, "nearest", [1024] * 2, [0.0] * 2, [0.125] * 2, dtype, 1_000_000_000,
                        oob_fraction=0.05, plant_special=True)  # fmt: skip
    if name == "c5_nearest3d_reg128":
        return _regular(name, "nearest", [128] * 3, [0.0] * 3, [0.5] * 3, dtype, 1_000_000_000,
                        oob_fraction=0.05, plant_special=True)  # fmt: skip
    if name == "c5_nearest2d_rect1024":
        return _rectilinear(name, "nearest", 2, 1024, dtype, 1_000_000_000, oob_fraction=0.05, plant_special=True)
    if name == "c5_nearest3d_rect128":
        return _rectilinear(name, "nearest", 3, 128, dtype, 1_000_000_000, oob_fraction=0.05, plant_special=True)
    # Extra shapes on L2-resident grids (not BASELINE.json configurations): the same kernels on the other
    # 3-D / 4-D method x grid-kind combinations, for parity at scale and for the per-kernel measurements.
    if name == "x_linear3d_reg100":
        return _regular(name, "linear", [100] * 3, [0.0] * 3, [100.0 / 99.0] * 3, dtype, 100_000_000, oob_fraction=0.10)
    if name == "x_linear4d_reg32":
        return _regular(name, "linear", [32] * 4, [0.0] * 4, [1.0] * 4, dtype, 100_000_000, oob_fraction=0.10)
    if name == "x_cubic4d_reg32":
        return _regular(name, "cubic", [32] * 4, [0.0] * 4, [1.0] * 4, dtype, 100_000_000,
                        linearize=True, oob_fraction=0.10)  # fmt: skip
    if name == "x_cubic2d_reg1024":
        return _regular(name, "cubic", [1024] * 2, [0.0] * 2, [0.125] * 2, dtype, 100_000_000,
                        linearize=True, oob_fraction=0.10)  # fmt: skip
    if name == "x_cubic2d_rect1024":
        return _rectilinear(name, "cubic", 2, 1024, dtype, 100_000_000, linearize=True, oob_fraction=0.10)
    if name == "x_cubic3d_rect100":
        return _rectilinear(name, "cubic", 3, 100, dtype, 100_000_000, linearize=True, oob_fraction=0.10)
    if name == "x_cubic4d_rect32":
        return _rectilinear(name, "cubic", 4, 32, dtype, 100_000_000, linearize=False, oob_fraction=0.10)
    if name == "x_linear4d_rect32":
        return _rectilinear(name, "linear", 4, 32, dtype, 100_000_000, oob_fraction=0.10)
    # Grids a little beyond L2 (134 / 134 / 95 MB in f64): the slab passes of launch_linear.cu on regular grids
    if name == "x_linear3d_reg256":
        return _regular(name, "linear", [256] * 3, [0.0] * 3, [1.0] * 3, dtype, 100_000_000, oob_fraction=0.10)
    if name == "x_linear4d_reg64":
        return _regular(name, "linear", [64] * 4, [0.0] * 4, [1.0] * 4, dtype, 100_000_000, oob_fraction=0.10)
    if name == "x_linear5d_reg26":
        return _regular(name, "linear", [26] * 5, [0.0] * 5, [1.0] * 5, dtype, 100_000_000)
    raise KeyError(name)


EXTRA = ["x_linear3d_reg100", "x_linear4d_reg32", "x_cubic4d_reg32", "x_cubic3d_rect100", "x_cubic4d_rect32", "x_linear4d_rect32",
         "x_cubic2d_reg1024", "x_cubic2d_rect1024",
         "x_linear3d_reg256", "x_linear4d_reg64", "x_linear5d_reg26"]

ALL = [
    "c1_linear3d_reg20",
    "c2_cubic3d_reg100",
    "c3_linear4d_rect64",
    "c3_cubic4d_rect64",
    "c4_linear6d_reg24",
    "c5_nearest2d_reg1024",
    "c5_nearest3d_reg128",
    "c5_nearest2d_rect1024",
    "c5_nearest3d_rect128",
]


def algorithmic_bytes(w: Workload, n: int) -> int:
    """Roofline numerator: query bytes in + output bytes out + every grid/axis byte
    once (capped by what n points can touch)."""
    e = w.dtype.itemsize
    fp = 4 if w.method == "cubic" else (1 if w.method == "nearest" else 2)
    grid = min(w.nvals, n * fp ** w.ndims) * e
    axes = sum(len(g) for g in w.grids) * e if w.rect else 0
    return n * (w.ndims + 1) * e + grid + axes
