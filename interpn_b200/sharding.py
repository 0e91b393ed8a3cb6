"""Multi-GPU plumbing: one process per GPU, query batch sharded, grid replicated once.

Every query point is independent (`interp_one` is pure, /root/reference/src/multilinear/regular.rs:296),
so the path shards by contiguous index ranges of the query batch and needs no collective while
evaluating. The only exchange is at interpolator creation: rank 0 owns the grid description and
the values; every other rank allocates uninitialised resident storage
(`INTERPN_B200_VALS_UNINIT`) and one broadcast (NCCL over NVLink on the GPU box) fills it in
place. SURVEY.md §8(e).

Nothing here computes: the functions take the process group and a factory for the resident
interpolator, so the same code runs under `gloo` on CPU in tests/test_sharding_gloo.py (with a
stand-in factory) and under `nccl` in bench.py.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable

import numpy as np


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard `[rank*n/world, (rank+1)*n/world)` of `n` query points (sizes differ by at
    most one point; the union over ranks is exactly `[0, n)`)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("need 0 <= rank < world and n >= 0")
    return (rank * n) // world, ((rank + 1) * n) // world


@dataclass
class GridSpec:
    """What a rank needs to allocate its replica: everything except the values."""

    method: str                          # "linear" | "cubic" | "nearest"
    rect: bool
    dtype: str                           # "float64" | "float32"
    linearize_extrapolation: bool = True
    dims: list[int] = field(default_factory=list)
    starts: list[float] | None = None    # regular
    steps: list[float] | None = None
    grids: list[list[float]] | None = None   # rectilinear axes

    @property
    def nvals(self) -> int:
        return int(np.prod(self.dims, dtype=np.int64)) if self.dims else 0


def broadcast_spec(spec: GridSpec | None, src: int = 0, group=None) -> GridSpec:
    """Rank `src` passes its GridSpec; every rank returns the same spec."""
    import torch.distributed as dist

    box = [spec if dist.get_rank(group) == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    if box[0] is None:
        raise RuntimeError("grid spec broadcast delivered nothing")
    return box[0]


def replicate(spec: GridSpec | None, vals, make_resident: Callable[[GridSpec, Any], Any], src: int = 0, group=None):
    """Create this rank's resident interpolator and replicate the grid values into it.

    `make_resident(spec, vals_or_None)` returns an object with `vals_tensor()` (a tensor aliasing
    its resident storage) and `vals_updated(stream=...)`; rank `src` passes its values, the others
    pass None. Returns (resident, spec). One broadcast_object_list (a few hundred bytes) and one
    tensor broadcast (the values) are the only collectives this package ever issues.
    """
    import torch.distributed as dist

    rank = dist.get_rank(group)
    spec = broadcast_spec(spec if rank == src else None, src, group)
    resident = make_resident(spec, vals if rank == src else None)
    buf = resident.vals_tensor()
    if buf.numel() != spec.nvals:
        raise AssertionError("Dimension mismatch")
    dist.broadcast(buf, src=src, group=group)
    if rank != src:
        resident.vals_updated()
    return resident, spec


def make_interpolator(spec: GridSpec, vals):
    """The real factory: a grid resident on this process's CUDA device."""
    from .interpolator import Interpolator

    dt = np.dtype(spec.dtype)
    if spec.rect:
        grids = [np.asarray(g, dtype=dt) for g in spec.grids]
        return Interpolator.rectilinear(spec.method, grids, vals, spec.linearize_extrapolation, dtype=dt)
    return Interpolator.regular(spec.method, spec.dims, np.asarray(spec.starts, dtype=dt), np.asarray(spec.steps, dtype=dt),
                                vals, spec.linearize_extrapolation, dtype=dt)  # fmt: skip


def max_over_ranks(seconds: float, device=None, group=None) -> float:
    """Job time of a sharded evaluation = the slowest rank's device time."""
    import torch
    import torch.distributed as dist

    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
