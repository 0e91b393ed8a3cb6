"""Host-side multi-rank logic (interpn_b200/sharding.py) under `gloo`, world_size 2, on CPU.

The GPU path shards the query batch and replicates the grid with one broadcast.
Here the resident interpolator is replaced by a stand-in that only holds storage, so what is
tested is the plumbing: shard ranges, the spec + values broadcast into uninitialised storage, the
`vals_updated` notification on non-source ranks and the max-over-ranks time reduction.
"""

import os
import socket

import numpy as np
import pytest

from interpn_b200.sharding import GridSpec, shard_range


@pytest.mark.parametrize("n", [0, 1, 7, 1000, 10**9 + 7])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shard_ranges_partition_the_batch(n, world):
    edges = [shard_range(n, r, world) for r in range(world)]
    assert edges[0][0] == 0 and edges[-1][1] == n
    for (_, hi), (lo, _) in zip(edges[:-1], edges[1:]):
        assert hi == lo
    sizes = [hi - lo for lo, hi in edges]
    assert max(sizes) - min(sizes) <= 1


def test_shard_range_rejects_bad_ranks():
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)
    with pytest.raises(ValueError):
        shard_range(10, 0, 0)


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _Standin:
    """Holds resident storage like interpn_b200.Interpolator, computes nothing."""

    def __init__(self, spec, vals):
        import torch

        self.spec = spec
        self.updated = 0
        self.store = torch.full((spec.nvals,), float("nan"), dtype=torch.float64)
        if vals is not None:
            self.store.copy_(torch.from_numpy(np.asarray(vals, dtype=np.float64)))

    def vals_tensor(self):
        return self.store

    def vals_updated(self, stream=0):
        self.updated += 1


def _worker(rank: int, world: int, port: int, q):
    import torch.distributed as dist

    from interpn_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        spec = vals = None
        if rank == 0:
            spec = GridSpec("cubic", False, "float64", True, dims=[5, 4, 6], starts=[0.0, 1.0, 2.0], steps=[0.5, 0.25, 1.0])
            vals = np.arange(spec.nvals, dtype=np.float64) * 0.5 - 7.0
        resident, got = sharding.replicate(spec, vals, _Standin, src=0)
        lo, hi = shard_range(1001, rank, world)
        job = sharding.max_over_ranks(0.25 * (rank + 1))
        q.put((rank, got, resident.store.numpy().copy(), resident.updated, (lo, hi), job))
    finally:
        dist.destroy_process_group()


def test_grid_replication_and_timing_reduction_world2():
    import torch.multiprocessing as mp

    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted((q.get(timeout=120) for _ in range(world)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.arange(5 * 4 * 6, dtype=np.float64) * 0.5 - 7.0
    for rank, spec, store, updated, (lo, hi), job in results:
        assert spec.dims == [5, 4, 6] and spec.method == "cubic" and spec.steps == [0.5, 0.25, 1.0]
        assert np.array_equal(store, want)           # every replica holds rank 0's values
        assert updated == (0 if rank == 0 else 1)    # derived layouts rebuilt only where the broadcast wrote
        assert job == 0.5                            # slowest rank
    assert results[0][4] == (0, 500) and results[1][4] == (500, 1001)
