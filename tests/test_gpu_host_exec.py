"""The host-buffer executor (interpn_b200/csrc/host_exec.cuh): pageable vs pinned caller memory, many chunks, the
reference's failure semantics across chunks, and several GPUs behind ONE C call (skipped on a one-GPU box; run with
`gpurun --gpus 2`). Needs a B200: `pytest -m gpu`."""

import numpy as np
import pytest

from tests.test_gpu_parity import assert_same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ib():
    import interpn_b200

    return interpn_b200


def _case(n, seed=5):
    rng = np.random.default_rng(seed)
    dims = [40, 30, 20]
    starts = np.array([-1.0, 2.0, 0.0])
    steps = np.array([0.25, 0.5, 1.0])
    vals = rng.standard_normal(int(np.prod(dims)))
    obs = [rng.random(n) * 11.0 - 1.5, rng.random(n) * 16.0 + 1.5, rng.random(n) * 21.0 - 1.0]
    return dims, starts, steps, vals, obs


@pytest.mark.parametrize("method", ["linear", "cubic", "nearest"])
def test_pageable_and_pinned_callers_get_the_same_bits(ib, oracle, method):
    torch = pytest.importorskip("torch")
    n = 6_300_007  # several 8 MiB staging chunks, ragged tail
    dims, starts, steps, vals, obs = _case(n)
    want = oracle.interpn_regular(method, dims, starts, steps, vals, obs, linearize_extrapolation=True, nthreads=8)
    with ib.Interpolator.regular(method, dims, starts, steps, vals, True) as it:
        got = it.eval(obs)  # numpy arrays: pageable -> pinned staging + copy threads
        assert_same_bits(got, want, "pageable")
        pobs = [torch.from_numpy(o).pin_memory() for o in obs]
        pout = torch.empty(n, dtype=torch.float64).pin_memory()
        it.eval([p.numpy() for p in pobs], pout.numpy())  # DMA in place
        assert_same_bits(pout.numpy(), want, "pinned")
        mixed_out = np.full(n, -1.0)
        it.eval([pobs[0].numpy(), obs[1], pobs[2].numpy()], mixed_out)  # one pageable array is enough to stage the inputs
        assert_same_bits(mixed_out, want, "mixed")
    assert ib.copy_threads() >= 1


@pytest.mark.parametrize("pinned", [False, True])
def test_failure_semantics_across_chunks(ib, oracle, pinned):
    """Earlier outputs written, later untouched (multilinear/regular.rs:276-280) when the failing point sits in a middle
    chunk and later chunks were already in flight; the smallest index wins."""
    torch = pytest.importorskip("torch")
    n = 9_000_001
    dims, starts, steps, vals, obs = _case(n, seed=6)
    bad = 5_123_457
    obs[1][bad] = np.nan
    obs[0][bad + 2_000_000] = np.inf
    out = np.full(n, -7.0)
    if pinned:
        keep = [torch.from_numpy(o).pin_memory() for o in obs]
        obs = [k.numpy() for k in keep]
        pout = torch.full((n,), -7.0, dtype=torch.float64).pin_memory()
        out = pout.numpy()
    with ib.Interpolator.regular("linear", dims, starts, steps, vals) as it:
        with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
            it.eval(obs, out)
        assert it.first_bad == bad
    want = oracle.interpn_regular("linear", dims, starts, steps, vals, [o[:bad] for o in obs], nthreads=8)
    assert_same_bits(np.ascontiguousarray(out[:bad]), want)
    assert np.all(out[bad:] == -7.0)


def test_several_gpus_behind_one_call(ib, oracle):
    if ib.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    n = 40_000_003
    dims, starts, steps, vals, obs = _case(n, seed=7)
    try:
        ib.set_host_devices(1)
        with ib.Interpolator.regular("cubic", dims, starts, steps, vals, True) as it:
            one = it.eval(obs)
        ib.set_host_devices(0)
        assert ib.host_devices() == ib.device_count()
        with ib.Interpolator.regular("cubic", dims, starts, steps, vals, True) as it:
            many = it.eval(obs)
            again = it.eval(obs)  # replicas are reused
            assert_same_bits(many, one)
            assert_same_bits(again, one)
            # the failure rule is global over devices
            bad = n // 2 + 11
            obs2 = [o.copy() for o in obs]
            obs2[2][bad] = np.nan
            obs2[0][n - 5] = np.nan
            out = np.full(n, -7.0)
            with pytest.raises(AssertionError, match="Unrepresentable coordinate value"):
                it.eval(obs2, out)
            assert it.first_bad == bad
            assert_same_bits(np.ascontiguousarray(out[:bad]), np.ascontiguousarray(one[:bad]))
            assert np.all(out[bad:] == -7.0)
        sl = slice(0, n, 811)
        want = oracle.interpn_regular("cubic", dims, starts, steps, vals, [np.ascontiguousarray(o[sl]) for o in obs], nthreads=8)
        assert_same_bits(np.ascontiguousarray(many[sl]), want)
        # one-shot C entry points take the same route
        out = np.zeros(n)
        ib.raw.interpn_linear_regular_f64(dims, starts, steps, vals, obs, out)
        ib.set_host_devices(1)
        out1 = np.zeros(n)
        ib.raw.interpn_linear_regular_f64(dims, starts, steps, vals, obs, out1)
        assert_same_bits(out, out1)
    finally:
        ib.set_host_devices(0)
