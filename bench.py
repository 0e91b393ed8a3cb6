#!/usr/bin/env python
"""bench.py — headline benchmark of the InterpN hot path on B200.

A "step" is one pass of the hot path over one batch of synthetic query points:
BASELINE.json config[1] by default — 3-D multicubic, regular 100^3 f64 grid,
linearize_extrapolation=true, 1e8 query points per GPU of which 10 % lie outside the grid
(workload `c2_cubic3d_reg100`, interpn_b200/workloads.py).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--workload NAME] [--points P]

* `value`  : query points/s, whole job, inputs already resident in HBM, CUDA-event timed,
             max over ranks. Weak scaling: every rank evaluates `points` queries on its own replica
             of the grid (replicated once by an NCCL broadcast before the timed region).
* `e2e`    : the same metric through the public host-buffer API (`Interpolator.eval` ->
             interpn_b200_interp_eval_host_f64), pinned host arrays in, pinned host array out,
             H2D/D2H copies inside the timed region.
* `roofline`: algorithmic bytes (query bytes in + outputs out + every grid byte once) / kernel time,
             against the measured HBM copy bandwidth in MEASURED_PEAKS.json.
* `cpu_baseline`: the CPU oracle (a port of interpn 0.8.2's arithmetic; the Rust crate cannot be
             built in this image) timed on this box's host cores on a bounded sample.
* `--impl reference`: times that CPU implementation alone (all host threads) and prints the
             reference-arm line. Rank 0 only.

Only the cpu_baseline / --impl reference legs touch oracle/; the GPU arm never does (apart from
the bit-parity spot check that follows the timed region).
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2_cubic3d_reg100")
    ap.add_argument("--points", type=int, default=0, help="query points per GPU (default: the workload's full size, capped at 1e8)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target duration of the CPU baseline sample")
    ap.add_argument("--arithmetic", default=os.environ.get("INTERPN_B200_ARITHMETIC", "strict"), choices=["strict", "fma"],
                    help="reference build whose arithmetic is reproduced: crate default features (strict, the headline) "
                         "or the crate's `fma` feature = the Python wheel's build (libinterpn_b200_fma.so)")  # fmt: skip
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="element type (the headline metric is f64; C5 names both)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


METRIC = "query points/sec (f64) 3D/4D linear+cubic; % of HBM roofline at 1/2/4/8 GPU"
UNIT = "points/s"


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str):
    """DRAM bytes of the dominant kernel per launch (per step, summed over the launches, where a step is several) from the
    committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(p) as f:
            v = json.load(f).get(workload)
            return int(v) if isinstance(v, (int, float)) else None
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle) — also the whole of the `--impl reference` arm
# ------------------------------------------------------------------------------------------------


def cpu_eval(oracle, w, vals, obs, nthreads):
    if w.rect:
        return oracle.interpn_rectilinear(w.method, w.grids, vals, obs, linearize_extrapolation=w.linearize, nthreads=nthreads)
    return oracle.interpn_regular(w.method, w.dims, w.starts, w.steps, vals, obs, linearize_extrapolation=w.linearize, nthreads=nthreads)


def cpu_baseline(w, target_seconds: float, nthreads: int | None = None, steps: int = 1, warmup: int = 0):
    """Time the CPU oracle on a bounded sample of workload `w`; returns (points/s, dict)."""
    from oracle import oracle

    oracle.build()
    cores = nthreads or max(1, oracle.max_threads())
    vals = w.vals("np")
    probe_n = 200_000
    obs = w.queries(0, probe_n, "np")
    t0 = time.perf_counter()
    cpu_eval(oracle, w, vals, obs, cores)
    rate = probe_n / max(time.perf_counter() - t0, 1e-6)
    n = int(min(w.n_full, max(probe_n, rate * target_seconds)))
    n = min(n, 50_000_000)
    obs = w.queries(0, n, "np")
    for _ in range(warmup):
        cpu_eval(oracle, w, vals, obs, cores)
    times = []
    for _ in range(max(1, steps)):
        t0 = time.perf_counter()
        cpu_eval(oracle, w, vals, obs, cores)
        times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    info = {
        "value": n / dt,
        "unit": UNIT,
        "cores": cores,
        "kind": "port",
        "sample": f"first {n} query points of {w.name} (same grid, same generator), {len(times)} pass(es), "
                  f"oracle port of interpn 0.8.2 ({'fma' if oracle.DEFAULT_FMA else 'strict'} arithmetic, -O3 -march=x86-64-v3), std::thread x{cores}",
        "ms_per_pass": dt * 1e3,
    }
    return n / dt, info, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from interpn_b200 import workloads as W

    w = W.get(args.workload)
    # bounded so that steps+warmup passes end within a few minutes
    per_pass = max(2.0, min(args.cpu_seconds, 150.0 / max(1, args.steps + args.warmup)))
    value, info, dt = cpu_baseline(w, per_pass, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": dt * 1e3,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": w.name, "points_per_step": int(round(value * dt)), "grid": w.dims, "method": w.method,
                   "note": "CPU path; a step is one pass over the bounded sample described in cpu_baseline.sample"},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# Clock sampling during the timed region (B200_PROFILING.md "clocks line")
# ------------------------------------------------------------------------------------------------


class ClockSampler:
    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------


def run_b200(args):
    import torch
    import torch.distributed as dist

    import interpn_b200 as ib
    from interpn_b200 import workloads as W

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: interpn_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ib.set_device(local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    w = W.get(args.workload, np.float32 if args.dtype == "f32" else np.float64)
    n = args.points or min(w.n_full, 100_000_000)
    tdtype = torch.float32 if args.dtype == "f32" else torch.float64

    # ---- grid: built on rank 0, replicated once by an NCCL broadcast straight into each rank's
    # resident storage (the only collective; none on the evaluation path).
    t_setup = time.perf_counter()
    from interpn_b200 import sharding

    spec = sharding.GridSpec(
        w.method, w.rect, "float32" if args.dtype == "f32" else "float64", bool(w.linearize), dims=list(w.dims),
        starts=None if w.rect else [float(v) for v in w.starts], steps=None if w.rect else [float(v) for v in w.steps],
        grids=[[float(v) for v in g] for g in w.grids] if w.rect else None,
    )  # fmt: skip
    bcast_ms = None
    if distributed:
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        interp, _ = sharding.replicate(spec if rank == 0 else None, w.vals("torch", dev) if rank == 0 else None,
                                       sharding.make_interpolator, src=0)  # fmt: skip
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    else:
        interp = sharding.make_interpolator(spec, w.vals("torch", dev))

    # ---- this rank's shard of the query batch, generated on the device
    base = rank * n
    obs = [torch.empty(n, dtype=tdtype, device=dev) for _ in range(w.ndims)]
    blk = 1 << 24
    for lo in range(0, n, blk):
        cnt = min(blk, n - lo)
        q = w.queries(base + lo, cnt, "torch", dev)
        for d in range(w.ndims):
            obs[d][lo : lo + cnt] = q[d]
        del q
    out = torch.empty(n, dtype=tdtype, device=dev)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup

    stream = torch.cuda.current_stream(dev)

    def step():
        interp.eval_torch(obs, out)

    # ncu --profile-from-start off sees only warm-up + timed region (input generation is torch's)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(max(3, args.warmup)):
        step()
    interp.status(stream.cuda_stream)

    # ---- timed region: K steps, CUDA events on the launching stream, barrier + sync both sides
    torch.cuda.synchronize()
    if distributed:
        dist.barrier()
        torch.cuda.synchronize()
    launches0 = ib.launch_count()
    swept0 = ib.swept_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    with ClockSampler(local_rank) as clocks:
        ev[0].record(stream)
        for k in range(args.steps):
            step()
            ev[k + 1].record(stream)
        torch.cuda.synchronize()
    if distributed:
        dist.barrier()
    launches = ib.launch_count() - launches0
    swept = ib.swept_launch_count() - swept0
    torch.cuda.cudart().cudaProfilerStop()
    interp.status(stream.cuda_stream)
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_total_ms = float(t.item())
    value = world * n * args.steps / (max_total_ms * 1e-3)

    # ---- parity spot check against the oracle (outside the timed region)
    parity = None
    if rank == 0:
        try:
            from oracle import oracle

            oracle.build()
            sl = slice(0, n, max(1, n // 50_000))
            o = [x[sl].contiguous().cpu().numpy() for x in obs]
            vals_h = interp.vals_tensor().cpu().numpy()
            want = cpu_eval(oracle, w, vals_h, o, max(1, oracle.max_threads()))
            got = out[sl].contiguous().cpu().numpy()
            parity = {"sample_points": int(got.size), "bit_identical": bool(np.array_equal(got.view(np.uint32 if got.dtype == np.float32 else np.uint64), want.view(np.uint32 if want.dtype == np.float32 else np.uint64))),
                      "max_abs_diff": float(np.max(np.abs(got - want)))}
        except Exception as e:  # the oracle is a checker, never a dependency of the measured path
            parity = {"error": repr(e)}

    # ---- end-to-end through the host-buffer API (pinned host memory)
    e2e = None
    if not args.no_e2e:
        hobs = [torch.empty(n, dtype=tdtype).pin_memory() for _ in range(w.ndims)]
        hout = torch.empty(n, dtype=tdtype).pin_memory()
        for d in range(w.ndims):
            hobs[d].copy_(obs[d])
        torch.cuda.synchronize()
        nobs = [h.numpy() for h in hobs]
        nout = hout.numpy()
        interp.eval(nobs, nout)  # warm-up: sizes the copy pipeline
        if distributed:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            interp.eval(nobs, nout)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        same = bool(torch.equal(hout.to(dev), out))
        e2e = {
            "value": world * n * args.e2e_steps / dt,
            "unit": UNIT,
            "h2d_bytes_per_step": int(n * w.ndims * 8),
            "d2h_bytes_per_step": int(n * 8),
            "ms_per_step": dt / args.e2e_steps * 1e3,
            "steps": args.e2e_steps,
            "api": "interpn_b200.Interpolator.eval -> interpn_b200_interp_eval_host_f64 (pinned host buffers, 3-slot copy/compute pipeline)",
            "matches_device_path": same,
        }
        del hobs, hout

    if rank != 0:
        if distributed:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_peak_hbm()
    kernel_ms = float(np.mean(step_ms))
    abytes = W.algorithmic_bytes(w, n)
    achieved = abytes / (kernel_ms * 1e-3) / 1e9
    roofline = {
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak,
        "traffic": ncu_traffic(w.name),
        "peak_source": peak_src,
        "kernel": "dominant evaluation kernel of the step (one launch per step)" if launches <= args.steps
        else f"all {launches / max(args.steps, 1):g} launches of a step (sort + evaluation of the bin-swept path, or the slab passes): algorithmic bytes per step over the step's device time",
        "algorithmic_bytes_per_launch": abytes,
        "kernel_ms": kernel_ms,
    }
    line = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": max_total_ms / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": args.dtype,
        "data": "synthetic",
        "config": {
            "workload": w.name,
            "method": w.method,
            "grid": w.dims,
            "grid_kind": "rectilinear" if w.rect else "regular",
            "points_per_gpu": n,
            "out_of_bounds_fraction": w.oob_fraction,
            "linearize_extrapolation": bool(w.linearize),
            "arithmetic": args.arithmetic,
            "l2": f"query arrays ({n * (w.ndims + 1) * w.dtype.itemsize / 1e9:.2f} GB per step) exceed L2; the {w.nvals * w.dtype.itemsize / 1e6:.0f} MB grid is reused across steps by design",
            "parallelism": f"query batch sharded over {world} GPU(s), grid replicated by one NCCL broadcast" if distributed else "single GPU",
        },
        "roofline": roofline,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "swept_launches": int(swept),
        "clocks": clocks.summary(),
        "parity": parity,
        "setup_s": setup_s,
        "grid_broadcast_ms": bcast_ms,
        "step_ms": step_ms,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            _, info, _ = cpu_baseline(w, args.cpu_seconds)
            line["cpu_baseline"] = {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")}
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    # read by interpn_b200/_lib.py (which library is loaded) and by oracle/oracle.py (which arithmetic the
    # parity spot check and the CPU legs run in) when they are first imported
    os.environ["INTERPN_B200_ARITHMETIC"] = args.arithmetic
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
