#!/usr/bin/env python
"""bench.py — benchmark of the InterpN hot path on B200.

A "step" is one pass of the hot path over one batch of synthetic query points. The headline is BASELINE.json
config[1] — 3-D multicubic, regular 100^3 f64 grid, linearize_extrapolation=true, 1e8 query points per GPU of which
10 % lie outside the grid (workload `c2_cubic3d_reg100`, interpn_b200/workloads.py).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--workload NAME] [--points P] [--suite auto|all|none] ...

One JSON line (rank 0):
* `value`   : query points/s, whole job, inputs already resident in HBM, CUDA events on the launching stream, max over
              ranks. Weak scaling: every rank evaluates `points` queries on its own replica of the grid (replicated once
              by an NCCL broadcast before the timed region; no collective on the evaluation path).
* `sustained`: the same over a longer back-to-back run (default 200 steps): clocks settle below the burst figure.
* `roofline`: algorithmic bytes (query bytes in + outputs out + every grid byte once) / mean step time,
              against the measured HBM copy bandwidth in MEASURED_PEAKS.json; `traffic` = DRAM bytes per step of the
              committed ncu capture of the same workload (profiles/ncu_traffic.json names the capture).
* `e2e`     : the same metric through the public host-buffer API (`Interpolator.eval` ->
              interpn_b200_interp_eval_host_f64): PINNED host arrays in and out, copies inside the timed region;
              `e2e.pageable` repeats it with ordinary (pageable) numpy arrays — what a drop-in caller passes; at N > 1
              `e2e.single_call` is ONE C call from rank 0 that shards its batch over all N GPUs in-process.
* `cpu_baseline`: the CPU oracle (a port of interpn 0.8.2's arithmetic; the Rust crate cannot be built in this image)
              timed on this box's host cores on a bounded sample, all cores and single-threaded.
* `workloads`: the other BASELINE.json configurations measured the same way in the same run (C1 through a CUDA graph of
              100 launches, C3 linear + cubic, C4 at 1.25e8 points/GPU, C5 2-D/3-D x regular/rectilinear x f32/f64 at
              1e9 points), each with its own roofline, parity and clocks. `--suite none` skips them.
* `--impl reference`: times that CPU implementation alone (all host threads) and prints the reference-arm line. Rank 0
              only; never loads the CUDA library.

Only the cpu_baseline / --impl reference legs and the parity check that FOLLOWS each timed region touch oracle/.
"""

from __future__ import annotations

import argparse
import importlib.util
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "query points/sec (f64) 3D/4D linear+cubic; % of HBM roofline at 1/2/4/8 GPU"
UNIT = "points/s"
HEADLINE = "c2_cubic3d_reg100"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=HEADLINE)
    ap.add_argument("--points", type=int, default=0, help="query points per GPU (default: the workload's full size, capped at 1e8; C4 1.25e8)")
    ap.add_argument("--suite", default="auto", choices=["auto", "all", "none"],
                    help="also measure the other BASELINE configurations (auto: all at N=1 on the default workload, C4 at N>1)")  # fmt: skip
    ap.add_argument("--sustained-steps", type=int, default=200, help="length of the back-to-back sustained run (0 = skip)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="target duration of the all-core CPU baseline sample")
    ap.add_argument("--arithmetic", default=os.environ.get("INTERPN_B200_ARITHMETIC", "strict"), choices=["strict", "fma"],
                    help="reference build whose arithmetic is reproduced: crate default features (strict, the headline) "
                         "or the crate's `fma` feature = the Python wheel's build (libinterpn_b200_fma.so)")  # fmt: skip
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"], help="element type (the headline metric is f64; C5 names both)")
    ap.add_argument("--graph-launches", type=int, default=0, help="time the step as a CUDA graph of this many launches (C1: 100)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def load_workloads():
    """interpn_b200/workloads.py by file path: numpy only. Importing it as `interpn_b200.workloads` would run the package's
    __init__ and load the CUDA library, which the reference arm must never do."""
    name = "_interpn_b200_workloads"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "interpn_b200", "workloads.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy bandwidth)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str, dtype: str, n: int):
    """DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum summed over the step's launches) from the committed
    ncu capture of this workload, scaled to this run's point count when the capture used another; (bytes, source)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            table = json.load(f)
        v = table.get(f"{workload}:{dtype}") or table.get(workload)
        if isinstance(v, (int, float)):
            return (int(v), None) if dtype == "f64" else (None, None)
        if isinstance(v, dict):
            scale = n / v["points"] if v.get("points") else 1.0
            return int(v["bytes"] * scale), v.get("source")
    except Exception:
        pass
    return None, None


# The unit that actually binds each workload's dominant kernel, as (microbenchmark key in
# profiles/r2_microbench_b200.json, units per query point, description). Reported next to the HBM roofline as
# roofline.binding = {unit, per_point, measured_peak, frac}: how close the kernel is to the ceiling of what it has to move.
BINDING = {
    "c2_cubic3d_reg100": ("quad128B_ldg256_L2_Gsectors_s", 16.75, "L2 -> L1 32-byte sectors (16 coefficient sectors + coordinates per point)"),
    "c3_cubic4d_rect64": ("quad128B_ldg256_L2_Gsectors_s", 65.0, "L2 -> L1 32-byte sectors (64 coefficient sectors + coordinates per point)"),
    "c3_linear4d_rect64": ("quad128B_aligned_hbm_Glines_s", 1.0, "random aligned 128-byte lines from HBM (one hypercube block per point)"),
    "c4_linear6d_reg24": ("quad128B_aligned_hbm_Glines_s", 4.0, "random aligned 128-byte lines from HBM (four hypercube blocks per point)"),
    "c5_nearest2d_reg1024": ("gather8B_random_Gloads_s", 1.0, "L1 wavefronts of lone gathers (one node per point)"),
    "c5_nearest3d_reg128": ("gather8B_random_Gloads_s", 1.0, "L1 wavefronts of lone gathers (one node per point)"),
    "c5_nearest2d_rect1024": ("gather8B_random_Gloads_s", 1.0, "L1 wavefronts of lone gathers (one node per point; the axis search adds shared-memory wavefronts)"),
    "c5_nearest3d_rect128": ("gather8B_random_Gloads_s", 1.0, "L1 wavefronts of lone gathers (one node per point; the axis search adds shared-memory wavefronts)"),
}


def binding_ceiling(workload: str, points_per_s_per_gpu: float):
    try:
        key, per_point, what = BINDING[workload]
        with open(os.path.join(ROOT, "profiles", "r2_microbench_b200.json")) as f:
            peak = float(json.load(f)[key]) * 1e9
        return {"unit": what, "per_point": per_point, "measured_peak_per_s": peak, "source": f"profiles/r2_microbench_b200.json {key}",
                "frac": points_per_s_per_gpu * per_point / peak}
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# CPU baseline (oracle) — also the whole of the `--impl reference` arm
# ------------------------------------------------------------------------------------------------


def cpu_eval(oracle, w, vals, obs, nthreads):
    if w.rect:
        return oracle.interpn_rectilinear(w.method, w.grids, vals, obs, linearize_extrapolation=w.linearize, nthreads=nthreads)
    return oracle.interpn_regular(w.method, w.dims, w.starts, w.steps, vals, obs, linearize_extrapolation=w.linearize, nthreads=nthreads)


def cpu_baseline(w, target_seconds: float, nthreads: int | None = None, steps: int = 1, warmup: int = 0, vals=None,
                 max_points: int = 50_000_000):
    """Time the CPU oracle on a bounded sample of workload `w`; returns (points/s, info dict, seconds per pass)."""
    from oracle import oracle

    oracle.build()
    cores = nthreads or max(1, oracle.max_threads())
    if vals is None:
        vals = w.vals("np")
    probe_n = 20_000 * cores
    obs = w.queries(0, probe_n, "np")
    t0 = time.perf_counter()
    cpu_eval(oracle, w, vals, obs, cores)
    rate = probe_n / max(time.perf_counter() - t0, 1e-6)
    n = int(min(w.n_full, max_points, max(probe_n, rate * target_seconds)))
    obs = w.queries(0, n, "np")
    for _ in range(warmup):
        cpu_eval(oracle, w, vals, obs, cores)
    if steps == 1 and n / rate < 0.5 * target_seconds:  # the sample is capped: repeat it to fill the time
        steps = int(min(50, max(1, target_seconds * rate / n)))
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
        "sample": f"first {n} query points of {w.name} ({w.dtype.name}; same grid, same generator), {len(times)} pass(es), "
                  f"oracle port of interpn 0.8.2 ({'fma' if oracle.DEFAULT_FMA else 'strict'} arithmetic, dimensionality a compile-time "
                  f"constant like the crate's const-generic structs, g++ -O3 -march=x86-64-v3 -ffp-contract=off), std::thread x{cores}",
        "ms_per_pass": dt * 1e3,
    }  # fmt: skip
    return n / dt, info, dt


def cpu_baseline_both(w, seconds: float, vals=None, max_points: int = 50_000_000):
    """All host cores (the headline CPU figure) and one thread — the reference's real execution model
    (multilinear/regular.rs:276-280: a serial loop, no threading anywhere in the crate)."""
    keys = ("value", "unit", "cores", "kind", "sample")
    if vals is None:
        vals = w.vals("np")
    _, info, _ = cpu_baseline(w, seconds, vals=vals, max_points=max_points)
    res = {k: info[k] for k in keys}
    _, one, _ = cpu_baseline(w, max(1.0, seconds / 2), nthreads=1, vals=vals, max_points=max_points)
    res["single_thread"] = {k: one[k] for k in keys}
    return res


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    W = load_workloads()
    w = W.get(args.workload, np.float32 if args.dtype == "f32" else np.float64)
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
        "dtype": args.dtype,
        "data": "synthetic",
        "config": {"workload": w.name, "points_per_step": int(round(value * dt)), "grid": w.dims, "method": w.method,
                   "note": "CPU path; a step is one pass over the bounded sample described in cpu_baseline.sample"},
        "cpu_baseline": {k: info[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "cuda_library_loaded": any(m == "interpn_b200" or m.startswith("interpn_b200.") for m in sys.modules),
    }  # fmt: skip
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
        return {"sm_mhz": float(np.median(self.samples)), "sm_min_mhz": float(np.min(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}  # fmt: skip


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------


class Ctx:
    """Process-wide state of the GPU arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import interpn_b200 as ib

        self.torch, self.dist, self.ib, self.args = torch, dist, ib, args
        self.W = load_workloads()
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py --impl b200 needs a CUDA device: interpn_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        ib.set_device(self.local_rank)
        self.distributed = self.world > 1
        if self.distributed:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
            ib.set_host_devices(1)  # one process per GPU: a rank's host-buffer calls stay on its own device
        self.stream = torch.cuda.current_stream(self.dev)
        self.peak, self.peak_src = measured_peak_hbm()
        self.last_vals = None

    def barrier(self):
        if self.distributed:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        if not self.distributed:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(self, ok: bool) -> bool:
        if not self.distributed:
            return ok
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int32, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())


def build_interp(cx: Ctx, w, dtype: str):
    """Grid built on rank 0 and replicated once by an NCCL broadcast straight into each rank's resident storage
    (the only collective; none on the evaluation path). Returns (interp, broadcast_ms)."""
    from interpn_b200 import sharding

    spec = sharding.GridSpec(
        w.method, w.rect, "float32" if dtype == "f32" else "float64", bool(w.linearize), dims=list(w.dims),
        starts=None if w.rect else [float(v) for v in w.starts], steps=None if w.rect else [float(v) for v in w.steps],
        grids=[[float(v) for v in g] for g in w.grids] if w.rect else None,
    )  # fmt: skip
    torch = cx.torch
    if not cx.distributed:
        return sharding.make_interpolator(spec, w.vals("torch", cx.dev)), None
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    interp, _ = sharding.replicate(spec if cx.rank == 0 else None, w.vals("torch", cx.dev) if cx.rank == 0 else None,
                                   sharding.make_interpolator, src=0)  # fmt: skip
    e1.record()
    torch.cuda.synchronize()
    return interp, e0.elapsed_time(e1)


def gen_queries(cx: Ctx, w, base: int, n: int, tdtype):
    """This rank's shard of the query batch, generated on the device (counter-based: any index range, any device)."""
    torch = cx.torch
    obs = [torch.empty(n, dtype=tdtype, device=cx.dev) for _ in range(w.ndims)]
    blk = 1 << 24
    for lo in range(0, n, blk):
        cnt = min(blk, n - lo)
        q = w.queries(base + lo, cnt, "torch", cx.dev)
        for d in range(w.ndims):
            obs[d][lo : lo + cnt] = q[d]
        del q
    return obs


def parity_check(cx: Ctx, w, interp, obs, out, n: int):
    """Bit comparison with the oracle on a strided sample of THIS rank's shard, after the timed region: 1e7 points for the
    1e9-point nearest runs, fewer where the oracle is slower; plus the wrapping 64-bit sum of the sample's
    output bits on both sides and of the whole output array on the device."""
    torch = cx.torch
    try:
        from oracle import oracle

        oracle.build()
        want_n = {"nearest": 10_000_000, "linear": 400_000 if w.ndims <= 4 else 100_000, "cubic": 100_000 if w.ndims <= 3 else 20_000}[w.method]
        stride = max(1, n // want_n)
        sl = slice(0, n, stride)
        o = [x[sl].contiguous().cpu().numpy() for x in obs]
        vals_h = cx.last_vals = interp.vals_tensor().cpu().numpy()  # reused by the CPU baseline of the same workload
        threads = max(1, oracle.max_threads() // max(1, cx.world))
        want = cpu_eval(oracle, w, vals_h, o, threads)
        got = out[sl].contiguous().cpu().numpy()
        ibits = np.uint32 if got.dtype == np.float32 else np.uint64
        gb, wb = got.view(ibits), want.view(ibits)
        itorch = torch.int32 if got.dtype == np.float32 else torch.int64
        full = int(out.view(itorch).to(torch.int64).sum().item()) & ((1 << 64) - 1)
        return {
            "sample_points": int(got.size),
            "stride": stride,
            "bit_identical": bool(np.array_equal(gb, wb)),
            "max_abs_diff": float(np.max(np.abs(got - want))),
            "sample_checksum_gpu": f"{int(gb.astype(np.uint64).sum(dtype=np.uint64)):016x}",
            "sample_checksum_oracle": f"{int(wb.astype(np.uint64).sum(dtype=np.uint64)):016x}",
            "output_checksum": f"{full:016x}",
        }
    except Exception as e:  # the oracle is a checker, never a dependency of the measured path
        return {"error": repr(e), "bit_identical": False}


def measure(cx: Ctx, name: str, dtype: str, n: int, steps: int, warmup: int, sustained_steps: int, graph_launches: int = 0,
            keep: bool = False):  # fmt: skip
    """One workload: build/replicate the grid, generate this rank's queries on the device, warm up, time `steps` steps with
    CUDA events on the launching stream (barrier + synchronize on both sides, max over ranks), a longer sustained run,
    then the parity check on every rank. Returns a dict; with keep=True also the live objects for the e2e legs."""
    torch, ib = cx.torch, cx.ib
    w = cx.W.get(name, np.float32 if dtype == "f32" else np.float64)
    tdtype = torch.float32 if dtype == "f32" else torch.float64
    t_setup = time.perf_counter()
    interp, bcast_ms = build_interp(cx, w, dtype)
    obs = gen_queries(cx, w, cx.rank * n, n, tdtype)
    out = torch.empty(n, dtype=tdtype, device=cx.dev)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    stream = cx.stream

    def one_launch():
        interp.eval_torch(obs, out)

    graph, kernels_per_eval = None, 1.0
    if graph_launches > 1:
        # C1-class batches are launch-latency bound (roofline time of 1e6 points: 4.9 us): time `graph_launches` launches
        # replayed as ONE CUDA graph and report the per-launch time.
        one_launch()
        torch.cuda.synchronize()
        side = torch.cuda.Stream(cx.dev)
        graph = torch.cuda.CUDAGraph()
        k0 = ib.launch_count()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for _ in range(graph_launches):
                    interp.eval_torch(obs, out, stream=side)
        kernels_per_eval = (ib.launch_count() - k0) / graph_launches  # replays are not counted by the library
        torch.cuda.synchronize()

    def step():
        if graph is not None:
            graph.replay()
        else:
            one_launch()

    per_step_launches = graph_launches if graph is not None else 1
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()  # ncu --profile-from-start off sees only warm-up + timed region
    for _ in range(max(3, warmup)):
        step()
    interp.status(stream.cuda_stream)

    cx.barrier()
    torch.cuda.synchronize()
    launches0, swept0 = ib.launch_count(), ib.swept_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    with ClockSampler(cx.local_rank) as clocks:
        ev[0].record(stream)
        for k in range(steps):
            step()
            ev[k + 1].record(stream)
        torch.cuda.synchronize()
    cx.barrier()
    launches = (ib.launch_count() - launches0) if graph is None else int(steps * graph_launches * kernels_per_eval)
    swept = ib.swept_launch_count() - swept0
    torch.cuda.cudart().cudaProfilerStop()
    interp.status(stream.cuda_stream)
    total_ms = cx.max_over_ranks(ev[0].elapsed_time(ev[-1]))
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]
    work_per_step = cx.world * n * per_step_launches
    value = work_per_step * steps / (total_ms * 1e-3)

    sustained = None
    if sustained_steps > 0:
        # long enough for the clocks to settle, bounded to ~3 s for slow workloads
        k_sus = int(max(steps, min(sustained_steps, 3000.0 / max(total_ms / steps, 1e-3))))
        cx.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(cx.local_rank) as sclk:
            s0.record(stream)
            for _ in range(k_sus):
                step()
            s1.record(stream)
            torch.cuda.synchronize()
        sus_ms = cx.max_over_ranks(s0.elapsed_time(s1))
        interp.status(stream.cuda_stream)
        sustained = {"steps": k_sus, "value": work_per_step * k_sus / (sus_ms * 1e-3), "ms_per_step": sus_ms / k_sus,
                     "sm_mhz": sclk.summary().get("sm_mhz")}  # fmt: skip

    parity = parity_check(cx, w, interp, obs, out, n)
    parity["every_rank_bit_identical"] = cx.all_true(bool(parity.get("bit_identical")))
    parity["ranks_checked"] = cx.world

    mean_ms = float(np.mean(step_ms))
    abytes = cx.W.algorithmic_bytes(w, n) * per_step_launches
    achieved = abytes / (mean_ms * 1e-3) / 1e9
    traffic, traffic_src = ncu_traffic(w.name, dtype, n)
    n_launch = launches / max(steps, 1)
    res = {
        "workload": w.name,
        "dtype": dtype,
        "method": w.method,
        "grid": w.dims,
        "grid_kind": "rectilinear" if w.rect else "regular",
        "points_per_gpu": n,
        "n_gpus": cx.world,
        "value": value,
        "unit": UNIT,
        "steps": steps,
        "ms_per_step": total_ms / steps,
        "sustained": sustained,
        "roofline": {
            "bound": "hbm",
            "achieved": achieved,
            "peak": cx.peak,
            "unit": "GB/s",
            "frac": achieved / cx.peak,
            "traffic": traffic * per_step_launches if traffic is not None else None,
            "traffic_source": traffic_src,
            "peak_source": cx.peak_src,
            "kernel": ("the evaluation kernel (one launch per step)" if n_launch <= 1.01 else
                       f"all {n_launch:g} launches of a step (CUDA graph of {graph_launches} evaluations)" if graph is not None else
                       f"all {n_launch:g} launches of a step (sort + evaluation of the bin-swept path, or the slab passes): "
                       "algorithmic bytes per step over the step's device time"),
            "algorithmic_bytes_per_launch": abytes,
            "kernel_ms": mean_ms,
            "binding": binding_ceiling(w.name, n * per_step_launches / (mean_ms * 1e-3)),
        },
        "gpu_launches": int(launches),
        "swept_launches": int(swept),
        "clocks": clocks.summary(),
        "parity": parity,
        "setup_s": setup_s,
        "grid_broadcast_ms": bcast_ms,
        "step_ms": step_ms if steps <= 40 else step_ms[:20] + step_ms[-20:],
    }  # fmt: skip
    if graph is not None:
        res["graph"] = {"launches_per_replay": graph_launches, "per_launch_us": total_ms / steps / graph_launches * 1e3,
                        "roofline_us_per_launch": cx.W.algorithmic_bytes(w, n) / (cx.peak * 1e9) * 1e6}  # fmt: skip
        del graph
    if keep:
        return res, (w, interp, obs, out)
    interp.close()
    del obs, out, interp
    torch.cuda.empty_cache()
    return res


def e2e_legs(cx: Ctx, w, interp, obs, out, n: int, dtype: str):
    """The same metric through the host-buffer C call (`Interpolator.eval` -> interpn_b200_interp_eval_host_*), copies inside
    the timed region: pinned host arrays (the contract's e2e), pageable numpy arrays (what a drop-in caller passes), and at
    N > 1 one single call from rank 0 over all GPUs."""
    torch, ib, args = cx.torch, cx.ib, cx.args
    tdtype = out.dtype
    item = w.dtype.itemsize
    api = f"interpn_b200.Interpolator.eval -> interpn_b200_interp_eval_host_{dtype}"

    def timed(nobs, nout, steps):
        interp.eval(nobs, nout)  # warm-up: sizes the copy pipeline / staging
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            interp.eval(nobs, nout)
        return cx.max_over_ranks(time.perf_counter() - t0)

    hobs = [torch.empty(n, dtype=tdtype).pin_memory() for _ in range(w.ndims)]
    hout = torch.empty(n, dtype=tdtype).pin_memory()
    for d in range(w.ndims):
        hobs[d].copy_(obs[d])
    torch.cuda.synchronize()
    dt = timed([h.numpy() for h in hobs], hout.numpy(), args.e2e_steps)
    e2e = {
        "value": cx.world * n * args.e2e_steps / dt,
        "unit": UNIT,
        "h2d_bytes_per_step": int(n * w.ndims * item),
        "d2h_bytes_per_step": int(n * item),
        "ms_per_step": dt / args.e2e_steps * 1e3,
        "steps": args.e2e_steps,
        "host_memory": "pinned",
        "api": api + " (pinned host buffers DMA'd in place, 3-slot copy/compute pipeline per device)",
        "matches_device_path": cx.all_true(bool(torch.equal(hout.to(cx.dev), out))),
    }
    # pageable: plain numpy arrays
    pobs = [np.empty(n, dtype=w.dtype) for _ in range(w.ndims)]
    for d in range(w.ndims):
        pobs[d][:] = hobs[d].numpy()
    pout = np.empty(n, dtype=w.dtype)
    dt = timed(pobs, pout, args.e2e_steps)
    e2e["pageable"] = {
        "value": cx.world * n * args.e2e_steps / dt,
        "ms_per_step": dt / args.e2e_steps * 1e3,
        "host_memory": f"pageable numpy arrays -> pinned staging ring, {ib.copy_threads()} copy threads per process",
        "fraction_of_pinned": (cx.world * n * args.e2e_steps / dt) / e2e["value"],
        "matches_device_path": cx.all_true(bool(np.array_equal(pout.view(np.uint8), hout.numpy().view(np.uint8)))),
    }
    del pobs, pout
    if cx.distributed:
        # ONE C call from rank 0 over all N GPUs (in-process sharding, grid replicated device-to-device on first use);
        # the other ranks wait on the CPU (a store key, not an NCCL barrier whose kernel would spin on their GPU).
        store = cx.dist.distributed_c10d._get_default_store()
        m = min(n, 50_000_000)
        cx.barrier()
        if cx.rank == 0:
            try:
                ib.set_host_devices(cx.world)
                tot = m * cx.world
                sobs = [torch.empty(tot, dtype=tdtype).pin_memory() for _ in range(w.ndims)]
                sout = torch.empty(tot, dtype=tdtype).pin_memory()
                for d in range(w.ndims):
                    for r in range(cx.world):
                        sobs[d][r * m : (r + 1) * m].copy_(obs[d][:m])
                torch.cuda.synchronize()
                nobs, nout = [h.numpy() for h in sobs], sout.numpy()
                interp.eval(nobs, nout)  # replicates the grid, sizes every device's slots
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    interp.eval(nobs, nout)
                dt1 = time.perf_counter() - t0
                ok = all(bool(torch.equal(sout[r * m : (r + 1) * m], hout[:m])) for r in range(cx.world))
                e2e["single_call"] = {
                    "value": tot * args.e2e_steps / dt1,
                    "points": tot,
                    "devices": ib.host_devices(),
                    "ms_per_step": dt1 / args.e2e_steps * 1e3,
                    "api": api + ": one call from one process, the batch sharded over all devices inside the library",
                    "matches_device_path": ok,
                }
                del sobs, sout
            except Exception as e:
                e2e["single_call"] = {"error": repr(e)}
            finally:
                ib.set_host_devices(1)
                store.set("interpn_b200_single_call_done", "1")
        else:
            store.wait(["interpn_b200_single_call_done"])
        cx.barrier()
    del hobs, hout
    return e2e


def suite_plan(cx: Ctx):
    """(name, dtype, points per GPU, steps, graph launches) of the non-headline BASELINE configurations."""
    a = cx.args
    mode = a.suite
    if mode == "auto":
        mode = "none" if a.workload != HEADLINE or a.points else ("all" if cx.world == 1 else "c4")
    if mode == "none":
        return []
    c4 = ("c4_linear6d_reg24", "f64", 125_000_000, 5, 0)
    if mode == "c4":
        return [c4]
    plan = [("c1_linear3d_reg20", "f64", 1_000_000, 10, 100), ("c3_linear4d_rect64", "f64", 100_000_000, 5, 0),
            ("c3_cubic4d_rect64", "f64", 100_000_000, 3, 0), c4]  # fmt: skip
    for nm in ("c5_nearest2d_reg1024", "c5_nearest3d_reg128", "c5_nearest2d_rect1024", "c5_nearest3d_rect128"):
        for dt in ("f64", "f32"):
            plan.append((nm, dt, 1_000_000_000, 5, 0))
    return plan


def run_b200(args):
    cx = Ctx(args)
    torch = cx.torch
    w0 = cx.W.get(args.workload, np.float32 if args.dtype == "f32" else np.float64)
    n = args.points or (125_000_000 if w0.name == "c4_linear6d_reg24" else min(w0.n_full, 100_000_000))
    head, (w, interp, obs, out) = measure(cx, args.workload, args.dtype, n, args.steps, args.warmup, args.sustained_steps,
                                          graph_launches=args.graph_launches, keep=True)  # fmt: skip
    head_vals = cx.last_vals
    e2e = None if args.no_e2e else e2e_legs(cx, w, interp, obs, out, n, args.dtype)
    interp.close()
    del obs, out, interp
    torch.cuda.empty_cache()

    suite = []
    for name, dt, pts, steps, graph_launches in suite_plan(cx):
        try:
            r = measure(cx, name, dt, pts, steps, 3, min(args.sustained_steps, 200), graph_launches=graph_launches)
            for k in ("setup_s", "step_ms"):
                r.pop(k, None)
            if cx.rank == 0 and cx.world == 1 and not args.no_cpu_baseline:
                wk = cx.W.get(name, np.float32 if dt == "f32" else np.float64)
                try:
                    r["cpu_baseline"] = cpu_baseline_both(wk, 1.5, vals=cx.last_vals if cx.last_vals is not None and cx.last_vals.size == wk.nvals else None,
                                                           max_points=4_000_000)  # generating the sample on the host dominates beyond that
                except Exception as e:
                    r["cpu_baseline"] = {"error": repr(e)}
            suite.append(r)
        except Exception as e:
            suite.append({"workload": name, "dtype": dt, "error": repr(e)})
            torch.cuda.empty_cache()

    if cx.rank != 0:
        if cx.distributed:
            cx.dist.destroy_process_group()
        return 0

    line = {
        "metric": METRIC,
        "value": head["value"],
        "unit": UNIT,
        "n_gpus": cx.world,
        "steps": args.steps,
        "warmup": max(3, args.warmup),
        "ms_per_step": head["ms_per_step"],
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
            "parallelism": f"query batch sharded over {cx.world} GPU(s), grid replicated by one NCCL broadcast" if cx.distributed else "single GPU",
        },
        "sustained": head["sustained"],
        "roofline": head["roofline"],
        "e2e": e2e,
        "gpu_launches": head["gpu_launches"],
        "swept_launches": head["swept_launches"],
        "clocks": head["clocks"],
        "parity": head["parity"],
        "setup_s": head["setup_s"],
        "grid_broadcast_ms": head["grid_broadcast_ms"],
        "step_ms": head["step_ms"],
    }  # fmt: skip
    if "graph" in head:
        line["graph"] = head["graph"]
    if not args.no_cpu_baseline and cx.world == 1:
        try:
            line["cpu_baseline"] = cpu_baseline_both(w, args.cpu_seconds, vals=head_vals)
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
    if suite:
        line["workloads"] = suite
    print(json.dumps(line), flush=True)
    if cx.distributed:
        cx.dist.destroy_process_group()
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
