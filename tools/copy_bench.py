#!/usr/bin/env python
"""Copy-only microbenchmark: what the HOST side of the end-to-end path can deliver, without any kernel.

The end-to-end metric of bench.py moves (N+1) arrays per point over PCIe. On one GPU it runs at the PCIe rate; on 8 GPUs
of one box its scaling stops early (round 1: 0.29 efficiency). This script says which host-side resource is the limit:
per device alone — pinned H2D, D2H and both directions at once — then ALL devices at once from one process (one thread
and two streams per device), and the host's own memcpy bandwidth with 1..k threads (the rate at which pageable caller
memory can be staged). Prints one JSON object.

    python tools/copy_bench.py [--mb 1024] [--reps 5]
"""

import argparse
import json
import threading
import time

import torch


def bw(nbytes, seconds):
    return nbytes / seconds / 1e9


def device_alone(dev, nbytes, reps):
    torch.cuda.set_device(dev)
    h_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    d_a = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
    d_b = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{dev}")
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    res = {}

    def run(h2d, d2h):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize(dev)
        return time.perf_counter() - t0

    run(True, True)
    res["h2d_gbs"] = bw(nbytes * reps, run(True, False))
    res["d2h_gbs"] = bw(nbytes * reps, run(False, True))
    res["both_gbs"] = bw(2 * nbytes * reps, run(True, True))
    return res, (h_in, h_out, d_a, d_b, s1, s2)


def all_devices(bufs, nbytes, reps, h2d=True, d2h=True):
    ndev = len(bufs)
    barrier = threading.Barrier(ndev + 1)

    def worker(dev):
        h_in, h_out, d_a, d_b, s1, s2 = bufs[dev]
        torch.cuda.set_device(dev)
        barrier.wait()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_a.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize(dev)
        barrier.wait()

    th = [threading.Thread(target=worker, args=(d,)) for d in range(ndev)]
    for t in th:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    return bw((int(h2d) + int(d2h)) * nbytes * reps * ndev, dt)


def host_memcpy(nbytes, threads, reps):
    src = [torch.empty(nbytes // threads, dtype=torch.uint8).random_() for _ in range(threads)]
    dst = [torch.empty(nbytes // threads, dtype=torch.uint8) for _ in range(threads)]
    for d in dst:
        d.zero_()
    barrier = threading.Barrier(threads + 1)

    def worker(i):
        barrier.wait()
        for _ in range(reps):
            dst[i].copy_(src[i])
        barrier.wait()

    th = [threading.Thread(target=worker, args=(i,)) for i in range(threads)]
    for t in th:
        t.start()
    barrier.wait()
    t0 = time.perf_counter()
    barrier.wait()
    dt = time.perf_counter() - t0
    for t in th:
        t.join()
    return bw((nbytes // threads) * threads * reps, dt)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    nbytes = a.mb << 20
    ndev = torch.cuda.device_count()
    torch.set_num_threads(1)
    out = {"devices": ndev, "bytes_per_copy": nbytes, "reps": a.reps, "per_device": [], "host_cpus": torch.get_num_threads()}
    import os

    out["host_cpus"] = len(os.sched_getaffinity(0))
    bufs = []
    for d in range(ndev):
        r, b = device_alone(d, nbytes, a.reps)
        out["per_device"].append(r)
        bufs.append(b)
    if ndev > 1:
        out["all_devices_h2d_gbs"] = all_devices(bufs, nbytes, a.reps, True, False)
        out["all_devices_d2h_gbs"] = all_devices(bufs, nbytes, a.reps, False, True)
        out["all_devices_both_gbs"] = all_devices(bufs, nbytes, a.reps, True, True)
        for k in (2, 4):
            if k < ndev:
                out[f"first_{k}_devices_both_gbs"] = all_devices(bufs[:k], nbytes, a.reps, True, True)
    out["host_memcpy_gbs"] = {str(t): host_memcpy(nbytes, t, 3) for t in (1, 2, 4, 8, 16, 32) if t <= out["host_cpus"]}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
