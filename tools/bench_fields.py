#!/usr/bin/env python
"""Fused multi-field evaluation against field-by-field calls: K fields over one grid and one query
batch, CUDA-event timed, bit-compared. Usage (on the GPU box): python tools/bench_fields.py [K]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import interpn_b200 as ib
from interpn_b200 import workloads as W

K = int(sys.argv[1]) if len(sys.argv) > 1 else 6
dev = torch.device("cuda", 0)
res = []
def small(name, method, ndims, size, rect):
    """A grid small enough that K of its gathered copies sit in L2 together."""
    if rect:
        return W._rectilinear(name, method, ndims, size, np.float64, 100_000_000, oob_fraction=0.05)
    return W._regular(name, method, [size] * ndims, [0.0] * ndims, [1.0] * ndims, np.float64, 100_000_000, oob_fraction=0.05)


CASES = [(W.get("x_linear3d_reg100"), 50_000_000), (W.get("c5_nearest3d_reg128"), 50_000_000), (W.get("c5_nearest3d_rect128"), 50_000_000),
         (small("linear3d_reg48", "linear", 3, 48, False), 50_000_000), (small("linear4d_rect16", "linear", 4, 16, True), 50_000_000),
         (small("nearest3d_reg64", "nearest", 3, 64, False), 50_000_000), (small("nearest2d_rect512", "nearest", 2, 512, True), 50_000_000)]
for w, n in CASES:
    name = w.name
    obs = w.queries(0, n, "torch", dev)
    rng = np.random.default_rng(1)
    interps = []
    for k in range(K):
        v = rng.standard_normal(w.nvals)
        interps.append(ib.Interpolator.rectilinear(w.method, w.grids, v, True) if w.rect
                       else ib.Interpolator.regular(w.method, w.dims, w.starts, w.steps, v, True))
    outs_f = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(K)]
    outs_s = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(K)]

    def fused():
        ib.Interpolator.eval_fields_torch(interps, obs, outs_f)

    def separate():
        for it, o in zip(interps, outs_s):
            it.eval_torch(obs, o)

    t = {}
    for label, fn in (("fused", fused), ("separate", separate)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t[label] = e0.elapsed_time(e1) / 5
    same = all(torch.equal(a.view(torch.int64), b.view(torch.int64)) for a, b in zip(outs_f, outs_s))
    row = {"workload": name, "fields": K, "points": n, "fused_ms": t["fused"], "separate_ms": t["separate"],
           "fused_field_points_per_s": K * n / t["fused"] * 1e3, "separate_field_points_per_s": K * n / t["separate"] * 1e3,
           "speedup": t["separate"] / t["fused"], "bit_identical": bool(same)}
    res.append(row)
    print(json.dumps(row), flush=True)
    for it in interps:
        it.close()
