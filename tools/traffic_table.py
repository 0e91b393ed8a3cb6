#!/usr/bin/env python
"""Builds profiles/ncu_traffic.json from the metrics-only ncu passes of tools/profile_round.sh: for every workload the
DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and the device time of ONE step, summed over the step's launches
(the pass profiles 3 warm-up steps + 1 timed step: the last quarter of the launches is the timed step).
Usage: python tools/traffic_table.py gpurun_out/r2_prof profiles/ncu_traffic.json"""
import csv
import glob
import json
import os
import re
import sys


def parse(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10 and r[0].isdigit()]
    launches = {}
    for r in rows:
        d = launches.setdefault(int(r[0]), {"kernel": r[4].split("(")[0]})
        name, unit, val = r[-3], r[-2], float(r[-1].replace(",", ""))
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1)
        d[name] = val * scale
    ids = sorted(launches)
    step = ids[len(ids) * 3 // 4 :]
    kern = {}
    for i in step:
        L = launches[i]
        k = kern.setdefault(L["kernel"][:60], [0, 0.0, 0.0])
        k[0] += 1
        k[1] += L.get("dram__bytes_read.sum", 0) + L.get("dram__bytes_write.sum", 0)
        k[2] += L.get("gpu__time_duration.sum", 0)
    return {"launches_per_step": len(step), "bytes": int(sum(v[1] for v in kern.values())), "device_ms": sum(v[2] for v in kern.values()),
            "kernels": {k: {"launches": v[0], "dram_bytes": int(v[1]), "ms": round(v[2], 4)} for k, v in kern.items()}}


def main():
    src, dst = sys.argv[1], sys.argv[2]
    points = {}
    for line in open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profile_round.sh")):
        m = re.match(r"\s*traffic (\S+) (f\d\d) (\d+)", line)
        if m:
            points[(m.group(1), m.group(2))] = int(m.group(3))
    table = {"_comment": "DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and device time of ONE step of each bench workload, summed over "
             "the step's launches, from the metrics-only ncu pass of tools/profile_round.sh (serialised launches, cold caches between "
             "kernels); `points` is the batch the pass ran on — bench.py scales `bytes` to its own batch and copies it into roofline.traffic"}
    for path in sorted(glob.glob(os.path.join(src, "traffic_*.csv"))):
        m = re.match(r"traffic_(.+)_(f\d\d)\.csv", os.path.basename(path))
        wl, dt = m.group(1), m.group(2)
        pts = points.get((wl, dt))
        if pts is None:
            for (w2, d2), p in points.items():
                if w2 == wl:
                    pts = p
        try:
            rec = parse(path)
        except Exception as e:
            print("skip", path, e)
            continue
        rec.update(points=pts, source=f"profiles/r2_traffic/{os.path.basename(path)} (ncu metrics pass of tools/profile_round.sh)")
        table[f"{wl}:{dt}"] = rec
        print(wl, dt, pts, "%.3f GB" % (rec["bytes"] / 1e9), "%.3f ms" % rec["device_ms"], rec["launches_per_step"], "launches")
    json.dump(table, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()
