#!/bin/bash
# Slab passes against the direct kernel on regular grids a little beyond L2. Usage (under gpurun): bash tools/gpu_slabreg.sh <tag>
tag=${1:-slabreg}; out=gpurun_out/$tag; mkdir -p $out
for wl in x_linear3d_reg256 x_linear4d_reg64 x_linear5d_reg26; do
  for kb in 0 46080; do
    name=${wl}_pass$kb
    INTERPN_B200_SLAB_PASS_KB=$kb timeout 600 python bench.py --workload $wl --points 100000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/$name.json 2> $out/$name.err
    python - <<PY
import json
try:
    d = json.loads(open("$out/$name.json").read().strip().splitlines()[-1])
    print("$name", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), "launches", d["gpu_launches"], "swept", d.get("swept_launches"))
except Exception as e:
    print("$name FAILED", e); print(open("$out/$name.err").read()[-600:])
PY
  done
done
timeout 600 python bench.py --workload c3_linear4d_rect64 --points 100000000 --steps 5 --warmup 3 --no-cpu-baseline > $out/c3_linear.json 2> $out/c3_linear.err; tail -c 1200 $out/c3_linear.json
