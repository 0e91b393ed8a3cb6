#!/bin/bash
# Bench a list of workloads against several builds of the library (kernel-tuning experiments).
# Usage (under gpurun): bash tools/gpu_variants.sh <tag> "<lib1> <lib2> ..." "<wl1> <wl2> ..."
tag=$1; libs=$2; wls=$3
out=gpurun_out/$tag; mkdir -p $out
for lib in $libs; do
  for wl in $wls; do
    pts=100000000; [ $wl = c3_cubic4d_rect64 ] && pts=20000000; [ $wl = c1_linear3d_reg20 ] && pts=1000000
    name=$(basename $lib .so)
    INTERPN_B200_LIBRARY=$PWD/$lib timeout 600 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/${name}_$wl.json 2> $out/${name}_$wl.err
    python - <<PY
import json
try:
    d = json.loads(open("$out/${name}_$wl.json").read().strip().splitlines()[-1])
    print("$name $wl", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$name $wl FAILED", e)
PY
  done
done
