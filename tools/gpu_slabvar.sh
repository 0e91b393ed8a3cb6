#!/bin/bash
# C3-linear against several builds of the library (slab-pass kernel hooks). Usage (under gpurun): bash tools/gpu_slabvar.sh <tag> <lib.so>...
tag=$1; shift; out=gpurun_out/$tag; mkdir -p $out
for lib in "$@"; do
  name=$(basename $lib .so)
  INTERPN_B200_LIBRARY=$PWD/$lib INTERPN_B200_SLAB_PASS_KB=${PASS_KB:-46080} timeout 600 python bench.py --workload c3_linear4d_rect64 --points 100000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/$name.json").read().strip().splitlines()[-1])
    print("$name", "%.3f Gpts/s" % (d["value"]/1e9), "bit_identical", d["parity"].get("bit_identical"), "launches", d["gpu_launches"])
except Exception as e:
    print("$name FAILED", e); print(open("$out/$name.err").read()[-600:])
PY
done
