#!/bin/bash
# parity suite, C2 after the location rewrite, and points-per-thread variants of the multilinear kernels
tag=${1:-mix}; out=gpurun_out/$tag; mkdir -p $out
bash tools/gpu_tests.sh $tag c2_cubic3d_reg100:100000000 x_cubic4d_reg32:50000000
for lib in "" interpn_b200/variants/lin_p2_2.so interpn_b200/variants/lin_p4_4.so interpn_b200/variants/lin_p2_1.so; do
  for spec in x_linear3d_reg100:100000000 x_linear4d_reg32:100000000 x_linear4d_rect32:100000000 c1_linear3d_reg20:100000000 c3_linear4d_rect64:100000000; do
    wl=${spec%%:*}; pts=${spec##*:}; name=$(basename "${lib:-default}" .so)
    if [ -n "$lib" ]; then export INTERPN_B200_LIBRARY=$PWD/$lib; else unset INTERPN_B200_LIBRARY; fi
    timeout 600 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/${name}_$wl.json 2> $out/${name}_$wl.err
    python - <<PY
import json
try:
    d = json.loads(open("$out/${name}_$wl.json").read().strip().splitlines()[-1])
    print("$name $wl", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"))
except Exception as e:
    print("$name $wl FAILED", e)
PY
  done
done
