#!/bin/bash
# f32 bench lines (C5 names f32 and f64; the other kernels for reference)
tag=${1:-f32}; out=gpurun_out/$tag; mkdir -p $out
for spec in c5_nearest3d_reg128:100000000 c5_nearest2d_reg1024:100000000 c5_nearest3d_rect128:100000000 c5_nearest2d_rect1024:100000000 c2_cubic3d_reg100:100000000 x_linear3d_reg100:100000000 c3_cubic4d_rect64:20000000; do
  wl=${spec%%:*}; pts=${spec##*:}
  timeout 900 python bench.py --dtype f32 --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/f32_$wl.json 2> $out/f32_$wl.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/f32_$wl.json").read().strip().splitlines()[-1])
    print("f32 $wl", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), d["dtype"])
except Exception as e:
    print("f32 $wl FAILED", e); print(open("$out/f32_$wl.err").read()[-800:])
PY
done
