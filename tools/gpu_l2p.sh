#!/bin/bash
tag=${1:-l2p}; out=gpurun_out/$tag; mkdir -p $out
run() {
  name=$1; wl=$2; pts=$3; shift 3
  env "$@" timeout 900 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/$name.json").read().strip().splitlines()[-1])
    print("$name", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), "launches", d["gpu_launches"], "swept", d.get("swept_launches"))
except Exception as e:
    print("$name FAILED", e); print(open("$out/$name.err").read()[-600:])
PY
}
run c5n3 c5_nearest3d_reg128 100000000 A=1
run c5n2 c5_nearest2d_reg1024 100000000 A=1
run c1_1e8 c1_linear3d_reg20 100000000 A=1
run c1_1e8_win c1_linear3d_reg20 100000000 INTERPN_B200_WINDOW_MIN_KB=0
run c3lin c3_linear4d_rect64 100000000 A=1
run c5n3_rect c5_nearest3d_rect128 100000000 A=1
