#!/bin/bash
tag=${1:-axes}; out=gpurun_out/$tag; mkdir -p $out
run() {
  name=$1; wl=$2; pts=$3; shift 3
  env "$@" timeout 900 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/$name.json").read().strip().splitlines()[-1])
    print("$name", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"))
except Exception as e:
    print("$name FAILED", e); print(open("$out/$name.err").read()[-600:])
PY
}
run c3lin_smem c3_linear4d_rect64 100000000 A=1
run c3lin_global c3_linear4d_rect64 100000000 INTERPN_B200_AXES_SMEM_KB=0
run xl4r_smem x_linear4d_rect32 100000000 A=1
run xl4r_global x_linear4d_rect32 100000000 INTERPN_B200_AXES_SMEM_KB=0
run c5n3r_global c5_nearest3d_rect128 100000000 INTERPN_B200_AXES_SMEM_KB=0
