#!/bin/bash
tag=${1:-lin2}; out=gpurun_out/$tag; mkdir -p $out
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
run c1 c1_linear3d_reg20 1000000 A=1
run c1_1e8 c1_linear3d_reg20 100000000 A=1
run c3lin_patch_dram c3_linear4d_rect64 100000000 INTERPN_B200_WINDOW_L2_MB=4096
run c3lin_swept c3_linear4d_rect64 100000000 INTERPN_B200_SWEEP_MIN_ROWS=4
run c3lin c3_linear4d_rect64 100000000 A=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'linear_kernel' -s 3 -c 1 -o $out/xl3 -f python bench.py --workload x_linear3d_reg100 --points 20000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/xl3_ncu.log 2>&1; echo "ncu exit $?"
