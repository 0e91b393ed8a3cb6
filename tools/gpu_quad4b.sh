#!/bin/bash
tag=${1:-quad4b}; out=gpurun_out/$tag; mkdir -p $out
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
run c2_q4_m5 c2_cubic3d_reg100 100000000 INTERPN_B200_QUAD4_MINB=5
run c2_q4_m6 c2_cubic3d_reg100 100000000 INTERPN_B200_QUAD4_MINB=6
run c2_q4_m4 c2_cubic3d_reg100 100000000 INTERPN_B200_QUAD4_MINB=4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'quad4' -s 3 -c 1 -o $out/c2_quad4 -f python bench.py --workload c2_cubic3d_reg100 --points 20000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c2_quad4_ncu.log 2>&1; echo "ncu exit $?"
