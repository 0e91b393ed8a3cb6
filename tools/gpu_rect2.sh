#!/bin/bash
tag=${1:-rect2}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
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
run c3c_m2 c3_cubic4d_rect64 20000000 INTERPN_B200_QUAD4_MINB=2
run c3c_m3 c3_cubic4d_rect64 20000000 INTERPN_B200_QUAD4_MINB=3
run x3r_m2 x_cubic3d_rect100 50000000 INTERPN_B200_QUAD4_MINB=2
run x3r_m3 x_cubic3d_rect100 50000000 INTERPN_B200_QUAD4_MINB=3
run x4r_m2 x_cubic4d_rect32 20000000 INTERPN_B200_QUAD4_MINB=2
run xl3 x_linear3d_reg100 100000000 A=1
run c1_1e8 c1_linear3d_reg20 100000000 A=1
run c4 c4_linear6d_reg24 100000000 A=1
