#!/bin/bash
# A/B of the cubic kernels on C2 (and a 4-D regular grid): one-point-per-quad vs four-points-per-quad, register budgets.
tag=${1:-quad4}; out=gpurun_out/$tag; mkdir -p $out
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
run c2_old c2_cubic3d_reg100 100000000 INTERPN_B200_CUBIC_QUAD=1
run c2_q4_m4 c2_cubic3d_reg100 100000000 INTERPN_B200_QUAD4_MINB=4
run c2_q4_m3 c2_cubic3d_reg100 100000000 INTERPN_B200_QUAD4_MINB=3
run c2_q4_m2 c2_cubic3d_reg100 100000000 INTERPN_B200_QUAD4_MINB=2
for extra in "$@"; do :; done
