#!/bin/bash
# Parity + multilinear workloads with the patch layout (2x2 sectors) against the row-pair window of the previous commit.
tag=${1:-lin}; out=gpurun_out/$tag; mkdir -p $out
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
run xl3 x_linear3d_reg100 100000000 A=1
run xl4 x_linear4d_reg32 100000000 A=1
run xl4r x_linear4d_rect32 100000000 A=1
run c1 c1_linear3d_reg20 1000000 A=1
run c1_1e8 c1_linear3d_reg20 100000000 A=1
run c4 c4_linear6d_reg24 100000000 A=1
run c3lin c3_linear4d_rect64 100000000 A=1
run c3lin_nowin c3_linear4d_rect64 100000000 INTERPN_B200_WINDOW_MB=0
