#!/bin/bash
# Bin-swept vs direct evaluation on the grids that exceed L2 (C3, C4).
# Usage (under gpurun): bash tools/gpu_sweep.sh <tag>
tag=${1:-sweep}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "swept or baseline" > $out/pytest_sweep.log 2>&1; echo "pytest exit $?" >> $out/pytest_sweep.log
tail -5 $out/pytest_sweep.log
run() {  # name workload points extra-env...
  name=$1; wl=$2; pts=$3; shift 3
  env "$@" timeout 900 python bench.py --workload $wl --points $pts --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/$name.json").read().strip().splitlines()[-1])
    print("$name", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), "launches", d["gpu_launches"], "swept", d.get("swept_launches"))
except Exception as e:
    print("$name FAILED", e); print(open("$out/$name.err").read()[-600:])
PY
}
for wl in c3_linear4d_rect64 c3_cubic4d_rect64 c4_linear6d_reg24; do
  pts=100000000; [ $wl = c3_cubic4d_rect64 ] && pts=20000000
  run ${wl}_direct $wl $pts INTERPN_B200_SWEEP_MIN_MB=100000000
  run ${wl}_swept $wl $pts
  run ${wl}_swept_t512k $wl $pts INTERPN_B200_SWEEP_TILE=524288
  run ${wl}_swept_slab2m $wl $pts INTERPN_B200_SWEEP_SLAB_KB=2048
done
