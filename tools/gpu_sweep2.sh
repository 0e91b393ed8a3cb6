#!/bin/bash
tag=${1:-sweep}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "swept or baseline" > $out/pytest_sweep.log 2>&1; echo "pytest exit $?" >> $out/pytest_sweep.log
tail -3 $out/pytest_sweep.log
run() {
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
run c4_auto c4_linear6d_reg24 100000000 A=1
run c4_chunk16m c4_linear6d_reg24 100000000 INTERPN_B200_SWEEP_CHUNK=16777216
run c4_chunk64m c4_linear6d_reg24 100000000 INTERPN_B200_SWEEP_CHUNK=67108864
run c4_slab3m c4_linear6d_reg24 100000000 INTERPN_B200_SWEEP_SLAB_KB=3072
run c4_slab12m c4_linear6d_reg24 100000000 INTERPN_B200_SWEEP_SLAB_KB=12288
run c4_nowin c4_linear6d_reg24 100000000 INTERPN_B200_WINDOW_MB=64
run c3lin_swept c3_linear4d_rect64 100000000 INTERPN_B200_SWEEP_MIN_ROWS=4
run c3lin_direct c3_linear4d_rect64 100000000 A=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sweep_scatter' -s 3 -c 1 -o $out/c4_scatter -f python bench.py --workload c4_linear6d_reg24 --points 40000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c4_scatter.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'linear_kernel' -s 3 -c 1 -o $out/c4_linear -f python bench.py --workload c4_linear6d_reg24 --points 40000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c4_linear.log 2>&1; echo "ncu exit $?"
