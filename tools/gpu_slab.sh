#!/bin/bash
tag=${1:-slab}; out=gpurun_out/$tag; mkdir -p $out
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
for kb in 6144 12288 24576 49152 98304; do
  run c4_slab$kb c4_linear6d_reg24 100000000 INTERPN_B200_SWEEP_SLAB_KB=$kb
done
for kb in 6144 24576 49152; do
  run c3c_slab$kb c3_cubic4d_rect64 20000000 INTERPN_B200_SWEEP_SLAB_KB=$kb
done
run c4_chunk25 c4_linear6d_reg24 100000000 INTERPN_B200_SWEEP_CHUNK=25000000
run c4_chunk100 c4_linear6d_reg24 100000000 INTERPN_B200_SWEEP_CHUNK=100000000
