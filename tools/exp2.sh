#!/bin/bash
# round-2 experiment 2: 4-D cubic stash variants x register budgets; launch lists + scatter ncu of the swept paths
out=gpurun_out/r2_exp2; mkdir -p $out
L=$PWD/interpn_b200
line() { # label env... -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py "$@" --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 > $out/$label.json 2> $out/$label.err
  python - "$out/$label.json" "$label" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "%.3f G/s" % (d["value"] / 1e9), "ms %.3f" % d["ms_per_step"], "parity", d["parity"].get("bit_identical"), "launches", d["gpu_launches"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for lib in libinterpn_b200 lib_nostash lib_inner2; do
  for mb in 2 3 4; do
    line x4reg_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic4d_reg32 --points 50000000
    line x4rect_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic4d_rect32 --points 30000000
    line c3c_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload c3_cubic4d_rect64 --points 50000000
  done
done
