#!/bin/bash
# round-2 experiment 6: software-pipelined group loop of the cubic quad kernel against the unpipelined build, per register budget
out=gpurun_out/${TAG:-r2_exp6}; mkdir -p $out
L=$PWD/interpn_b200
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_instantiations.py -m gpu -x -q -k "cubic or Cubic or baseline or plateau" ) > $out/pytest_cubic.log 2>&1; tail -n 3 $out/pytest_cubic.log
line() { # label env... -- bench args
  label=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py "$@" --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 --suite none > $out/$label.json 2> $out/$label.err
  python - "$out/$label.json" "$label" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "%.3f G/s" % (d["value"] / 1e9), "ms %.3f" % d["ms_per_step"], "parity", d["parity"].get("bit_identical"), "launches", d["gpu_launches"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for lib in libinterpn_b200 lib_nopipe; do
for mb in ${MBS:-2 3 4}; do
  line c2_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload c2_cubic3d_reg100 --points 100000000
  line x3rect_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic3d_rect100 --points 50000000
  line x4reg_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic4d_reg32 --points 50000000
  line x4rect_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic4d_rect32 --points 30000000
  line c3c_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload c3_cubic4d_rect64 --points 50000000
done; done
