#!/bin/bash
# round-2 experiment 4: coefficient-layout cubic kernels — parity tests, then bench lines per register budget
out=gpurun_out/${TAG:-r2_exp4}; mkdir -p $out
( time timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_instantiations.py tests/test_ref_docs_golden.py tests/test_gpu_reference_suite.py -m gpu -x -q -k "cubic or Cubic or docs or suite or baseline or plateau or sweep or swept" ) > $out/pytest_cubic.log 2>&1; tail -n 6 $out/pytest_cubic.log
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
for mb in 0 ${MBS:-3 4 5 2}; do
  line c2_mb$mb INTERPN_B200_QUAD4_MINB=$mb -- --workload c2_cubic3d_reg100 --points 100000000
  line x3rect_mb$mb INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic3d_rect100 --points 50000000
  line x4reg_mb$mb INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic4d_reg32 --points 50000000
  line x4rect_mb$mb INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic4d_rect32 --points 30000000
  line c3c_mb$mb INTERPN_B200_QUAD4_MINB=$mb -- --workload c3_cubic4d_rect64 --points 50000000
done
line c2_fma INTERPN_B200_ARITHMETIC=fma -- --workload c2_cubic3d_reg100 --points 100000000 --arithmetic fma
