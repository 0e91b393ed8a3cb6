#!/bin/bash
# round-2 experiment 3: software-pipelined 4-D cubic rows; pageable staging with/without non-temporal stores
out=gpurun_out/r2_exp3; mkdir -p $out
L=$PWD/interpn_b200
line() {
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
for lib in libinterpn_b200 lib_prefetch; do
  for mb in 2 1; do
    [ $lib = libinterpn_b200 ] && [ $mb = 1 ] && continue
    line x4reg_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic4d_reg32 --points 50000000
    line x4rect_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload x_cubic4d_rect32 --points 30000000
    line c3c_${lib}_mb$mb INTERPN_B200_LIBRARY=$L/$lib.so INTERPN_B200_QUAD4_MINB=$mb -- --workload c3_cubic4d_rect64 --points 50000000
  done
done
for nt in 1 0; do for t in 8 14; do
  INTERPN_B200_COPY_NT=$nt INTERPN_B200_COPY_THREADS=$t timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustained-steps 0 --suite none 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('NT=$nt threads=$t: pinned', e['value']/1e9, 'pageable', e['pageable']['value']/1e9, e['pageable']['fraction_of_pinned'])"
done; done
timeout 900 python -m pytest tests/test_gpu_host_exec.py tests/test_gpu_instantiations.py -m gpu -x -q 2>&1 | tail -3
