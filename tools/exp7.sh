#!/bin/bash
# round-2 experiment 7: C4 bin-swept evaluation — chunk size, slab size and the window of keys in flight (L2 hit rate of the evaluation kernel was 55 %)
out=gpurun_out/${TAG:-r2_exp7}; mkdir -p $out
line() { # label env... -- bench args
  label=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 python bench.py "$@" --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --sustained-steps 0 --suite none > $out/$label.json 2> $out/$label.err
  python - "$out/$label.json" "$label" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "%.3f G/s" % (d["value"] / 1e9), "ms %.3f" % d["ms_per_step"], "parity", d["parity"].get("bit_identical"), "launches", d["gpu_launches"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
W="--workload c4_linear6d_reg24 --points 125000000"
line base A=1 -- $W
line chunk125 INTERPN_B200_SWEEP_CHUNK=125000000 -- $W
line chunk125_ctas4 INTERPN_B200_SWEEP_CHUNK=125000000 INTERPN_B200_SWEEP_EVAL_CTAS=4 -- $W
line chunk125_ctas2 INTERPN_B200_SWEEP_CHUNK=125000000 INTERPN_B200_SWEEP_EVAL_CTAS=2 -- $W
line ctas4 INTERPN_B200_SWEEP_EVAL_CTAS=4 -- $W
line ctas2 INTERPN_B200_SWEEP_EVAL_CTAS=2 -- $W
line slab24 INTERPN_B200_SWEEP_SLAB_KB=24576 -- $W
line slab24_ctas4 INTERPN_B200_SWEEP_SLAB_KB=24576 INTERPN_B200_SWEEP_EVAL_CTAS=4 -- $W
line slab24_chunk125 INTERPN_B200_SWEEP_SLAB_KB=24576 INTERPN_B200_SWEEP_CHUNK=125000000 -- $W
line slab12_chunk125 INTERPN_B200_SWEEP_SLAB_KB=12288 INTERPN_B200_SWEEP_CHUNK=125000000 -- $W
line chunk31 INTERPN_B200_SWEEP_CHUNK=31250000 -- $W
W3="--workload c3_linear4d_rect64 --points 100000000"
line c3l_swept INTERPN_B200_SWEEP_MIN_ROWS=4 -- $W3
line c3l_swept_chunk100 INTERPN_B200_SWEEP_MIN_ROWS=4 INTERPN_B200_SWEEP_CHUNK=100000000 -- $W3
line c3l_swept_chunk100_ctas4 INTERPN_B200_SWEEP_MIN_ROWS=4 INTERPN_B200_SWEEP_CHUNK=100000000 INTERPN_B200_SWEEP_EVAL_CTAS=4 -- $W3
