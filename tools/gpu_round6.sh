#!/bin/bash
# Session-6 checkpoint: full GPU suite, smoke, default bench line, C3-linear (slab passes) bench line in both arithmetic
# flavours, launch list of its timed region and ncu --set full of one step's slab-pass launches.
# Usage (under gpurun): bash tools/gpu_round6.sh <tag>
tag=${1:-s6_round}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -n 6 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $out/smoke.log
( time timeout 900 python bench.py ) > $out/bench_default.json 2> $out/bench_default.err; tail -c 1500 $out/bench_default.json
line() {
  name=$1; shift
  timeout 600 python bench.py "$@" > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/$name.json").read().strip().splitlines()[-1])
    print("$name", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), "launches", d["gpu_launches"], "e2e", (d.get("e2e") or {}).get("value"))
except Exception as e:
    print("$name FAILED", e); print(open("$out/$name.err").read()[-600:])
PY
}
line c3_linear --workload c3_linear4d_rect64 --points 100000000 --steps 5 --warmup 3 --no-cpu-baseline
line c3_linear_fma --workload c3_linear4d_rect64 --points 100000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --arithmetic fma
INTERPN_B200_SLAB_PASS_KB=0 line c3_linear_direct --workload c3_linear4d_rect64 --points 100000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches_c3_linear.csv python bench.py --workload c3_linear4d_rect64 --points 100000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/c3_under_ncu.log 2>&1
tail -n 4 $out/launches_c3_linear.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_slab -s 9 -c 3 -o $out/c3l_slab -f python bench.py --workload c3_linear4d_rect64 --points 100000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c3l_slab.log 2>&1; echo "ncu exit $?"
