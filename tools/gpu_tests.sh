#!/bin/bash
# GPU parity suite + a few bench lines. Usage (under gpurun): bash tools/gpu_tests.sh <tag> [workload:points ...]
tag=${1:-tests}; shift; out=gpurun_out/$tag; mkdir -p $out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -12 $out/pytest_gpu.log
for spec in "$@"; do
  wl=${spec%%:*}; pts=${spec##*:}
  timeout 900 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/$wl.json 2> $out/$wl.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/$wl.json").read().strip().splitlines()[-1])
    print("$wl", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), "launches", d["gpu_launches"], "swept", d.get("swept_launches"))
except Exception as e:
    print("$wl FAILED", e); print(open("$out/$wl.err").read()[-600:])
PY
done
