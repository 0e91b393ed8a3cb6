#!/bin/bash
# Full GPU suite (strict + the fma flavour re-run in a child process) and bench lines in both arithmetic flavours.
tag=${1:-fma}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -15 $out/pytest_gpu.log
for ar in strict fma; do
  for spec in c2_cubic3d_reg100:100000000 x_linear3d_reg100:100000000 x_cubic4d_reg32:50000000 x_cubic3d_rect100:50000000 c3_cubic4d_rect64:20000000 c5_nearest3d_reg128:100000000; do
    wl=${spec%%:*}; pts=${spec##*:}
    timeout 900 python bench.py --arithmetic $ar --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/${ar}_$wl.json 2> $out/${ar}_$wl.err
    python - <<PY
import json
try:
    d = json.loads(open("$out/${ar}_$wl.json").read().strip().splitlines()[-1])
    print("$ar $wl", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), d["config"].get("arithmetic"))
except Exception as e:
    print("$ar $wl FAILED", e); print(open("$out/${ar}_$wl.err").read()[-800:])
PY
  done
done
