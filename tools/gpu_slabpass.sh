#!/bin/bash
# Slab passes for multilinear on grids a little beyond L2 (launch_linear.cu, kernels.cuh linear_slab_kernel):
# parity tests of the path, then pass size sweep on C3-linear. Usage (under gpurun): bash tools/gpu_slabpass.sh <tag> [pass_kb ...]
tag=${1:-slabpass}; shift; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "slab or c3_linear" > $out/pytest_slab.log 2>&1; echo "pytest exit $?" >> $out/pytest_slab.log
tail -n 4 $out/pytest_slab.log
run() {
  name=$1; wl=$2; pts=$3; shift 3
  env "$@" timeout 600 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/$name.json").read().strip().splitlines()[-1])
    print("$name", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "bit_identical", d["parity"].get("bit_identical"), "launches", d["gpu_launches"])
except Exception as e:
    print("$name FAILED", e); print(open("$out/$name.err").read()[-600:])
PY
}
for kb in ${@:-0 71680 46080 35840}; do
  run c3l_pass$kb c3_linear4d_rect64 100000000 INTERPN_B200_SLAB_PASS_KB=$kb
done
