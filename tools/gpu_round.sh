#!/bin/bash
# One GPU-box session: parity tests, microbenchmarks, bench lines for several workloads/variants.
# Usage (under gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
timeout 300 ./tools/microbench > $out/microbench.json 2>&1
cat $out/microbench.json
for wl in c2_cubic3d_reg100 c1_linear3d_reg20 c3_linear4d_rect64 c3_cubic4d_rect64 c5_nearest3d_reg128 c5_nearest2d_reg1024; do
  pts=100000000; [ $wl = c3_cubic4d_rect64 ] && pts=20000000; [ $wl = c1_linear3d_reg20 ] && pts=1000000
  timeout 600 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_$wl.json 2> $out/bench_$wl.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$wl.json").read().strip().splitlines()[-1])
    print("$wl", "%.3f Gpts/s" % (d["value"]/1e9), "frac %.4f" % d["roofline"]["frac"], "parity", d["parity"], "clk", d["clocks"]["sm_mhz"])
except Exception as e:
    print("$wl FAILED", e)
PY
done
INTERPN_B200_WINDOW_MB=0 timeout 600 python bench.py --workload c2_cubic3d_reg100 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/bench_c2_nowin.json 2> $out/bench_c2_nowin.err
python -c "
import json; d=json.loads(open('$out/bench_c2_nowin.json').read().strip().splitlines()[-1]); print('c2 nowin %.3f Gpts/s'%(d['value']/1e9), d['parity'])"
