#!/bin/bash
# ncu --set full capture of the evaluation kernel of each listed workload (one launch each).
# Usage (under gpurun): bash tools/gpu_profile.sh <tag> <workload>...
tag=$1; shift
out=gpurun_out/$tag
mkdir -p $out
for wl in "$@"; do
  pts=20000000; [ $wl = c3_cubic4d_rect64 ] && pts=4000000; [ $wl = c1_linear3d_reg20 ] && pts=1000000
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'(cubic|linear|nearest)_kernel' -s 3 -c 1 \
      -o $out/$wl -f python bench.py --workload $wl --points $pts --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/$wl.log 2>&1
  echo "$wl ncu exit $?"
done
