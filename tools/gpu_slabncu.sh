#!/bin/bash
# ncu --set full of one step's slab-pass launches on C3-linear (kernels.cuh linear_slab_kernel). Usage (under gpurun): bash tools/gpu_slabncu.sh <tag> [pass_kb]
tag=${1:-slabncu}; kb=${2:-46080}; out=gpurun_out/$tag; mkdir -p $out
INTERPN_B200_SLAB_PASS_KB=$kb timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_slab -s 9 -c 3 \
    -o $out/c3l_slab -f python bench.py --workload c3_linear4d_rect64 --points 100000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c3l_slab.log 2>&1
echo "ncu exit $?"; tail -n 3 $out/c3l_slab.log
