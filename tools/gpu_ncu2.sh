#!/bin/bash
tag=${1:-ncu2}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'nearest_kernel' -s 3 -c 1 -o $out/c5n3_rect -f python bench.py --workload c5_nearest3d_rect128 --points 20000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c5n3_rect.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'quad4' -s 3 -c 1 -o $out/x4 -f python bench.py --workload x_cubic4d_reg32 --points 20000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/x4.log 2>&1; echo "ncu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'sweep_scatter' -s 6 -c 1 -o $out/c4_scatter -f python bench.py --workload c4_linear6d_reg24 --points 100000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c4_scatter.log 2>&1; echo "ncu exit $?"
