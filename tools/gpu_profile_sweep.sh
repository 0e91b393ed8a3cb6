#!/bin/bash
# ncu --set full of the evaluation kernel inside the bin-swept path (one launch each).
tag=$1; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "swept" > $out/pytest_sweep.log 2>&1; echo "pytest exit $?" >> $out/pytest_sweep.log; tail -3 $out/pytest_sweep.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'cubic_kernel' -s 4 -c 1 -o $out/c3_cubic_swept -f python bench.py --workload c3_cubic4d_rect64 --points 8000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c3_cubic_swept.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'linear_kernel' -s 4 -c 1 -o $out/c4_linear_swept -f python bench.py --workload c4_linear6d_reg24 --points 8000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c4_linear_swept.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'sweep_|linear_kernel' -s 20 -c 10 --csv --log-file $out/c4_launches.csv python bench.py --workload c4_linear6d_reg24 --points 8000000 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c4_launches.log 2>&1
cat $out/c4_launches.csv | tail -10 | cut -d, -f5,15- | cut -c1-150
