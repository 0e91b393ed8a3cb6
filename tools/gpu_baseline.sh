#!/bin/bash
# Round checkpoint on the GPU box: parity tests, smoke, default bench line (both arms), launch list of the
# timed region, and one ncu --set full capture of the headline kernel at the bench's launch size.
# Usage (under gpurun): bash tools/gpu_baseline.sh <tag>
tag=${1:-base}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
nproc > $out/nproc.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -4 $out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; tail -2 $out/smoke.log
( time timeout 900 python bench.py --impl reference ) > $out/bench_reference.json 2> $out/bench_reference.err; tail -c 600 $out/bench_reference.json
( time timeout 900 python bench.py ) > $out/bench_default.json 2> $out/bench_default.err; tail -c 3000 $out/bench_default.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
tail -4 $out/launches.csv | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'cubic_quad' -s 3 -c 1 -o $out/c2_full -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/c2_full.log 2>&1; echo "ncu exit $?"
