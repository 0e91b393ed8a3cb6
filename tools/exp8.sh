#!/bin/bash
# round-2 experiment 8: nearest kernels with / without the next-group coordinate prefetch, all C5 variants
out=gpurun_out/${TAG:-r2_exp8}; mkdir -p $out
L=$PWD/interpn_b200
for lib in libinterpn_b200 lib_nopf; do
for wl in c5_nearest2d_reg1024 c5_nearest3d_reg128 c5_nearest2d_rect1024 c5_nearest3d_rect128; do
for dt in f64 f32; do
  INTERPN_B200_LIBRARY=$L/$lib.so timeout 600 python bench.py --workload $wl --dtype $dt --points 200000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 --suite none > $out/${wl}_${dt}_$lib.json 2> $out/${wl}_${dt}_$lib.err
  python - "$out/${wl}_${dt}_$lib.json" "$wl $dt $lib" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "%.2f G/s" % (d["value"] / 1e9), "frac %.3f" % d["roofline"]["frac"], "parity", d["parity"].get("bit_identical"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done; done; done
