#!/bin/bash
# round-2 experiment 1: sweep tile/block variants, swept C3-linear, pageable copy threads, f32 rect nearest P
out=gpurun_out/r2_exp1; mkdir -p $out
L=interpn_b200
run() { # label lib extra-env workload points
  env INTERPN_B200_LIBRARY=$2 $3 timeout 600 python bench.py --workload $4 --points $5 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 > $out/$1.json 2> $out/$1.err
  python - "$out/$1.json" "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], "%.3f G/s" % (d["value"] / 1e9), "ms %.3f" % d["ms_per_step"], "parity", d["parity"].get("bit_identical"), "launches", d["gpu_launches"], "swept", d["swept_launches"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for v in "" sweepA sweepB sweepC; do
  lib=$PWD/$L/libinterpn_b200.so; [ -n "$v" ] && lib=$PWD/$L/lib_$v.so
  run c4_$v $lib "A=1" c4_linear6d_reg24 125000000
  run c3c_$v $lib "A=1" c3_cubic4d_rect64 100000000
  run c3l_swept_$v $lib "INTERPN_B200_SWEEP_MIN_ROWS=4" c3_linear4d_rect64 100000000
done
run c3l_slab $PWD/$L/libinterpn_b200.so "A=1" c3_linear4d_rect64 100000000
run c3l_swept_24mb $PWD/$L/libinterpn_b200.so "INTERPN_B200_SWEEP_MIN_ROWS=4 INTERPN_B200_SWEEP_SLAB_KB=24576" c3_linear4d_rect64 100000000
run c3l_swept_96mb $PWD/$L/libinterpn_b200.so "INTERPN_B200_SWEEP_MIN_ROWS=4 INTERPN_B200_SWEEP_SLAB_KB=98304" c3_linear4d_rect64 100000000
for wl in c5_nearest2d_rect1024 c5_nearest3d_rect128; do
  INTERPN_B200_LIBRARY=$PWD/$L/libinterpn_b200.so timeout 600 python bench.py --workload $wl --dtype f32 --points 200000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl f32 P4', d['value']/1e9, d['parity']['bit_identical'])"
  INTERPN_B200_LIBRARY=$PWD/$L/lib_nrf2.so timeout 600 python bench.py --workload $wl --dtype f32 --points 200000000 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$wl f32 P2', d['value']/1e9, d['parity']['bit_identical'])"
done
for t in 8 12 15; do
  INTERPN_B200_COPY_THREADS=$t timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustained-steps 0 --suite none 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('copy threads $t: pinned', e['value']/1e9, 'pageable', e['pageable']['value']/1e9, e['pageable']['fraction_of_pinned'])"
done
for kb in 4096 16384; do
  INTERPN_B200_COPY_THREADS=12 INTERPN_B200_CHUNK_PAGEABLE_KB=$kb timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --sustained-steps 0 --suite none 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('chunk $kb KB x12 threads: pageable', e['pageable']['value']/1e9, e['pageable']['fraction_of_pinned'])"
done
