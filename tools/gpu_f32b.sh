#!/bin/bash
tag=${1:-f32b}; out=gpurun_out/$tag; mkdir -p $out
( time timeout 2400 python -m pytest tests -m gpu -x -q ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
tail -6 $out/pytest_gpu.log
bash tools/gpu_f32.sh $tag
for spec in c5_nearest3d_reg128:100000000 c1_linear3d_reg20:100000000 x_linear3d_reg100:100000000; do
  wl=${spec%%:*}; pts=${spec##*:}
  timeout 900 python bench.py --workload $wl --points $pts --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > $out/f64_$wl.json 2> $out/f64_$wl.err
  python -c "
import json; d=json.loads(open('$out/f64_$wl.json').read().strip().splitlines()[-1]); print('f64 $wl %.3f Gpts/s'%(d['value']/1e9), d['parity'].get('bit_identical'))"
done
