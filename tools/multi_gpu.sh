#!/bin/bash
# Multi-GPU session (under `gpurun --gpus N`): the several-GPUs-behind-one-call test, bench.py at N ranks (per-rank e2e,
# single-call e2e, C4 in the workloads array), and the copy-only microbenchmark. Usage: bash tools/multi_gpu.sh <N> [tag]
N=${1:-2}; tag=${2:-r2_n$N}; out=gpurun_out/$tag; mkdir -p $out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > $out/gpus.csv 2>&1
nvidia-smi topo -m > $out/topo.txt 2>&1
( timeout 900 python -m pytest tests/test_gpu_host_exec.py -m gpu -x -q -rs ) > $out/pytest_host_exec.log 2>&1; tail -n 4 $out/pytest_host_exec.log
timeout 600 python tools/copy_bench.py > $out/copy_bench.json 2> $out/copy_bench.err; cat $out/copy_bench.json | cut -c1-1500
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 ) > $out/bench_n$N.json 2> $out/bench_n$N.err
tail -n 3 $out/bench_n$N.err
python - "$out/bench_n$N.json" <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    e = d["e2e"]
    print("N", d["n_gpus"], "value %.2f G/s" % (d["value"] / 1e9), "sustained %.2f" % (d["sustained"]["value"] / 1e9), "parity", d["parity"]["every_rank_bit_identical"], d["parity"]["ranks_checked"])
    print("  e2e pinned %.2f G/s, pageable %.2f G/s, single_call %s" % (e["value"] / 1e9, e["pageable"]["value"] / 1e9, json.dumps(e.get("single_call"))[:300]))
    for w in d.get("workloads", []):
        print("  ", w["workload"], "%.2f G/s" % (w["value"] / 1e9), "frac %.3f" % w["roofline"]["frac"], w["parity"]["every_rank_bit_identical"])
except Exception as ex:
    print("bench parse failed", ex)
PY
