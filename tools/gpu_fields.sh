python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or resident" 2>&1 | tail -3; python -m pytest tests -m gpu -x -q 2>&1 | tail -2; mkdir -p gpurun_out/s5_fields; python tools/bench_fields.py 6 > gpurun_out/s5_fields/fields6.jsonl 2> gpurun_out/s5_fields/err.log; python - <<'PY'
import json
for l in open("gpurun_out/s5_fields/fields6.jsonl"):
    d=json.loads(l); print(d["workload"], "fused %.1f G field-pts/s"%(d["fused_field_points_per_s"]/1e9), "separate %.1f"%(d["separate_field_points_per_s"]/1e9), "x%.2f"%d["speedup"], d["bit_identical"])
PY
tail -3 gpurun_out/s5_fields/err.log
