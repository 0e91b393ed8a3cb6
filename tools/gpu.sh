#!/bin/bash
# One parameterised GPU-box runner (replaces the single-use tools/gpu_*.sh of round 1). Run under gpurun:
#   gpurun --timeout 1800 -- 'bash tools/gpu.sh <tag> <step> [<step> ...]'
# Everything is written under gpurun_out/<tag>/. Steps (each is one word, arguments joined with ':'):
#   tests[:pytest-args]             GPU suite (pytest -m gpu -x -q [args]); tests:tests/test_x.py runs one file
#   smoke                           __graft_entry__.smoke()
#   bench[:name[:args...]]          python bench.py [args]  -> <name>.json (default name "bench"); args use ',' for spaces
#   line:<workload>:<points>[:args] one quick bench line (5 steps, no CPU baseline, no e2e) with a one-line summary
#   ll:<workload>:<points>[:args]   ncu launch list of one timed step (gpu__time_duration per kernel) + per-kernel shares
#   ncu:<workload>:<points>:<kernel-regex>[:skip[:count[:args]]]   ncu --set full of <count> launches after <skip>
#   sanitize[:tool]                 compute-sanitizer (memcheck|racecheck|initcheck|synccheck, default all four) over tools/sanitize_run.py
#   env VAR=VALUE                   export a variable for the following steps (e.g. env:INTERPN_B200_ARITHMETIC=fma)
#   sh:<command,with,commas>        any other command
tag=${1:?tag}; shift
out=gpurun_out/$tag; mkdir -p $out
summary() {  # <json file> <label>
python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(sys.argv[2], "%.3f G/s" % (d["value"] / 1e9), "ms %.3f" % d["ms_per_step"], "frac %.4f" % (r.get("frac") or 0), "parity", (d.get("parity") or {}).get("bit_identical"),
          "launches", d.get("gpu_launches"), "e2e", ((d.get("e2e") or {}).get("value") or 0) / 1e9)
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
shares() {  # <launch csv>
python - "$1" <<'PY'
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4].split("(")[0][:70]; agg.setdefault(name, [0, 0.0]); agg[name][0] += 1; agg[name][1] += float(r[-1].replace(",", "")) / 1e6
tot = sum(v[1] for v in agg.values()) or 1.0
print("total %.3f ms in %d launches" % (tot, len(rows)))
for k, v in agg.items(): print("  %-70s x%-4d %9.3f ms %5.1f%%" % (k, v[0], v[1], 100 * v[1] / tot))
PY
}
for step in "$@"; do
  IFS=: read -r kind a b c d e f <<< "$step"
  case $kind in
    env) export "$a"; echo "export $a" ;;
    tests)
      ( time timeout 2400 python -m pytest ${a:-tests} -m gpu -x -q ${b//,/ } ) > $out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/pytest_gpu.log
      tail -n 8 $out/pytest_gpu.log ;;
    smoke) timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $out/smoke.log ;;
    bench)
      name=${a:-bench}
      ( time timeout 1800 python bench.py ${b//,/ } ) > $out/$name.json 2> $out/$name.err; echo "bench $name exit $?"
      summary $out/$name.json $name; tail -n 4 $out/$name.err ;;
    line)
      timeout 900 python bench.py --workload $a --points $b --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 ${c//,/ } > $out/$a.json 2> $out/$a.err
      summary $out/$a.json "$a${c:+ [$c]}" ;;
    ll)
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches_$a.csv \
        python bench.py --workload $a --points $b --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 ${c//,/ } > $out/ll_$a.log 2>&1
      echo "launch list $a:"; shares $out/launches_$a.csv ;;
    ncu)
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$c" -s ${d:-3} -c ${e:-1} -o $out/${a}_ncu -f \
        python bench.py --workload $a --points $b --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 ${f//,/ } > $out/ncu_$a.log 2>&1
      echo "ncu $a exit $?" ;;
    sanitize)
      for tool in ${a:-memcheck racecheck initcheck synccheck}; do
        timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 20 python tools/sanitize_run.py > $out/sanitize_$tool.log 2>&1
        echo "compute-sanitizer $tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize_run:" $out/sanitize_$tool.log | tail -n 3
      done ;;
    sh) cmd=${step#sh:}; bash -c "${cmd//,/ }" 2>&1 | tail -n 20 ;;
    *) echo "unknown step $step" ;;
  esac
done
