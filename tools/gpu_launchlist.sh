#!/bin/bash
# Per-kernel durations (ncu launch list) of one timed step of the swept workloads.
tag=${1:-ll}; out=gpurun_out/$tag; mkdir -p $out
for spec in c4_linear6d_reg24:100000000 c3_cubic4d_rect64:20000000; do
  wl=${spec%%:*}; pts=${spec##*:}
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches_$wl.csv python bench.py --workload $wl --points $pts --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > $out/$wl.log 2>&1
  python - <<PY
import csv, collections
rows=[r for r in csv.reader(open("$out/launches_$wl.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split("(")[0][:60]; agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=float(r[-1])/1e6
tot=sum(v[1] for v in agg.values())
print("$wl total %.3f ms"%tot)
for k,v in agg.items(): print("  %-60s x%-3d %8.3f ms %5.1f%%"%(k,v[0],v[1],100*v[1]/tot))
PY
done
