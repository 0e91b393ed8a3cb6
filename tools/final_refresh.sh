bash tools/gpu.sh r2_final2 tests smoke bench:default bench:reference:--impl,reference
out=gpurun_out/r2_final2
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/traffic_c2_cubic3d_reg100_f64.csv python bench.py --workload c2_cubic3d_reg100 --dtype f64 --points 100000000 $B > $out/traffic_c2.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $out/launches_default_bench.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 --suite none > $out/launches_default_bench.log 2>&1
for spec in "c2_coef c2_cubic3d_reg100 100000000" "x4rect_coef x_cubic4d_rect32 30000000" "x3rect_coef x_cubic3d_rect100 50000000"; do
  set -- $spec
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:cubic_quad4 -s 3 -c 1 -o $out/$1 -f python bench.py --workload $2 --points $3 $B > $out/full_$1.log 2>&1
  python tools/ncu_summary.py $out/$1.ncu-rep $out/$1_ncu.json > /dev/null 2>&1
  python tools/ncu_hot.py $out/$1.ncu-rep 16 > $out/$1_hot.txt 2>&1
  rm -f $out/$1.ncu-rep
done
ls $out
