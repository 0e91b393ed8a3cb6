#!/bin/bash
# Round-2 profile session (one GPU, under gpurun): DRAM traffic per step of every bench workload (metrics-only ncu pass),
# launch lists, and ncu --set full captures of the dominant kernels. Output: gpurun_out/r2_prof/. Summaries are made
# on the build machine with tools/ncu_summary.py / tools/traffic_table.py and copied to profiles/.
out=gpurun_out/r2_prof; mkdir -p $out
B="--steps 1 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0"
traffic() { # workload dtype points
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/traffic_$1_$2.csv python bench.py --workload $1 --dtype $2 --points $3 $B > $out/traffic_$1_$2.log 2>&1
  echo "traffic $1 $2 exit $?"
}
traffic c2_cubic3d_reg100 f64 100000000
traffic c1_linear3d_reg20 f64 1000000
traffic c3_linear4d_rect64 f64 100000000
traffic c3_cubic4d_rect64 f64 100000000
traffic c4_linear6d_reg24 f64 125000000
for w in c5_nearest2d_reg1024 c5_nearest3d_reg128 c5_nearest2d_rect1024 c5_nearest3d_rect128; do
  traffic $w f64 200000000; traffic $w f32 200000000
done
# launch list of the headline command itself (suite off: the timed region of the default line)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $out/launches_default_bench.csv \
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --sustained-steps 0 --suite none > $out/launches_default_bench.log 2>&1; echo "launch list exit $?"
full() { # name workload dtype points kernel-regex skip count  -> <name>_ncu.json (summary) + <name>_hot.txt (per-instruction view)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$5" -s $6 -c $7 -o $out/$1 -f \
    python bench.py --workload $2 --dtype $3 --points $4 $B > $out/full_$1.log 2>&1; echo "ncu full $1 exit $?"
  python tools/ncu_summary.py $out/$1.ncu-rep $out/$1_ncu.json > /dev/null 2>&1
  python tools/ncu_hot.py $out/$1.ncu-rep 16 > $out/$1_hot.txt 2>&1
  [ -n "$8" ] || rm -f $out/$1.ncu-rep   # gpurun_out/ comes back only below 64 MiB: keep the reports marked "keep"
}
full c2_coef c2_cubic3d_reg100 f64 100000000 cubic_quad4 3 1 keep
full x4reg_coef x_cubic4d_reg32 f64 50000000 cubic_quad4 3 1
full x3rect_coef x_cubic3d_rect100 f64 50000000 cubic_quad4 3 1
full x4rect_coef x_cubic4d_rect32 f64 30000000 cubic_quad4 3 1
full c3c_coef c3_cubic4d_rect64 f64 100000000 cubic_quad4 12 1
full c3l_hyper c3_linear4d_rect64 f64 100000000 linear_hyper 3 1
full c4_hyper c4_linear6d_reg24 f64 125000000 linear_hyper 3 1
full c3c_scatter c3_cubic4d_rect64 f64 100000000 sweep_scatter 3 1
full c5_n2rect_f32 c5_nearest2d_rect1024 f32 200000000 nearest_kernel 3 1
full c5_n3reg_f64 c5_nearest3d_reg128 f64 200000000 nearest_kernel 3 1
full c1_linear c1_linear3d_reg20 f64 1000000 linear_kernel 3 1
du -sh $out; ls $out | head -n 80
