#!/bin/bash
# ncu pass: launch lists of the default bench command per workload + full captures of K3 / K2 (tag = $1, default r2)
TAG=${1:-r2}
mkdir -p gpurun_out
for w in f4k h4k h2k_batch e2k; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_$w.csv python bench.py --steps 1 --warmup 1 --rounds 4 --no-cpu --workload $w > gpurun_out/ncu_launch_$w.log 2>&1
done
cap() { # name kernel-regex workload
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o gpurun_out/${TAG}_prof_$1 -f python bench.py --steps 1 --warmup 1 --rounds 2 --no-cpu --workload $3 > gpurun_out/ncu_full_$1.log 2>&1
}
cap k3_f4k k_score_main f4k
cap k3_h4k k_score_main h4k
cap k3_h2k k_score_main h2k_batch
cap k3_e2k k_score_main e2k
cap k2_f7 k_solve_f7 f4k
cap k2_h4 k_solve_h4 h2k_batch
cap k2_e5 k_solve_e5 e2k
ls -la gpurun_out | tail -20
