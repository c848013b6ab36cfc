#!/bin/bash
# GPU pass: tests, smoke, bench (every workload + reference arms), ncu launch lists, ncu full captures of K3 / K2
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_f4k.log 2>&1; echo "exit $?" >> gpurun_out/bench_f4k.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_f4k_ref.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --workload h4k > gpurun_out/bench_h4k.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --workload h2k_batch > gpurun_out/bench_h2k.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --workload e2k > gpurun_out/bench_e2k.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 1 --workload f4k_pairs > gpurun_out/bench_f4k_pairs.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 1 --workload h2k_pairs > gpurun_out/bench_h2k_pairs.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --workload f4k_pairs > gpurun_out/bench_f4k_pairs_ref.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --workload h2k_pairs > gpurun_out/bench_h2k_pairs_ref.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 1 --workload e2k_pairs > gpurun_out/bench_e2k_pairs.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --workload e2k_pairs > gpurun_out/bench_e2k_pairs_ref.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 1 --workload cfg1_pairs > gpurun_out/bench_cfg1_pairs.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --workload cfg1_pairs > gpurun_out/bench_cfg1_pairs_ref.log 2>&1
timeout 300 python tools/k3_trace.py f4k > gpurun_out/k3_trace_f4k.txt 2>&1
for w in f4k h4k h2k_batch e2k; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv python bench.py --steps 2 --warmup 1 --no-cpu --workload $w > gpurun_out/ncu_launch_$w.log 2>&1
done
cap() { # name kernel-regex workload
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$2 -s 1 -c 1 -o gpurun_out/prof_$1 -f python bench.py --steps 2 --warmup 1 --no-cpu --workload $3 > gpurun_out/ncu_full_$1.log 2>&1
}
cap k3_f4k k_score_main f4k
cap k3_h2k k_score_main h2k_batch
cap k3_h4k k_score_main h4k
cap k3_e2k k_score_main e2k
cap k2_f7 k_solve_f7 f4k
cap k2_h4 k_solve_h4 h2k_batch
cap k2_e5 k_solve_e5 e2k
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
