#!/bin/bash
# quick GPU check: tests + one bench line per workload + launch lists
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_f4k.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload h2k_batch > gpurun_out/bench_h2k.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --workload e2k > gpurun_out/bench_e2k.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --workload f4k_pairs > gpurun_out/bench_f4k_pairs.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --workload h2k_pairs > gpurun_out/bench_h2k_pairs.log 2>&1
timeout 600 python bench.py --steps 2 --warmup 1 --workload e2k_pairs > gpurun_out/bench_e2k_pairs.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --workload f4k_pairs > gpurun_out/bench_f4k_pairs_ref.log 2>&1
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 --workload e2k_pairs > gpurun_out/bench_e2k_pairs_ref.log 2>&1
for w in f4k h2k_batch e2k; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$w.csv python bench.py --steps 2 --warmup 1 --no-cpu --workload $w > gpurun_out/ncu_launch_$w.log 2>&1
done
tail -n 5 gpurun_out/pytest_gpu.log
