#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for k in 65536 262144; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload e2k --samples $k > gpurun_out/bench_e2k_k$k.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_e2k.csv python bench.py --steps 2 --warmup 1 --no-cpu --workload e2k > gpurun_out/ncu_launch_e.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_e5 -s 1 -c 1 -o gpurun_out/prof_k2_e5 -f python bench.py --steps 2 --warmup 1 --no-cpu --workload e2k > gpurun_out/ncu_full_e5.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
