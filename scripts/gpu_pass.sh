#!/bin/bash
# GPU pass: tests, smoke, bench (all K3 variants), ncu launch list, ncu full capture of K3 / K2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_f4k.log 2>&1; echo "exit $?" >> gpurun_out/bench_f4k.log
for v in 0 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --variant $v > gpurun_out/bench_f4k_v$v.log 2>&1; done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --exact-counts > gpurun_out/bench_f4k_exact.log 2>&1
for v in 0 1; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload h2k_batch --variant $v > gpurun_out/bench_h2k_v$v.log 2>&1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_f4k.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h2k.csv python bench.py --steps 2 --warmup 1 --no-cpu --workload h2k_batch > gpurun_out/ncu_launch_h.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_main -s 1 -c 1 -o gpurun_out/prof_k3_f4k -f python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_score_main -s 1 -c 1 -o gpurun_out/prof_k3_h2k -f python bench.py --steps 2 --warmup 1 --no-cpu --workload h2k_batch > gpurun_out/ncu_full_h.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve_f7 -s 1 -c 1 -o gpurun_out/prof_k2_f7 -f python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_k2.log 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log
