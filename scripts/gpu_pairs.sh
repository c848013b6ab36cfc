#!/bin/bash
# pipeline workloads: host-thread sweep
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for w in 1 2 4 8 16; do
  timeout 600 python bench.py --steps 2 --warmup 1 --workload f4k_pairs --workers $w > gpurun_out/bench_f4k_pairs_w$w.log 2>&1
  timeout 600 python bench.py --steps 2 --warmup 1 --workload h2k_pairs --workers $w > gpurun_out/bench_h2k_pairs_w$w.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log
