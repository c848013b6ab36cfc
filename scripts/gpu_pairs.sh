#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python scripts/prof_pairs.py 2>&1 | tail -4 | cut -c1-300
for w in f4k_pairs h2k_pairs e2k_pairs; do
  timeout 600 python bench.py --steps 2 --warmup 1 --workload $w > gpurun_out/bench_$w.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log
