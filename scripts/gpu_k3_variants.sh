#!/bin/bash
mkdir -p gpurun_out
for v in 0 3; do for e in "" "--no-exact"; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --variant $v $e > gpurun_out/bench_f4k_v${v}${e}.log 2>&1; done; done
for v in 0 3; do timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload h2k_batch --variant $v > gpurun_out/bench_h2k_v$v.log 2>&1; done
