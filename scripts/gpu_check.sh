#!/bin/bash
# quick GPU pass: tests, smoke, default bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "exit $?" >> gpurun_out/bench_default.err
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log; tail -c 600 gpurun_out/bench_default.log
