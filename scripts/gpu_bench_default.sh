#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/bench_reference.err
timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/bench_default.err; echo "exit $?" >> gpurun_out/bench_default.err
tail -2 gpurun_out/bench_default.err
