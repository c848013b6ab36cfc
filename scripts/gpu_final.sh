#!/bin/bash
# final pass of a round: tests, smoke, the default bench line + the reference arm, ncu launch lists and full captures
TAG=${1:-r2}
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/bench_reference.err
timeout 1200 python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/bench_default.err; echo "bench exit $?" >> gpurun_out/bench_default.err
bash scripts/gpu_prof.sh $TAG > /dev/null 2>&1
tail -n 3 gpurun_out/pytest_gpu.log gpurun_out/smoke.log gpurun_out/bench_default.err
