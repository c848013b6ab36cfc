#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q > gpurun_out/pytest_batch.log 2>&1; echo "exit $?" >> gpurun_out/pytest_batch.log
tail -30 gpurun_out/pytest_batch.log
bash scripts/gpu_pairs_one.sh "cfg1_pairs h2k_pairs f4k_pairs e2k_pairs" "96 2;64 1;128 4"
