#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2 3; do
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
done
for rep in 1 2 3 4 5; do
timeout 300 python -m pytest tests/test_gpu_batch.py -q -x -p no:cacheprovider 2>&1 | tail -1
done
