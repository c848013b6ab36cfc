#!/bin/bash
for mode in LAZY EAGER; do
SECONDS=0
CUDA_MODULE_LOADING=$mode python bench.py --no-extras --no-cpu --steps 5 --warmup 3 > /dev/null
echo "$mode bench f4k --no-extras wall $SECONDS s"
done
