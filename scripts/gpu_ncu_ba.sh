#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_refit -s 6 -c 1 -o gpurun_out/r2_prof_k6_fba -f python scripts/prof_pipeline_kernels.py f > gpurun_out/ncu_ba.log 2>&1
tail -3 gpurun_out/ncu_ba.log
