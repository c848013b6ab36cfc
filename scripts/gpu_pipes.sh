#!/bin/bash
mkdir -p gpurun_out
M=sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__pipe_fmalite_cycles_active.sum,sm__pipe_fmaheavy_cycles_active.sum,sm__inst_executed_pipe_fma.sum,sm__pipe_fma_cycles_active.sum,sm__cycles_active.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_issued.sum
ncu --metrics $M --clock-control none -c 8 --csv --log-file gpurun_out/pipes_ubench.csv ./tools/ubench_fma > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k_score_main -s 1 -c 1 --csv --log-file gpurun_out/pipes_k3.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-exact > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:k_score_main -s 1 -c 1 --csv --log-file gpurun_out/pipes_k3_scalar.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-exact --variant 3 > /dev/null 2>&1
