#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l; nproc
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?" >> gpurun_out/bench_n$N.err
tail -2 gpurun_out/bench_n$N.err
