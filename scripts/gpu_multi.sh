#!/bin/bash
# multi-GPU check (run with gpurun --gpus N): weak-scaling pairs and the hypothesis-sharded single problem
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --workload f1m_sharded > gpurun_out/bench_f1m_n1.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --workload f1m_sharded > gpurun_out/bench_f1m_n$N.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_f4k_n$N.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --workload h2k_batch > gpurun_out/bench_h2k_n$N.log 2>&1
tail -n 4 gpurun_out/bench_f1m_n1.log gpurun_out/bench_f1m_n$N.log gpurun_out/bench_f4k_n$N.log gpurun_out/bench_h2k_n$N.log | cut -c1-400
