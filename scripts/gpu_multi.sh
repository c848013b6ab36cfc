#!/bin/bash
# multi-GPU pass (gpurun --gpus N): the driver's launch line for every workload that shards
N=${1:-2}
mkdir -p gpurun_out
run() { # workload steps warmup
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps $2 --warmup $3 --no-cpu --workload $1 > gpurun_out/bench_$1_n$N.log 2>&1
}
run f4k 20 3
run h2k_batch 5 3
run f1m_sharded 2 1
if [ "$2" != "short" ]; then
run f4k_pairs 2 1
run h2k_pairs 2 1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_f4k_ref_n$N.log 2>&1
fi
grep -h -o '"value": [0-9.e+]*' gpurun_out/bench_*_n$N.log
