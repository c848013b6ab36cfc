#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?" >> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_bench_reference_n2.json 2> gpurun_out/bench_ref_n2.err; echo "exit $?" >> gpurun_out/bench_ref_n2.err
tail -2 gpurun_out/bench_ref_n2.err
timeout 600 python -m pytest tests/test_gpu_select.py -q -m gpu 2>&1 | tail -3
