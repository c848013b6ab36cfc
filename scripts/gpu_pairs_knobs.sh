#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batch.py -x -q > gpurun_out/pytest_batch.log 2>&1; echo "exit $?" >> gpurun_out/pytest_batch.log
tail -5 gpurun_out/pytest_batch.log
run() { # workload pairs seats engines threads
  GCR_BATCH_STATS=1 GCR_BATCH_THREADS=$5 timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --workload $1 --pairs-api batch --pairs $2 --seats $3 --engines $4 > gpurun_out/knob_$1_$2_$3_$4_$5.log 2>&1
  python - <<PY
import json
v=None
for line in open("gpurun_out/knob_$1_$2_$3_$4_$5.log"):
    if line.startswith("{"): v=json.loads(line)["value"]
print("$1 pairs $2 seats $3 engines $4 threads $5 -> %s pairs/s" % ("%.0f" % v if v else "FAILED"))
PY
}
for w in h2k_pairs cfg1_pairs f4k_pairs; do
  run $w 128 128 1 14
  run $w 128 128 2 14
  run $w 128 128 2 16
  run $w 512 128 1 14
  run $w 512 256 2 14
  run $w 512 512 2 14
  run $w 512 512 4 12
done
run e2k_pairs 64 64 2 14
