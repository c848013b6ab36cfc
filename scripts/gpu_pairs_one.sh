#!/bin/bash
# usage: gpu_pairs_one.sh "<workloads>" "<seats engines>;..."
mkdir -p gpurun_out
IFS=';' read -ra CFGS <<< "$2"
for w in $1; do
  for cfg in "${CFGS[@]}"; do
    set -- $cfg
    GCR_BATCH_STATS=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-cpu --workload $w --pairs-api batch --seats $1 --engines $2 > gpurun_out/pairs_${w}_batch_$1_$2.log 2>&1
    echo "== $w seats $1 engines $2"; grep -A5 "batch call" gpurun_out/pairs_${w}_batch_$1_$2.log | tail -6 | cut -c1-700
    python - <<PY
import json
for line in open("gpurun_out/pairs_${w}_batch_$1_$2.log"):
    if line.startswith("{"): print("   -> %.0f pairs/s" % json.loads(line)["value"])
PY
  done
done
